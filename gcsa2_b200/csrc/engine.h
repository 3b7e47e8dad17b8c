/*
  engine.h -- shared between the CUDA translation units of the engine (engine.cu: handles, index creation; find.cu;
  ops.cu: LF / count; locate.cu; lcp.cu: LCPArray queries and the MEM-style scan; kmers.cu): handles, error state,
  the library's memory pool, small host helpers.  Not part of the C ABI.
*/
#ifndef GCSA2_B200_ENGINE_H
#define GCSA2_B200_ENGINE_H

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <omp.h>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/gcsa2_b200.h"
#include "internal.h"

// Device views and primitives (device functions only; each kernel header is included by one translation unit).
#include "device/layout.cuh"

//------------------------------------------------------------------------------
// Errors
//------------------------------------------------------------------------------

extern thread_local std::string g_last_error;     // engine.cu

inline int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
  return fail(GCSA_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } } while(0)

//------------------------------------------------------------------------------
// Stream-ordered temporaries come from a pool of the library's own (one per device, created on first use, never
// trimmed between calls): the host application's default pool and its attributes are left alone.
//------------------------------------------------------------------------------

cudaError_t enginePoolAlloc(void** p, size_t bytes, cudaStream_t stream);      // engine.cu
void enginePoolTrim();                                                         // engine.cu: idle pool memory back to the driver
size_t engineFreeMemory();                                                     // engine.cu: free memory incl. what the pool holds idle
template<class T> inline cudaError_t engineMallocAsync(T** p, size_t bytes, cudaStream_t stream) { return enginePoolAlloc((void**)p, bytes, stream); }

//------------------------------------------------------------------------------
// Host side: handles
//------------------------------------------------------------------------------

/*
  One set of resources for a host-buffer call: SLOTS chunks can be in flight, each with its own stream, device input
  and result buffers, a pinned staging buffer for packed patterns and two events (input copied, results delivered).
  Buffers only grow.  Not shared between concurrent calls (gcsa_b200_index::takePipe / givePipe).
*/
struct HostPipe
{
  static const int SLOTS = 8;
  cudaStream_t stream[SLOTS] = {};
  cudaEvent_t copied[SLOTS] = {}, done[SLOTS] = {};
  void* d_in[SLOTS] = {}; size_t in_bytes[SLOTS] = {};
  void* d_off[SLOTS] = {}; size_t off_bytes[SLOTS] = {};
  void* d_res[SLOTS] = {}; size_t res_bytes[SLOTS] = {};
  static const int STAGING = 4;          // pinned buffers the packers fill (a ring, independent of the slots)
  void* staging[STAGING] = {}; size_t staging_bytes[STAGING] = {};
  cudaEvent_t staged[STAGING] = {};      // the copy engine has read the buffer
  bool staged_used[STAGING] = {};
  bool used[SLOTS] = {};                 // `done` has been recorded at least once
  bool ready = false;

  cudaError_t init()
  {
    if(ready) { return cudaSuccess; }
    for(int s = 0; s < SLOTS; s++)
    {
      cudaError_t e = cudaStreamCreateWithFlags(&stream[s], cudaStreamNonBlocking);
      if(e == cudaSuccess) { e = cudaEventCreateWithFlags(&copied[s], cudaEventDisableTiming); }
      if(e == cudaSuccess) { e = cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming); }
      if(e != cudaSuccess) { return e; }
    }
    for(int b = 0; b < STAGING; b++)
    {
      cudaError_t e = cudaEventCreateWithFlags(&staged[b], cudaEventDisableTiming);
      if(e != cudaSuccess) { return e; }
    }
    ready = true;
    return cudaSuccess;
  }
  static cudaError_t grow(void** p, size_t* have, size_t want, bool pinned)
  {
    if(*have >= want) { return cudaSuccess; }
    if(*p != nullptr) { if(pinned) { cudaFreeHost(*p); } else { cudaFree(*p); } *p = nullptr; *have = 0; }
    cudaError_t e = (pinned ? cudaHostAlloc(p, want, cudaHostAllocDefault) : cudaMalloc(p, want));
    if(e == cudaSuccess) { *have = want; } else { *p = nullptr; }
    return e;
  }
  void destroy()
  {
    for(int s = 0; s < SLOTS; s++)
    {
      if(stream[s]) { cudaStreamSynchronize(stream[s]); cudaStreamDestroy(stream[s]); }
      if(copied[s]) { cudaEventDestroy(copied[s]); }
      if(done[s]) { cudaEventDestroy(done[s]); }
      if(d_in[s]) { cudaFree(d_in[s]); }
      if(d_off[s]) { cudaFree(d_off[s]); }
      if(d_res[s]) { cudaFree(d_res[s]); }
    }
    for(int b = 0; b < STAGING; b++)
    {
      if(staged[b]) { cudaEventDestroy(staged[b]); }
      if(staging[b]) { cudaFreeHost(staging[b]); }
    }
  }
};

struct gcsa_b200_index
{
  int device = 0;
  int sm_count = 148;
  DevView view;
  std::vector<void*> allocations;
  u64 device_bytes = 0;
  gcsa_flat_index header;            // scalars only (pointers nulled)

  // Host-side 2-bit packing of fixed-length patterns (pack.cpp): byte -> comp - 1 or 0xFF, and
  // whether that table is exactly ACGT / acgt.
  u8 pack_code[256];
  bool pack_default = false;

  // Automatic choice between packing and raw copies in the host entry point of find() (GCSA_B200_HOST_PACK unset):
  // seconds per query of the recent large batches either way.  Packing moves fewer bytes over the link but more through
  // host memory (32 B read + 8 written + 8 read + 16 written per 32-mer against 32 + 16 raw), so it wins while the link
  // is the bottleneck and loses when several GPUs share one host's memory system; which one it is shows in the clock.
  mutable std::mutex policy_mutex;
  mutable double policy_seconds[2] = { 0.0, 0.0 };     // [0] raw only, [1] packing shares the batch; 0 = not measured yet
  mutable u64 policy_calls = 0;

  // Resources of the host-buffer entry points (streams, events, device chunk buffers, pinned staging): created on
  // first use, kept for the life of the handle and handed from call to call, one set per concurrent caller.
  mutable std::mutex pool_mutex;
  mutable std::vector<HostPipe*> pipes;
  HostPipe* takePipe() const
  {
    {
      std::lock_guard<std::mutex> lock(pool_mutex);
      if(!pipes.empty()) { HostPipe* p = pipes.back(); pipes.pop_back(); return p; }
    }
    return new HostPipe();
  }
  void givePipe(HostPipe* p) const
  {
    std::lock_guard<std::mutex> lock(pool_mutex);
    pipes.push_back(p);
  }
};

struct gcsa_b200_lcp
{
  int device = 0;
  int sm_count = 148;
  LcpView view;
  void* data = nullptr;
};

//------------------------------------------------------------------------------
// Small host helpers
//------------------------------------------------------------------------------

inline int gridFor(u64 n, int sm_count, int per_sm = 8)
{
  u64 blocks = (n + 255) / 256;
  u64 cap = (u64)sm_count * per_sm;
  return (int)std::max<u64>(1, std::min(blocks, cap));
}

struct DeviceGuard
{
  int prev = 0; bool ok = false;
  explicit DeviceGuard(int device) { ok = (cudaGetDevice(&prev) == cudaSuccess) && (cudaSetDevice(device) == cudaSuccess); }
  ~DeviceGuard() { if(ok) { cudaSetDevice(prev); } }
};


// ---- one process, several GPUs: blocks of a batch, one host thread per handle ----


// Block g of `count` over n items: [first, last)
inline void shardBlock(u64 n, int count, int g, u64* first, u64* last)
{
  u64 base = n / (u64)count, extra = n % (u64)count;
  *first = (u64)g * base + std::min<u64>((u64)g, extra);
  *last = *first + base + ((u64)g < extra ? 1 : 0);
}

// Runs work(g) on one thread per handle; returns the first failure (its message becomes the caller's last error).
template<class Work> int runPerHandle(int count, const char* what, Work work)
{
  std::vector<int> rcs(count, 0);
  std::vector<std::string> errors(count);
  std::vector<std::thread> threads;
  for(int g = 1; g < count; g++)
  {
    try { threads.emplace_back([&, g]() { rcs[g] = work(g); if(rcs[g] != 0) { errors[g] = g_last_error; } }); }
    catch(...) { rcs[g] = GCSA_B200_ERR_NOMEM; errors[g] = std::string(what) + ": cannot start a host thread"; }
  }
  rcs[0] = work(0);
  if(rcs[0] != 0) { errors[0] = g_last_error; }
  for(std::thread& t : threads) { t.join(); }
  for(int g = 0; g < count; g++) { if(rcs[g] != 0) { return fail(rcs[g], errors[g]); } }
  return 0;
}

inline int checkHandles(const gcsa_b200_index* const* indexes, int count, const char* what)
{
  if(indexes == nullptr || count < 1) { return fail(GCSA_B200_ERR_INVALID, std::string(what) + ": no handles"); }
  for(int g = 0; g < count; g++)
  {
    if(indexes[g] == nullptr) { return fail(GCSA_B200_ERR_INVALID, std::string(what) + ": null handle"); }
    if(indexes[g]->header.path_nodes != indexes[0]->header.path_nodes || indexes[g]->header.edge_count != indexes[0]->header.edge_count)
    {
      return fail(GCSA_B200_ERR_INVALID, std::string(what) + ": the handles are not replicas of one index");
    }
  }
  return 0;
}


//------------------------------------------------------------------------------
// Generic host wrapper: copy inputs, run, copy outputs
//------------------------------------------------------------------------------


struct Scratch
{
  cudaStream_t stream = nullptr;
  std::vector<void*> ptrs;
  int init() { return (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess ? 0 : -1); }
  template<class T> T* alloc(u64 count)
  {
    void* p = nullptr;
    if(engineMallocAsync(&p, std::max<u64>(count, 1) * sizeof(T), stream) != cudaSuccess) { return nullptr; }
    ptrs.push_back(p);
    return (T*)p;
  }
  template<class T> T* in(const T* host, u64 count)
  {
    T* p = alloc<T>(count);
    if(p && count) { cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream); }
    return p;
  }
  template<class T> void out(T* host, const T* dev, u64 count)
  {
    if(count) { cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, stream); }
  }
  cudaError_t finish()
  {
    for(void* p : ptrs) { cudaFreeAsync(p, stream); }
    ptrs.clear();
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream); stream = nullptr;
    return e;
  }
};


#define HOST_PROLOGUE(name, handle) \
  if((handle) == nullptr) { return fail(GCSA_B200_ERR_INVALID, name ": null handle"); } \
  DeviceGuard guard((handle)->device); \
  Scratch sc; if(sc.init()) { return fail(GCSA_B200_ERR_CUDA, name ": cannot create stream"); }

#define HOST_EPILOGUE(name, rc) \
  { cudaError_t e_ = sc.finish(); if((rc) != 0) { return (rc); } \
    if(e_ != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string(name ": ") + cudaGetErrorString(e_)); } return 0; }

template<class T> inline int scanExclusive(const T* in, T* out, u64 count, cudaStream_t st)
{
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, st));
  void* tmp = nullptr;
  CUDA_TRY(engineMallocAsync(&tmp, std::max<size_t>(bytes, 16), st));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, st);
  cudaFreeAsync(tmp, st);
  CUDA_TRY(e);
  return 0;
}

#endif
