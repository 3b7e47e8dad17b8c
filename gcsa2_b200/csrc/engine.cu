/*
  engine.cu -- B200 (sm_100a) batched backward-search engine behind the C ABI of
  include/gcsa2_b200.h.  Hand-written CUDA; no tensor cores (integer rank/select work).

  Device layout (DESIGN.md "Data layout in HBM"):

  * FUSED BWT BLOCKS.  The four fast characters (A,C,G,T; fast_bwt[1..4] of
    include/gcsa/gcsa.h:217-219) are interleaved: block b covers path nodes [87b, 87b+87) and is
    one 128-byte line of four 32-byte sectors, one per character c:
        w0 = (C[c] + rank(B_c, 87b))            [40 bits] | B_c bits 64..86   [23 bits] << 40
        w1 = B_c bits 0..63 of the block
        w2 = rank(edges, P - 1)                 [40 bits] | window bits 64..87 [24 bits] << 40
        w3 = window bits 0..63,   window bit t = edges[P - 1 + t],  P = C[c] + rank(B_c, 87b)
    so that one endpoint of GCSA::LF(range, c) (include/gcsa/gcsa.h:155-162, 253-274) -- the B_c
    rank AND the dependent rank on `edges` -- is ONE 32-byte sector read (LDG.E.256) instead of
    two cache-line probes in two vectors, and GCSA::LF(node) (gcsa.h:165-183) is one 128-byte line.
  * RANK VECTORS (edges, sampled_paths, SadaSparse::filter): 32-byte sectors {cumulative count,
    192 data bits}; a rank probe or bit access is one sector.
  * SELECT VECTORS (SadaSparse::values, SadaCount::data): rank vector + one hint per 512 ones.
  * samples/select (gcsa.h:235-236) is replaced by an explicit start-offset array per sampled node.
  * sparse characters ($, N, #; sparse_bwt of gcsa.h:221-223) are sorted position lists.
  * optional k-mer table: find() results of all 4^k ACGT strings of length k, 8 bytes each
    (sp in 40 bits, range length in 24 bits; an empty result always has ep = sp - 1); in its fused
    form 16 bytes: the same entry and the jump-table entry of its path node.
  * optional per-node tables: jump tables for find() (the unary backward path of a node, up to 16 and up
    to 4 steps), walk / locate tables for locate().

  This file is the host side (handles, construction of the layout, C ABI entry points and their
  pipelines); the kernels are in device/*.cuh, included below in dependency order.
*/
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

// Device code: views and primitives, then the kernels by operation.
#include "device/layout.cuh"
#include "device/find.cuh"
#include "device/two_step.cuh"
#include "device/lf_count.cuh"
#include "device/kmers.cuh"
#include "device/locate.cuh"
#include "device/lcp.cuh"
#include "device/mem.cuh"

//------------------------------------------------------------------------------
// Errors
//------------------------------------------------------------------------------

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
  return fail(GCSA_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } } while(0)

//------------------------------------------------------------------------------
// Stream-ordered temporaries come from a pool of the library's own (one per device, created on first use, never
// trimmed between calls): the host application's default pool and its attributes are left alone.
//------------------------------------------------------------------------------

static cudaError_t enginePoolAlloc(void** p, size_t bytes, cudaStream_t stream)
{
  static std::mutex mutex;
  static cudaMemPool_t pools[64] = {};
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if(e != cudaSuccess) { return e; }
  if(device < 0 || device >= 64) { return cudaErrorInvalidValue; }
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(mutex);
    if(pools[device] == nullptr)
    {
      cudaMemPoolProps props;
      std::memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      e = cudaMemPoolCreate(&pools[device], &props);
      if(e != cudaSuccess) { pools[device] = nullptr; return e; }
      uint64_t keep = ~0ull;      // do not hand the memory back to the driver between calls
      cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = pools[device];
  }
  return cudaMallocFromPoolAsync(p, bytes, pool, stream);
}
template<class T> static cudaError_t engineMallocAsync(T** p, size_t bytes, cudaStream_t stream) { return enginePoolAlloc((void**)p, bytes, stream); }

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

namespace {

struct HostBits
{
  const uint64_t* words; u64 n_bits;
  inline u64 word(u64 w) const
  {
    u64 n_words = (n_bits + 63) / 64;
    if(words == nullptr || w >= n_words) { return 0; }
    u64 x = words[w];
    u64 rem = n_bits - w * 64;
    if(rem < 64) { x &= ((1ull << rem) - 1); }
    return x;
  }
  // up to 64 bits starting at bit `start` (bits past the end read as 0)
  inline u64 get(u64 start, u32 len) const
  {
    if(len == 0) { return 0; }
    u64 w = start >> 6; u32 off = start & 63;
    u64 x = word(w) >> off;
    if(off && off + len > 64) { x |= word(w + 1) << (64 - off); }
    if(len < 64) { x &= ((1ull << len) - 1); }
    return x;
  }
};

// ones before each word; cum[n_words] = total
std::vector<u64> wordCum(const HostBits& b)
{
  u64 n_words = (b.n_bits + 63) / 64;
  std::vector<u64> cum(n_words + 2, 0);
  for(u64 w = 0; w < n_words; w++) { cum[w + 1] = cum[w] + __builtin_popcountll(b.word(w)); }
  cum[n_words + 1] = cum[n_words];
  return cum;
}

inline u64 hostRank(const HostBits& b, const std::vector<u64>& cum, u64 i)
{
  if(i >= b.n_bits) { return cum[(b.n_bits + 63) / 64]; }
  u64 w = i >> 6; u32 r = i & 63;
  return cum[w] + (r ? __builtin_popcountll(b.word(w) & ((1ull << r) - 1)) : 0);
}

int upload(gcsa_b200_index* idx, const void* host, size_t bytes, const void** dev)
{
  void* p = nullptr;
  size_t alloc = std::max<size_t>(bytes, 256);
  CUDA_TRY(cudaMalloc(&p, alloc));
  idx->allocations.push_back(p);
  idx->device_bytes += alloc;
  if(bytes) { CUDA_TRY(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice)); }
  *dev = p;
  return 0;
}

int buildRankVec(gcsa_b200_index* idx, const HostBits& b, RankVecDev* out, std::vector<ulonglong4>* keep = nullptr)
{
  u64 n_sec = b.n_bits / RV_W + 1;
  std::vector<ulonglong4> sec(n_sec);
  u64 cum = 0;
  for(u64 s = 0; s < n_sec; s++)
  {
    ulonglong4 q;
    q.x = cum; q.y = b.word(3 * s); q.z = b.word(3 * s + 1); q.w = b.word(3 * s + 2);
    cum += __builtin_popcountll(q.y) + __builtin_popcountll(q.z) + __builtin_popcountll(q.w);
    sec[s] = q;
  }
  const void* d = nullptr;
  int rc = upload(idx, sec.data(), sec.size() * sizeof(ulonglong4), &d);
  if(rc) { return rc; }
  out->sec = (const ulonglong4*)d; out->n_bits = b.n_bits; out->n_sec = n_sec;
  if(keep) { keep->swap(sec); }
  return 0;
}

int buildSelVec(gcsa_b200_index* idx, const HostBits& b, SelVecDev* out)
{
  std::vector<ulonglong4> sec;
  int rc = buildRankVec(idx, b, &out->rv, &sec);
  if(rc) { return rc; }
  u64 ones = 0;
  if(!sec.empty())
  {
    const ulonglong4& q = sec.back();
    ones = q.x + __builtin_popcountll(q.y) + __builtin_popcountll(q.z) + __builtin_popcountll(q.w);
  }
  u64 n_hints = ones / SEL_HINT + 2;
  std::vector<u32> hints(n_hints, (u32)(sec.size() - 1));
  u64 h = 0;
  for(u64 s = 0; s < sec.size() && h < n_hints; s++)
  {
    const ulonglong4& q = sec[s];
    u64 end = q.x + __builtin_popcountll(q.y) + __builtin_popcountll(q.z) + __builtin_popcountll(q.w);
    while(h < n_hints && h * SEL_HINT + 1 <= end) { if(h * SEL_HINT + 1 > q.x) { hints[h] = (u32)s; } h++; }
  }
  const void* d = nullptr;
  rc = upload(idx, hints.data(), hints.size() * sizeof(u32), &d);
  if(rc) { return rc; }
  out->hints = (const u32*)d; out->ones = ones;
  return 0;
}

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

} // namespace

//------------------------------------------------------------------------------
// C ABI
//------------------------------------------------------------------------------

const char* gcsa_b200_last_error(void) { return g_last_error.c_str(); }
void gcsa_b200_internal_set_error(const char* message) { g_last_error = (message != nullptr ? message : ""); }
const char* gcsa_b200_version(void) { return "gcsa2_b200 0.1 (sm_100a; GCSA v3 / LCP v1 semantics of gcsa2 1.3.0)"; }

int gcsa_b200_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void gcsa_b200_free(void* p) { std::free(p); }

int gcsa_b200_index_create(const gcsa_flat_index* host, int device, const gcsa_b200_options* options, gcsa_b200_index** out)
{
  if(host == nullptr || out == nullptr) { return fail(GCSA_B200_ERR_INVALID, "index_create: null argument"); }
  *out = nullptr;
  if(host->sigma != GCSA_B200_SIGMA || host->fast_chars != GCSA_B200_FAST_CHARS)
  {
    return fail(GCSA_B200_ERR_INVALID, "index_create: only the default alphabet (sigma 7, 4 fast characters) is supported");
  }
  if(host->path_nodes >= (1ull << 40) || host->edge_count >= (1ull << 40))
  {
    return fail(GCSA_B200_ERR_INVALID, "index_create: more than 2^40 path nodes or edges");
  }
  int n_dev = gcsa_b200_device_count();
  if(n_dev <= 0) { return fail(GCSA_B200_ERR_CUDA, "index_create: no CUDA device available (this engine has no CPU fallback)"); }
  if(device < 0 || device >= n_dev) { return fail(GCSA_B200_ERR_INVALID, "index_create: bad device ordinal"); }
  DeviceGuard guard(device);
  if(!guard.ok) { return fail(GCSA_B200_ERR_CUDA, "index_create: cudaSetDevice failed"); }

  if(const char* g = std::getenv("GCSA_B200_L2_FETCH"))
  {
    // Random 32-byte sector probes: ask the L2 not to over-fetch neighbouring sectors from HBM.
    size_t bytes = (size_t)std::atoi(g);
    if(bytes == 32 || bytes == 64 || bytes == 128) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, bytes); }
  }
  gcsa_b200_index* idx = new gcsa_b200_index();
  idx->device = device;
  cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, device);
  idx->header = *host;
  for(int c = 0; c < GCSA_B200_SIGMA; c++) { idx->header.bwt[c] = nullptr; }
  idx->header.edges = idx->header.sampled_paths = idx->header.stored_samples = idx->header.samples = nullptr;
  idx->header.extra_filter = idx->header.extra_values = idx->header.redundant = nullptr;

  DevView& v = idx->view;
  std::memset(&v, 0, sizeof(v));
  const u64 N = host->path_nodes;
  v.path_nodes = N; v.edge_count = host->edge_count;
  for(int c = 0; c <= GCSA_B200_SIGMA; c++) { v.C[c] = host->C[c]; }
  std::memcpy(v.char2comp, host->char2comp, 256);
  {
    u8 def[256]; gcsa_b200_default_char2comp(def);
    idx->pack_default = true;
    for(int i = 0; i < 256; i++)
    {
      u8 c = host->char2comp[i];
      bool fast = (c >= 1 && c <= GCSA_B200_FAST_CHARS);
      idx->pack_code[i] = (fast ? (u8)(c - 1) : (u8)0xFF);
      bool def_fast = (def[i] >= 1 && def[i] <= GCSA_B200_FAST_CHARS);
      if(fast != def_fast || (fast && c != def[i])) { idx->pack_default = false; }
    }
    v.default_alphabet = (idx->pack_default ? 1u : 0u);
  }
  for(int i = 0; i < 256; i++) { if(v.char2comp[i] >= GCSA_B200_SIGMA) { delete idx; return fail(GCSA_B200_ERR_INVALID, "index_create: char2comp value out of range"); } }

  int rc = 0;
  #define TRY_RC(expr) do { rc = (expr); if(rc) { gcsa_b200_index_destroy(idx); return rc; } } while(0)

  HostBits edges = { host->edges, host->edge_count };
  std::vector<u64> edge_cum = wordCum(edges);
  TRY_RC(buildRankVec(idx, edges, &v.edges));

  // charRange(comp) = pathNodeRange(C[comp], C[comp+1] - 1), gcsa.h:150-153; C[comp+1] == 0 -> (0, ~0)
  for(int c = 0; c < GCSA_B200_SIGMA; c++)
  {
    if(host->C[c + 1] == 0) { v.char_sp[c] = 0; v.char_ep[c] = ~0ull; }
    else { v.char_sp[c] = hostRank(edges, edge_cum, host->C[c]); v.char_ep[c] = hostRank(edges, edge_cum, host->C[c + 1] - 1); }
  }

  // fused BWT blocks
  {
    u64 n_blocks = N / BWT_W + 1;
    std::vector<ulonglong4> blocks(n_blocks * 4);
    for(int c = 1; c <= GCSA_B200_FAST_CHARS; c++)
    {
      HostBits B = { host->bwt[c], N };
      std::vector<u64> cum = wordCum(B);
      const u64 Cc = host->C[c];
      #pragma omp parallel for schedule(static)
      for(long long bb = 0; bb < (long long)n_blocks; bb++)
      {
        u64 b = (u64)bb, start = b * BWT_W;
        u64 cnt = hostRank(B, cum, start);
        u64 P = Cc + cnt;
        u64 blo = B.get(start, 64), bhi = B.get(start + 64, BWT_W - 64);
        u64 e0, wlo, whi;
        if(P == 0) { e0 = 0; wlo = edges.get(0, 63) << 1; whi = edges.get(63, 24); }
        else { e0 = hostRank(edges, edge_cum, P - 1); wlo = edges.get(P - 1, 64); whi = edges.get(P - 1 + 64, 24); }
        ulonglong4 q;
        q.x = (P & M40) | (bhi << 40); q.y = blo;
        q.z = (e0 & M40) | (whi << 40); q.w = wlo;
        blocks[b * 4 + (c - 1)] = q;
      }
    }
    const void* d = nullptr;
    TRY_RC(upload(idx, blocks.data(), blocks.size() * sizeof(ulonglong4), &d));
    v.bwt = (const ulonglong4*)d;
  }

  // sparse characters
  {
    const int comps[3] = { 0, 5, 6 };
    for(int s = 0; s < 3; s++)
    {
      HostBits B = { host->bwt[comps[s]], N };
      std::vector<u64> pos;
      u64 n_words = (N + 63) / 64;
      for(u64 w = 0; w < n_words; w++)
      {
        u64 x = B.word(w);
        while(x) { pos.push_back(w * 64 + __builtin_ctzll(x)); x &= x - 1; }
      }
      const void* d = nullptr;
      TRY_RC(upload(idx, pos.data(), pos.size() * sizeof(u64), &d));
      v.sparse_pos[s] = (const u64*)d; v.sparse_n[s] = pos.size();
    }
  }

  // samples
  {
    HostBits sampled = { host->sampled_paths, N };
    TRY_RC(buildRankVec(idx, sampled, &v.sampled));
    HostBits last = { host->samples, host->sample_count };
    std::vector<u64> start; start.push_back(0);
    u64 n_words = (host->sample_count + 63) / 64;
    for(u64 w = 0; w < n_words; w++)
    {
      u64 x = last.word(w);
      while(x) { start.push_back(w * 64 + __builtin_ctzll(x) + 1); x &= x - 1; }
    }
    start.push_back(host->sample_count);   // guard entry
    const void* d = nullptr;
    TRY_RC(upload(idx, start.data(), start.size() * sizeof(u64), &d));
    v.sample_start = (const u64*)d;
    TRY_RC(upload(idx, host->stored_samples, host->sample_count * sizeof(u64), &d));
    v.stored_samples = (const u64*)d; v.sample_count = host->sample_count;
  }

  // counting structures
  {
    HostBits filter = { host->extra_filter, N };
    TRY_RC(buildRankVec(idx, filter, &v.extra_filter));
    HostBits values = { host->extra_values, host->extra_values_len };
    TRY_RC(buildSelVec(idx, values, &v.extra_values));
    HostBits red = { host->redundant, host->redundant_len };
    TRY_RC(buildSelVec(idx, red, &v.redundant));
  }

  // locate walk table (optional; 4 or 8 bytes per path node)
  if(N > 0 && (options == nullptr || options->walk_table != 0))
  {
    bool narrow = (N < (1ull << 31));
    size_t bytes = (size_t)N * (narrow ? sizeof(u32) : sizeof(u64));
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    bool forced = (options != nullptr && options->walk_table > 0);
    if(forced || bytes < free_b / 4)
    {
      void* p = nullptr;
      cudaError_t e = cudaMalloc(&p, bytes);
      if(e == cudaSuccess)
      {
        if(narrow) { walk_table_kernel<u32><<<gridFor(N, idx->sm_count, 8), 256>>>(v, (u32*)p); }
        else { walk_table_kernel<u64><<<gridFor(N, idx->sm_count, 8), 256>>>(v, (u64*)p); }
        e = cudaDeviceSynchronize();
      }
      if(e != cudaSuccess)
      {
        if(p) { cudaFree(p); }
        gcsa_b200_index_destroy(idx);
        return fail(GCSA_B200_ERR_CUDA, std::string("walk table: ") + cudaGetErrorString(e));
      }
      idx->allocations.push_back(p); idx->device_bytes += bytes;
      if(narrow) { v.walk32 = (const u32*)p; } else { v.walk64 = (const u64*)p; }

      // Locate table (8 bytes per path node) from the walk table, which it then replaces: one load per
      // located node instead of one per LF step.  walk_table = 2 keeps the walk table instead.
      size_t loc_bytes = (size_t)N * sizeof(u64);
      cudaMemGetInfo(&free_b, &total_b);
      if((options == nullptr || options->walk_table != 2) && (forced || loc_bytes < free_b / 2))
      {
        u64* loc = nullptr; int* overflow = nullptr; int host_overflow = 0;
        e = cudaMalloc((void**)&loc, loc_bytes);
        if(e == cudaSuccess) { e = cudaMalloc((void**)&overflow, sizeof(int)); }
        if(e == cudaSuccess) { e = cudaMemset(overflow, 0, sizeof(int)); }
        if(e == cudaSuccess)
        {
          locate_table_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(v, loc, overflow);
          e = cudaMemcpy(&host_overflow, overflow, sizeof(int), cudaMemcpyDeviceToHost);
        }
        if(overflow) { cudaFree(overflow); }
        if(e == cudaSuccess && host_overflow == 0)
        {
          cudaFree(p); idx->allocations.pop_back(); idx->device_bytes -= bytes;
          v.walk32 = nullptr; v.walk64 = nullptr;
          idx->allocations.push_back(loc); idx->device_bytes += loc_bytes;
          v.loc64 = loc;
        }
        else
        {
          if(loc) { cudaFree(loc); }
          cudaGetLastError();               // keep the walk table
        }
      }
    }
  }

  // jump tables (optional; 8 bytes per path node each + as much again while they are built)
  {
    int want = (options != nullptr ? options->jump_table : 0);        // 0 = automatic, 1 = build, 2 = build with 16-byte entries, -1 = do not
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t bytes = (size_t)N * sizeof(u64);
    u32 tbits = 1; while((1ull << tbits) < N) { tbits++; }
    int max_len = std::min<int>(16, (59 - (int)tbits) / 2);
    // An 8-byte entry holds (59 - tbits) / 2 characters: 16 up to 2^27 path nodes, 13 at 3 G.  Beyond that the long
    // table gets 16-byte entries (16 characters again: a 32-mer is the k-mer table and one jump), memory permitting.
    // (while it is built: level 1, the short table and the 16-byte table, 4 x 8 bytes per node; the k-mer table comes after)
    size_t kmer_table_bytes = 0;
    if(options != nullptr && options->kmer_table_k > 0)
    {
      int tk = std::min(16, options->kmer_table_k);
      kmer_table_bytes = ((size_t)1 << (2 * tk)) * sizeof(u64) + ((size_t)1 << (2 * std::max(1, tk - 1))) * sizeof(ulonglong2);
    }
    bool wide = (want == 2 || (want >= 0 && max_len < 16 && (double)(4 * bytes + kmer_table_bytes) < 0.9 * (double)free_b));
    const int short_len = 4;
    if(N > 0 && want >= 0 && (max_len >= 2 || wide) && (want > 0 || wide || 2 * bytes < free_b / 2))
    {
      u64 *one = nullptr, *table = nullptr, *short_table = nullptr; ulonglong2* wide_table = nullptr;
      cudaError_t e = cudaMalloc((void**)&one, bytes);
      if(e == cudaSuccess) { e = cudaMalloc((void**)&table, bytes); }
      if(e == cudaSuccess && wide) { e = cudaMalloc((void**)&wide_table, 2 * bytes); }
      if(e == cudaSuccess && !wide && max_len > short_len && 3 * bytes < free_b / 2) { if(cudaMalloc((void**)&short_table, bytes) != cudaSuccess) { short_table = nullptr; cudaGetLastError(); } }
      if(e == cudaSuccess)
      {
        jump_init_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(v, tbits, one, table);
        // with 16-byte entries for the long paths, `table` only grows to the short length and IS the short table
        const int grow_to = (wide ? std::min(short_len, max_len) : max_len);
        for(int j = 1; j < grow_to; j++)
        {
          if(j == short_len && short_table != nullptr) { e = cudaMemcpyAsync(short_table, table, bytes, cudaMemcpyDeviceToDevice, 0); }   // paths of up to 4 steps
          jump_extend_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, (u32)j, one, table);
        }
        if(wide)
        {
          jump_wide_init_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, one, wide_table);
          for(int j = 1; j < 16; j++) { jump_wide_extend_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, (u32)j, one, wide_table); }
        }
        if(e == cudaSuccess) { e = cudaDeviceSynchronize(); }
      }
      if(one) { cudaFree(one); }
      if(e == cudaSuccess)
      {
        idx->allocations.push_back(table); idx->device_bytes += bytes;
        v.jump_tbits = tbits;
        if(wide)
        {
          idx->allocations.push_back(wide_table); idx->device_bytes += 2 * bytes;
          v.jump_wide = wide_table; v.jump_short = table; v.jump_k = 16;
        }
        else
        {
          v.jump = table; v.jump_k = (u32)max_len;
          if(short_table != nullptr) { idx->allocations.push_back(short_table); idx->device_bytes += bytes; v.jump_short = short_table; }
        }
      }
      else
      {
        if(table) { cudaFree(table); }
        if(short_table) { cudaFree(short_table); }
        if(wide_table) { cudaFree(wide_table); }
        cudaGetLastError();
        if(want > 0) { gcsa_b200_index_destroy(idx); return fail(GCSA_B200_ERR_CUDA, std::string("jump table: ") + cudaGetErrorString(e)); }
      }
    }
  }

  // two-step blocks (optional): built on the device from the one-step blocks
  // -1 = automatic: worth it once the one-step blocks are far beyond the L2 (the probe rate no longer
  // depends on the footprint there, so halving the probes halves the time; measured in DESIGN.md)
  // (measured at 3 G nodes, profiles/r02_cfg4_3gbp_option_variants.json: with the jump tables the two-step blocks LOSE --
  // the single steps left over are few and the 17.7 GB of extra sectors only dilute the TLB reach -- so the automatic
  // choice takes them only for an index without jump tables)
  bool want_two_step = (options != nullptr && (options->two_step > 0 ||
                        (options->two_step < 0 && N >= 400000000ull && v.jump == nullptr && v.jump_wide == nullptr)));
  if(want_two_step && N > 0)
  {
    u64 n_blocks = N / BWT_W + 1;
    unsigned short* m2 = nullptr; u32* blockpop = nullptr; u64* blockcnt = nullptr; u64* d_base = nullptr; u64* src = nullptr;
    u32* d_viol = nullptr; void* scan_tmp = nullptr; ulonglong4* blocks2 = nullptr;
    cudaError_t e = cudaSuccess;
    auto cleanup2 = [&]() { cudaFree(m2); cudaFree(blockpop); cudaFree(blockcnt); cudaFree(d_base); cudaFree(src); cudaFree(d_viol); cudaFree(scan_tmp); };
    #define TWO_TRY(expr) do { e = (expr); if(e != cudaSuccess) { cleanup2(); if(blocks2) { cudaFree(blocks2); } gcsa_b200_index_destroy(idx); \
      return fail(GCSA_B200_ERR_CUDA, std::string("two-step build: " #expr ": ") + cudaGetErrorString(e)); } } while(0)
    TWO_TRY(cudaMalloc(&m2, (N + 1) * sizeof(unsigned short)));
    TWO_TRY(cudaMalloc(&blockpop, 16 * n_blocks * sizeof(u32)));
    TWO_TRY(cudaMalloc(&blockcnt, 16 * n_blocks * sizeof(u64)));
    TWO_TRY(cudaMalloc(&d_base, 17 * sizeof(u64)));
    TWO_TRY(cudaMalloc(&d_viol, sizeof(u32)));
    TWO_TRY(cudaMemset(d_viol, 0, sizeof(u32)));
    two_step_mask_kernel<<<gridFor(n_blocks, idx->sm_count, 16), 128>>>(v, n_blocks, m2, blockpop);
    TWO_TRY(cudaGetLastError());
    size_t scan_bytes = 0;
    TWO_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, blockpop, blockcnt, n_blocks));
    TWO_TRY(cudaMalloc(&scan_tmp, std::max<size_t>(scan_bytes, 16)));
    u64 base[17]; base[0] = 0;
    for(int p = 0; p < 16; p++)
    {
      TWO_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, blockpop + (u64)p * n_blocks, blockcnt + (u64)p * n_blocks, n_blocks));
      u64 last_cnt = 0; u32 last_pop = 0;
      TWO_TRY(cudaMemcpy(&last_cnt, blockcnt + (u64)p * n_blocks + (n_blocks - 1), sizeof(u64), cudaMemcpyDeviceToHost));
      TWO_TRY(cudaMemcpy(&last_pop, blockpop + (u64)p * n_blocks + (n_blocks - 1), sizeof(u32), cudaMemcpyDeviceToHost));
      base[p + 1] = base[p] + last_cnt + last_pop;
    }
    TWO_TRY(cudaMemcpy(d_base, base, sizeof(base), cudaMemcpyHostToDevice));
    TWO_TRY(cudaMalloc(&src, std::max<u64>(base[16], 1) * sizeof(u64)));
    two_step_source_kernel<<<gridFor(n_blocks, idx->sm_count, 16), 128>>>(v, n_blocks, m2, blockcnt, d_base, src);
    two_step_validate_kernel<<<gridFor(base[16] / 16 + 1, idx->sm_count, 8), 256>>>(src, d_base, d_viol);
    u32 violations = 0;
    TWO_TRY(cudaMemcpy(&violations, d_viol, sizeof(u32), cudaMemcpyDeviceToHost));
    if(violations == 0)
    {
      TWO_TRY(cudaMalloc(&blocks2, n_blocks * 16 * sizeof(ulonglong4)));
      two_step_build_kernel<<<gridFor(n_blocks * 16, idx->sm_count, 8), 256>>>(N, n_blocks, m2, blockcnt, d_base, src, blocks2);
      TWO_TRY(cudaDeviceSynchronize());
      idx->allocations.push_back(blocks2); idx->device_bytes += n_blocks * 16 * sizeof(ulonglong4);
      v.bwt2 = blocks2;
    }
    cleanup2();
    #undef TWO_TRY
  }

  // k-mer table
  int k = (options ? options->kmer_table_k : 0);
  if(k < 0) { k = 0; }
  if(k > 16) { k = 16; }
  if(k > 0 && N > 0)
  {
    u64 entries = 1ull << (2 * k);
    u64 tmp_entries = (k == 1 ? 4 : 1ull << (2 * (k - 1)));
    // Fused form (16 bytes per entry: the jump entry of a singleton result rides along) when there is a jump table
    // and the doubled table still leaves most of the device free; fused_table = 1 forces it, -1 forbids it.
    int want_fused = (options != nullptr ? options->fused_table : 0);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    bool fused = (v.jump != nullptr && want_fused >= 0 &&
                  (want_fused > 0 || (entries + tmp_entries) * sizeof(ulonglong2) < free_b / 10 * 6));
    size_t entry_bytes = (fused ? sizeof(ulonglong2) : sizeof(u64));
    void* p = nullptr; void* tmp = nullptr;
    cudaError_t e = cudaMalloc(&p, entries * entry_bytes);
    if(e == cudaSuccess) { e = cudaMalloc(&tmp, tmp_entries * sizeof(ulonglong2)); }
    if(e != cudaSuccess)
    {
      if(p) { cudaFree(p); }
      gcsa_b200_index_destroy(idx);
      return fail(GCSA_B200_ERR_NOMEM, "index_create: k-mer table allocation failed");
    }
    idx->allocations.push_back(p); idx->device_bytes += entries * entry_bytes;
    table_init_kernel<<<1, 256>>>(v, (ulonglong2*)tmp);
    for(int j = 1; j + 1 < k; j++) { table_extend_kernel<<<gridFor(1ull << (2 * j), idx->sm_count, 8), 256>>>(v, j, (ulonglong2*)tmp); }
    table_final_kernel<<<gridFor(tmp_entries, idx->sm_count, 8), 256>>>(v, k, (const ulonglong2*)tmp, fused ? nullptr : (u64*)p, fused ? (ulonglong2*)p : nullptr);
    e = cudaDeviceSynchronize();
    cudaFree(tmp);
    if(e != cudaSuccess) { gcsa_b200_index_destroy(idx); return fail(GCSA_B200_ERR_CUDA, std::string("k-mer table kernels: ") + cudaGetErrorString(e)); }
    if(fused) { v.table2 = (const ulonglong2*)p; } else { v.table = (const u64*)p; }
    v.table_k = k;
  }
  #undef TRY_RC

  *out = idx;
  return 0;
}

void gcsa_b200_index_destroy(gcsa_b200_index* index)
{
  if(index == nullptr) { return; }
  DeviceGuard guard(index->device);
  for(void* p : index->allocations) { cudaFree(p); }
  for(HostPipe* p : index->pipes) { p->destroy(); delete p; }
  delete index;
}

int gcsa_b200_index_info(const gcsa_b200_index* index, gcsa_b200_info* info)
{
  if(index == nullptr || info == nullptr) { return fail(GCSA_B200_ERR_INVALID, "index_info: null argument"); }
  std::memset(info, 0, sizeof(*info));
  info->path_nodes = index->header.path_nodes; info->edge_count = index->header.edge_count;
  info->order = index->header.order; info->sample_count = index->header.sample_count;
  info->device_bytes = index->device_bytes; info->kmer_table_k = index->view.table_k;
  info->device = index->device; info->sm_count = index->sm_count;
  info->two_step = (index->view.bwt2 != nullptr ? 1 : 0);
  info->jump_k = (index->view.jump != nullptr || index->view.jump_wide != nullptr ? (int)index->view.jump_k : 0);
  info->fused_table = (index->view.table2 != nullptr ? 1 : 0);
  return 0;
}

int gcsa_b200_char_range(const gcsa_b200_index* index, uint64_t comp, uint64_t* sp, uint64_t* ep)
{
  if(index == nullptr || sp == nullptr || ep == nullptr) { return fail(GCSA_B200_ERR_INVALID, "char_range: null argument"); }
  if(comp >= GCSA_B200_SIGMA) { *sp = 0; *ep = ~0ull; return 0; }
  *sp = index->view.char_sp[comp]; *ep = index->view.char_ep[comp];
  return 0;
}

//------------------------------------------------------------------------------
// find
//------------------------------------------------------------------------------

// Measurement hook (not part of the C ABI): how many batches of this process took the two-kernel form.
static std::atomic<unsigned long long> g_fast_launches(0);
extern "C" unsigned long long gcsa_b200_internal_fast_launches(void) { return g_fast_launches.load(); }

static int launchFind(const gcsa_b200_index* index, const u8* d_chars, const u64* d_offsets, u64 char_base, u64 fixed_length,
                      u64 n, u64* d_sp, u64* d_ep, FindStatsDev* d_stats, cudaStream_t stream, bool packed = false)
{
  if(n == 0) { return 0; }
  // persistent grid: 4 CTAs of 256 threads per SM by default (59 registers, no spills; with the packed pattern
  // tail 5 CTAs/SM spill: 13.3 vs 13.1 G queries/s with the 16-mer table, but 7.0 vs 8.4 with the 14-mer table,
  // where more single steps run), one contiguous slice of queries per warp
  static const int min_blocks = []() { const char* e = std::getenv("GCSA_B200_FIND_MINBLOCKS"); return (e ? std::atoi(e) : 4); }();
  int per_sm = (min_blocks >= 8 ? 8 : (min_blocks <= 4 ? 4 : min_blocks));
  int grid = gridFor(n, index->sm_count, per_sm);
  // idle lanes of a warp are refilled together once this many are idle (16 measured best: DESIGN.md)
  static const int refill_at = []() { const char* e = std::getenv("GCSA_B200_FIND_REFILL"); int r = (e ? std::atoi(e) : 16); return std::min(32, std::max(1, r)); }();
  // Batches of k-mers (one length, at least the k-mer table's, the default alphabet): the two-kernel form -- one
  // probe (or two) per thread for everything, then the general kernel for the work list of what that left unfinished.
  static const bool fast_off = []() { const char* e = std::getenv("GCSA_B200_FIND_FAST"); return (e != nullptr && std::atoi(e) == 0); }();
  const DevView& v = index->view;
  if(!fast_off && d_offsets == nullptr && fixed_length <= 255 && v.table_k > 0 && fixed_length >= (u64)v.table_k &&
     (packed || v.default_alphabet != 0) && n >= 4096 && n < (1ull << 47))
  {
    // work list of the general kernel (8 bytes per entry), work list of the quad kernel (16), the two counters
    const bool use_quads = (v.jump_wide != nullptr || v.jump != nullptr);
    // (the 16-byte entries first: the allocation is aligned, the end of an odd number of 8-byte entries is not)
    u64* buffer = nullptr;
    CUDA_TRY(engineMallocAsync(&buffer, n * sizeof(u64) * (use_quads ? 3 : 1) + 256, stream));
    ulonglong2* quad_work = (use_quads ? (ulonglong2*)buffer : nullptr);
    u64* work = buffer + (use_quads ? 2 * n : 0);
    unsigned long long* count = (unsigned long long*)(work + n);
    unsigned long long* quad_count = count + 1;
    cudaError_t e = cudaMemsetAsync(count, 0, 2 * sizeof(unsigned long long), stream);
    if(e == cudaSuccess)
    {
      // queries per thread and round in the first kernel (GCSA_B200_FIND_UNROLL = 1, 2 or 4: measured in DESIGN.md)
      static const int unroll = []() { const char* e = std::getenv("GCSA_B200_FIND_UNROLL"); int u = (e ? std::atoi(e) : 4); return (u == 1 || u == 2 ? u : 4); }();
      int fast_grid = gridFor((n + unroll - 1) / unroll, index->sm_count, 8);
      int slow_grid = gridFor(n, index->sm_count, d_stats ? 1 : 4);
      const u32 L = (u32)fixed_length;
      #define LAUNCH_FAST(S, P, U) find_fast_kernel<S, P, U><<<fast_grid, 256, 0, stream>>>(v, d_chars, L, n, d_sp, d_ep, work, count, quad_work, quad_count, d_stats)
      #define LAUNCH_FAST_U(S, P) do { if(unroll == 1) { LAUNCH_FAST(S, P, 1); } else if(unroll == 2) { LAUNCH_FAST(S, P, 2); } else { LAUNCH_FAST(S, P, 4); } } while(0)
      if(d_stats)
      {
        if(packed) { LAUNCH_FAST(true, true, 4); } else { LAUNCH_FAST(true, false, 4); }
        if(use_quads) { find_quad_kernel<true><<<gridFor(n, index->sm_count, 8), 256, 0, stream>>>(v, L, quad_work, quad_count, d_sp, d_ep, work, count, d_stats); }
        if(packed) { find_kernel<true, 1, true, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, d_stats, refill_at, work, count); }
        else { find_kernel<true, 1, false, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, d_stats, refill_at, work, count); }
      }
      else
      {
        if(packed) { LAUNCH_FAST_U(false, true); } else { LAUNCH_FAST_U(false, false); }
        if(use_quads) { find_quad_kernel<false><<<gridFor(n, index->sm_count, 8), 256, 0, stream>>>(v, L, quad_work, quad_count, d_sp, d_ep, work, count, nullptr); }
        if(packed) { find_kernel<false, 4, true, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, nullptr, refill_at, work, count); }
        else { find_kernel<false, 4, false, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, nullptr, refill_at, work, count); }
      }
      #undef LAUNCH_FAST_U
      #undef LAUNCH_FAST
      e = cudaGetLastError();
    }
    cudaFreeAsync(buffer, stream);
    CUDA_TRY(e);
    g_fast_launches.fetch_add(1);
    return 0;
  }
  #define LAUNCH_FIND(S, B) find_kernel<S, B><<<grid, 256, 0, stream>>>(index->view, d_chars, d_offsets, char_base, fixed_length, n, d_sp, d_ep, d_stats, refill_at)
  if(packed)
  {
    find_kernel<false, 4, true><<<gridFor(n, index->sm_count, 4), 256, 0, stream>>>(index->view, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, nullptr, refill_at);
  }
  else if(d_stats) { LAUNCH_FIND(true, 1); }
  else if(per_sm == 8) { LAUNCH_FIND(false, 8); }
  else if(per_sm == 5) { LAUNCH_FIND(false, 5); }
  else if(per_sm == 4) { LAUNCH_FIND(false, 4); }
  else { LAUNCH_FIND(false, 6); }
  #undef LAUNCH_FIND
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_find_batch(const gcsa_b200_index* index, const uint8_t* d_chars, const uint64_t* d_offsets,
                         uint64_t n, uint64_t* d_sp, uint64_t* d_ep, void* stream)
{
  if(index == nullptr || (n > 0 && (d_chars == nullptr || d_offsets == nullptr || d_sp == nullptr || d_ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_batch: null argument");
  }
  DeviceGuard guard(index->device);
  return launchFind(index, d_chars, (const u64*)d_offsets, 0, 0, n, (u64*)d_sp, (u64*)d_ep, nullptr, (cudaStream_t)stream);
}

int gcsa_b200_find_fixed_batch(const gcsa_b200_index* index, const uint8_t* d_chars, uint64_t pattern_length,
                               uint64_t n, uint64_t* d_sp, uint64_t* d_ep, void* stream)
{
  if(index == nullptr || (n > 0 && (d_chars == nullptr || d_sp == nullptr || d_ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_fixed_batch: null argument");
  }
  DeviceGuard guard(index->device);
  return launchFind(index, d_chars, nullptr, 0, pattern_length, n, (u64*)d_sp, (u64*)d_ep, nullptr, (cudaStream_t)stream);
}

/*
  Host-buffer find.  The batch is cut into chunks that are pipelined over the slots of a HostPipe (the H2D copy of
  one chunk overlaps the kernel of another and the D2H copy of a third).  What bounds this entry point is the H2D
  copy of the patterns (32 pattern bytes in, 16 result bytes out per 32-mer), so fixed-length ACGT batches are also
  2-bit packed on the host (pack.cpp): 4x fewer bytes over the link -- for the chunks the host manages to pack.

  Raw copying and packing SHARE the batch.  The calling thread is the driver -- the only thread that talks to the
  CUDA runtime: it keeps a few raw chunks from the FRONT of the batch queued ahead of the copy engine and sends every
  packed chunk as soon as it is complete.  The other threads of its OpenMP team are packers: they work through the
  chunks the driver opens for them from the BACK of the batch, one sub-block of 8192 patterns at a time, into a ring
  of pinned staging buffers.  The two ends meet wherever the ratio of packing rate to link rate puts them (a packer as
  fast as the link leaves the batch at 0.57 of the raw transfer time, twice as fast at 0.4; with a slow host nearly
  everything goes raw).  Nobody waits for anybody: the first version had a helper thread for the raw copies that the
  packing team starved of a core, the second one packed and enqueued in turns on one thread (measured 7.1 and 5.5 ms
  per 10 M 32-mers, profiles/r02_bench_cfg2_*pack*.json).
  A chunk with any character other than ACGT/acgt is sent raw.
    GCSA_B200_HOST_PACK=0   no packing;   =N   N packing threads;
    unset or "auto"         all OpenMP threads (GCSA_B200_HOST_PACK_THREADS overrides the count).
*/
static int hostPackThreads()
{
  const char* e = std::getenv("GCSA_B200_HOST_PACK");
  if(e != nullptr && *e != 0 && std::strcmp(e, "auto") != 0) { return std::max(0, std::atoi(e)); }
  const char* t = std::getenv("GCSA_B200_HOST_PACK_THREADS");
  int threads = (t != nullptr && *t != 0 ? std::atoi(t) : omp_get_max_threads());
  return std::max(1, threads);
}

static inline void cpuRelax()
{
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}

// Measurement hook (not part of the C ABI): chunks of the last host-buffer find of this process that went packed, and all.
static std::atomic<unsigned long long> g_last_packed_chunks(0), g_last_chunks(0);
extern "C" void gcsa_b200_internal_pack_share(unsigned long long* packed, unsigned long long* total)
{
  if(packed) { *packed = g_last_packed_chunks.load(); }
  if(total) { *total = g_last_chunks.load(); }
}

static int findHost(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets, uint64_t fixed_length,
                    uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats, int pack_threads_override = -1)
{
  if(index == nullptr || (n > 0 && (chars == nullptr || sp == nullptr || ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_host: null argument");
  }
  if(stats) { std::memset(stats, 0, sizeof(*stats)); stats->queries = n; }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);

  const int pack_threads = (pack_threads_override >= 0 ? pack_threads_override : hostPackThreads());
  // (below three chunks of 128 k queries there is nothing to share)
  bool pack = (pack_threads > 0 && fixed_length > 0 && offsets == nullptr && stats == nullptr && n > (2u << 17));
  // No explicit policy: both ways are tried on the first large batches, then the faster one is used, and the other one
  // is tried again every 32nd batch (the load on the host changes).
  const char* policy_env = std::getenv("GCSA_B200_HOST_PACK");
  const bool auto_policy = pack && (policy_env == nullptr || *policy_env == 0 || std::strcmp(policy_env, "auto") == 0);
  if(auto_policy)
  {
    std::lock_guard<std::mutex> lock(index->policy_mutex);
    u64 call = index->policy_calls++;
    if(index->policy_seconds[1] == 0.0) { pack = true; }
    else if(index->policy_seconds[0] == 0.0) { pack = false; }
    else
    {
      bool best = (index->policy_seconds[1] <= index->policy_seconds[0]);
      pack = (call % 32 == 31 ? !best : best);
    }
  }
  const double policy_t0 = omp_get_wtime();
  const bool policy_packed = pack;
  // Chunks of >= 128 k queries (4 MB of 32-mers: the link is at its streaming rate), at most ~24 per batch (48 when
  // packing shares it): the H2D engine is the busy resource from the first byte on, so what the pipeline adds to
  // the transfer time is the kernel and the D2H of the LAST chunk -- the smaller the chunks, the smaller that tail.
  const u64 CHUNK = (pack ? std::max<u64>(1ull << 17, (n + 47) / 48) : std::max<u64>(1ull << 18, (n + 23) / 24));
  const u64 n_chunks = (n + CHUNK - 1) / CHUNK;
  const u64 words_per_pattern = (fixed_length + 31) / 32;
  const int SLOTS = HostPipe::SLOTS;

  HostPipe* pipe = index->takePipe();
  struct Return { const gcsa_b200_index* index; HostPipe* pipe; ~Return() { index->givePipe(pipe); } } give_back = { index, pipe };
  {
    cudaError_t e = pipe->init();
    if(e != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string("find_host: stream creation: ") + cudaGetErrorString(e)); }
  }
  FindStatsDev* d_stats = nullptr;
  if(stats)
  {
    if(cudaMalloc(&d_stats, sizeof(FindStatsDev)) != cudaSuccess || cudaMemset(d_stats, 0, sizeof(FindStatsDev)) != cudaSuccess)
    {
      if(d_stats) { cudaFree(d_stats); }
      return fail(GCSA_B200_ERR_CUDA, "find_host: out of device memory");
    }
  }

  int rc = 0;
  u64 issued = 0;                        // chunks enqueued so far: chunk number k uses slot k % SLOTS
  // One chunk through the next slot: H2D (raw bytes, or the words packed into the slot's staging buffer), kernel, D2H.
  auto enqueue = [&](u64 c, int staging_buffer) -> int
  {
    const bool packed = (staging_buffer >= 0);
    const int slot = (int)(issued % SLOTS);
    cudaStream_t st = pipe->stream[slot];
    u64 q0 = c * CHUNK, q1 = std::min(n, q0 + CHUNK), m = q1 - q0;
    u64 c0 = (offsets ? offsets[q0] : q0 * fixed_length), c1 = (offsets ? offsets[q1] : q1 * fixed_length);
    u64 bytes = (packed ? m * words_per_pattern * sizeof(u64) : c1 - c0);
    #define PIPE_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
      return fail(e_ == cudaErrorMemoryAllocation ? GCSA_B200_ERR_NOMEM : GCSA_B200_ERR_CUDA, std::string("find_host: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)
    if(pipe->used[slot]) { PIPE_TRY(cudaEventSynchronize(pipe->done[slot])); }       // the slot's previous chunk has left its buffers
    PIPE_TRY(HostPipe::grow(&pipe->d_in[slot], &pipe->in_bytes[slot], bytes + 16, false));
    PIPE_TRY(HostPipe::grow(&pipe->d_res[slot], &pipe->res_bytes[slot], 2 * m * sizeof(u64), false));
    if(offsets) { PIPE_TRY(HostPipe::grow(&pipe->d_off[slot], &pipe->off_bytes[slot], (m + 1) * sizeof(u64), false)); }
    u8* d_chars = (u8*)pipe->d_in[slot]; u64* d_off = (offsets ? (u64*)pipe->d_off[slot] : nullptr); u64* d_res = (u64*)pipe->d_res[slot];
    if(packed)
    {
      PIPE_TRY(cudaMemcpyAsync(d_chars, pipe->staging[staging_buffer], bytes, cudaMemcpyHostToDevice, st));
      PIPE_TRY(cudaEventRecord(pipe->staged[staging_buffer], st));       // the buffer may be packed into again
      pipe->staged_used[staging_buffer] = true;
    }
    else if(bytes) { PIPE_TRY(cudaMemcpyAsync(d_chars, chars + c0, bytes, cudaMemcpyHostToDevice, st)); }
    if(offsets) { PIPE_TRY(cudaMemcpyAsync(d_off, offsets + q0, (m + 1) * sizeof(u64), cudaMemcpyHostToDevice, st)); }
    PIPE_TRY(cudaEventRecord(pipe->copied[slot], st));
    int r = launchFind(index, d_chars, d_off, c0, fixed_length, m, d_res, d_res + m, d_stats, st, packed);
    if(r != 0) { return r; }
    PIPE_TRY(cudaMemcpyAsync(sp + q0, d_res, m * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PIPE_TRY(cudaMemcpyAsync(ep + q0, d_res + m, m * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PIPE_TRY(cudaEventRecord(pipe->done[slot], st));
    #undef PIPE_TRY
    pipe->used[slot] = true;
    issued++;
    return 0;
  };

  // Unclaimed chunks are [front, back): raw chunks are claimed from the front, packed ones from the back.  Only the
  // driver claims (it opens chunks for the packers), so front / back need no lock.
  u64 front = 0, back = n_chunks, packed_chunks = 0;
  std::vector<int> h2d_slots;            // slots of the chunks (raw or packed) whose H2D copy may still be queued, oldest first
  auto h2d_queued = [&]() -> size_t
  {
    while(!h2d_slots.empty() && cudaEventQuery(pipe->copied[h2d_slots.front()]) == cudaSuccess) { h2d_slots.erase(h2d_slots.begin()); }
    cudaGetLastError();                  // cudaErrorNotReady is not an error
    return h2d_slots.size();
  };
  auto send = [&](u64 c, int staging_buffer) -> int
  {
    const int slot = (int)(issued % SLOTS);
    int r = enqueue(c, staging_buffer);
    if(r == 0) { h2d_slots.push_back(slot); }
    return r;
  };
  auto send_raw = [&]() -> int
  {
    int r = send(front, -1);
    if(r == 0) { front++; }
    return r;
  };

  const int team = (pack ? std::min(pack_threads + 1, std::max(2, omp_get_max_threads())) : 1);     // the driver and the packers
  if(team < 2)
  {
    while(front < back && rc == 0) { rc = send_raw(); }
  }
  else
  {
    // Packed chunk j (the j-th from the back) is chunk n_chunks - 1 - j and uses staging buffer j % STAGING.
    const int STAGING = HostPipe::STAGING;
    const u64 SUB = 8192;                                            // patterns per work item
    const u64 subs_per_chunk = (CHUNK + SUB - 1) / SUB;
    std::vector<std::atomic<u32>> blocks_done(n_chunks), blocks_bad(n_chunks);
    for(u64 j = 0; j < n_chunks; j++) { blocks_done[j].store(0); blocks_bad[j].store(0); }
    std::atomic<u64> ticket(0), opened(0);
    std::atomic<bool> closing(false);
    for(int b = 0; b < STAGING && rc == 0; b++)
    {
      cudaError_t e = HostPipe::grow(&pipe->staging[b], &pipe->staging_bytes[b], CHUNK * words_per_pattern * sizeof(u64), true);
      if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("find_host: staging buffer: ") + cudaGetErrorString(e)); }
    }
    // A raw chunk is sent only when the copy engine is about to run dry (fewer than this many H2D copies queued,
    // packed ones included): every chunk the packers finish in time crosses the link at a quarter of the bytes, and
    // the raw chunks fill the gaps they leave.  (Keeping raw copies queued regardless gave the raw path half of the
    // batch however fast the packers were: profiles/r02_bench_cfg2_pipe3_pack_*.json.)
    const size_t feed_below = 2;

    #pragma omp parallel num_threads(team)
    {
      if(omp_get_thread_num() != 0)
      {
        // ---- packer: work items (chunk j, sub-block b) in order; wait until the driver has opened chunk j ----
        while(true)
        {
          u64 t = ticket.fetch_add(1), j = t / subs_per_chunk, b = t % subs_per_chunk;
          u32 spins = 0;
          while(j >= opened.load(std::memory_order_acquire) && !closing.load(std::memory_order_acquire))
          {
            if(++spins < 2000) { cpuRelax(); } else { std::this_thread::yield(); }
          }
          if(j >= opened.load(std::memory_order_acquire)) { break; }                       // closing: no more chunks
          u64 c = n_chunks - 1 - j, q0 = c * CHUNK, m = std::min(n, q0 + CHUNK) - q0;
          u64 first = b * SUB, last = std::min(m, first + SUB);
          if(first < last)
          {
            int good = gcsa_b200_internal_pack_range(chars + q0 * fixed_length, first, last, fixed_length, index->pack_code,
                                                     index->pack_default ? 1 : 0, (u64*)pipe->staging[j % STAGING]);
            if(!good) { blocks_bad[j].fetch_add(1, std::memory_order_relaxed); }
          }
          blocks_done[j].fetch_add(1, std::memory_order_release);
        }
      }
      else
      {
        // ---- driver ----
        u64 sent = 0;                                                // packed chunks handed to the copy engine
        while(rc == 0 && (front < back || sent < opened.load(std::memory_order_relaxed)))
        {
          bool progress = false;
          // a packed chunk is complete: send it (raw from the caller's buffer if it held another character)
          if(sent < opened.load(std::memory_order_relaxed) && blocks_done[sent].load(std::memory_order_acquire) == subs_per_chunk)
          {
            bool ok = (blocks_bad[sent].load() == 0);
            rc = send(n_chunks - 1 - sent, ok ? (int)(sent % STAGING) : -1);
            if(rc == 0 && ok) { packed_chunks++; }
            sent++;
            continue;
          }
          // open the next chunk for the packers: one being packed and one waiting is enough to keep them busy, and its
          // staging buffer must have been read by the copy engine (the packed chunk STAGING places before it)
          u64 open_now = opened.load(std::memory_order_relaxed);
          if(front < back && open_now - sent < 2)
          {
            const int buffer = (int)(open_now % STAGING);            // last used by packed chunk open_now - STAGING < sent
            bool free_buffer = true;
            if(pipe->staged_used[buffer])
            {
              if(cudaEventQuery(pipe->staged[buffer]) == cudaSuccess) { pipe->staged_used[buffer] = false; }
              else { free_buffer = false; cudaGetLastError(); }
            }
            if(free_buffer) { back--; opened.store(open_now + 1, std::memory_order_release); progress = true; }
          }
          // keep the copy engine fed
          if(front < back && h2d_queued() < feed_below) { rc = send_raw(); progress = true; }
          if(!progress) { cpuRelax(); }
        }
        closing.store(true, std::memory_order_release);
      }
    }
  }
  g_last_packed_chunks.store(packed_chunks); g_last_chunks.store(n_chunks);

  // everything that was enqueued must have left the caller's buffers before this returns, error or not
  cudaError_t err = cudaSuccess;
  for(int s = 0; s < SLOTS; s++)
  {
    if(pipe->stream[s]) { cudaError_t e = cudaStreamSynchronize(pipe->stream[s]); if(e != cudaSuccess) { err = e; } }
  }
  if(stats && err == cudaSuccess && rc == 0)
  {
    FindStatsDev h;
    err = cudaMemcpy(&h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    stats->found = h.found; stats->total_length = h.total_length; stats->lf_steps = h.lf_steps;
    stats->sector_probes = h.sector_probes; stats->table_hits = h.table_hits;
  }
  if(d_stats) { cudaFree(d_stats); }
  if(rc) { return rc; }
  if(err != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string("find_host: ") + cudaGetErrorString(err)); }
  if(auto_policy)
  {
    // seconds per query of this batch; a moving average over the batches sent the same way
    double per_query = (omp_get_wtime() - policy_t0) / (double)n;
    std::lock_guard<std::mutex> lock(index->policy_mutex);
    double& slot = index->policy_seconds[policy_packed ? 1 : 0];
    slot = (slot == 0.0 ? per_query : 0.75 * slot + 0.25 * per_query);
  }
  return 0;
}

int gcsa_b200_find_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                        uint64_t n, uint64_t* sp, uint64_t* ep)
{
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_host: null offsets"); }
  return findHost(index, chars, offsets, 0, n, sp, ep, nullptr);
}

int gcsa_b200_find_fixed_host(const gcsa_b200_index* index, const uint8_t* chars, uint64_t pattern_length,
                              uint64_t n, uint64_t* sp, uint64_t* ep)
{
  return findHost(index, chars, nullptr, pattern_length, n, sp, ep, nullptr);
}

int gcsa_b200_find_fixed_stats_host(const gcsa_b200_index* index, const uint8_t* chars, uint64_t pattern_length,
                                    uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats)
{
  if(stats == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_fixed_stats_host: null stats"); }
  return findHost(index, chars, nullptr, pattern_length, n, sp, ep, stats);
}

int gcsa_b200_find_stats_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                              uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats)
{
  if(stats == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_stats_host: null stats"); }
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_stats_host: null offsets"); }
  return findHost(index, chars, offsets, 0, n, sp, ep, stats);
}

//------------------------------------------------------------------------------
// One process, several GPUs: the batch is cut into contiguous blocks, one per handle (each handle on its own
// device, the index replicated), and one host thread per handle runs the single-device pipeline on its block,
// writing straight into the caller's arrays.  This is what a caller that parallelises over queries with OpenMP
// threads in one process (src/algorithms.cpp:113, 409; vg) can use; there is no exchange between the devices.
//------------------------------------------------------------------------------

namespace {

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

int checkHandles(const gcsa_b200_index* const* indexes, int count, const char* what)
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

} // namespace

int gcsa_b200_find_fixed_host_multi(const gcsa_b200_index* const* indexes, int count, const uint8_t* chars, uint64_t pattern_length,
                                    uint64_t n, uint64_t* sp, uint64_t* ep)
{
  int rc = checkHandles(indexes, count, "find_fixed_host_multi");
  if(rc != 0) { return rc; }
  if(count == 1) { return findHost(indexes[0], chars, nullptr, pattern_length, n, sp, ep, nullptr); }
  const int per_handle = std::max(1, hostPackThreads() / count);       // the packing threads are shared out
  return runPerHandle(count, "find_fixed_host_multi", [&](int g) -> int
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    if(q0 == q1) { return 0; }
    return findHost(indexes[g], chars + q0 * pattern_length, nullptr, pattern_length, q1 - q0, sp + q0, ep + q0, nullptr,
                    hostPackThreads() == 0 ? 0 : per_handle);
  });
}

int gcsa_b200_find_host_multi(const gcsa_b200_index* const* indexes, int count, const uint8_t* chars, const uint64_t* offsets,
                              uint64_t n, uint64_t* sp, uint64_t* ep)
{
  int rc = checkHandles(indexes, count, "find_host_multi");
  if(rc != 0) { return rc; }
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_host_multi: null offsets"); }
  if(count == 1) { return findHost(indexes[0], chars, offsets, 0, n, sp, ep, nullptr); }
  return runPerHandle(count, "find_host_multi", [&](int g) -> int
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    if(q0 == q1) { return 0; }
    return findHost(indexes[g], chars, offsets + q0, 0, q1 - q0, sp + q0, ep + q0, nullptr);      // offsets stay batch-wide
  });
}

//------------------------------------------------------------------------------
// Generic host wrapper: copy inputs, run, copy outputs
//------------------------------------------------------------------------------

namespace {

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

} // namespace

#define HOST_PROLOGUE(name, handle) \
  if((handle) == nullptr) { return fail(GCSA_B200_ERR_INVALID, name ": null handle"); } \
  DeviceGuard guard((handle)->device); \
  Scratch sc; if(sc.init()) { return fail(GCSA_B200_ERR_CUDA, name ": cannot create stream"); }

#define HOST_EPILOGUE(name, rc) \
  { cudaError_t e_ = sc.finish(); if((rc) != 0) { return (rc); } \
    if(e_ != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string(name ": ") + cudaGetErrorString(e_)); } return 0; }

int gcsa_b200_lf_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                       const uint8_t* d_comp, uint64_t n, uint64_t* d_sp_out, uint64_t* d_ep_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lf_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  lf_kernel<<<gridFor(n, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_sp, (const u64*)d_ep, d_comp, n, (u64*)d_sp_out, (u64*)d_ep_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_lf_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                      const uint8_t* comp, uint64_t n, uint64_t* sp_out, uint64_t* ep_out)
{
  HOST_PROLOGUE("lf_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n); u8* c = sc.in(comp, n);
  u64* oa = sc.alloc<u64>(n); u64* ob = sc.alloc<u64>(n);
  int rc = gcsa_b200_lf_batch(index, a, b, c, n, oa, ob, sc.stream);
  sc.out((u64*)sp_out, oa, n); sc.out((u64*)ep_out, ob, n);
  HOST_EPILOGUE("lf_host", rc);
}

int gcsa_b200_lf_node_batch(const gcsa_b200_index* index, const uint64_t* d_nodes, uint64_t n, uint64_t* d_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lf_node_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  lf_node_kernel<<<gridFor(n, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_nodes, n, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_lf_node_host(const gcsa_b200_index* index, const uint64_t* nodes, uint64_t n, uint64_t* out)
{
  HOST_PROLOGUE("lf_node_host", index);
  u64* a = sc.in((const u64*)nodes, n); u64* o = sc.alloc<u64>(n);
  int rc = gcsa_b200_lf_node_batch(index, a, n, o, sc.stream);
  sc.out((u64*)out, o, n);
  HOST_EPILOGUE("lf_node_host", rc);
}

int gcsa_b200_lf_multi_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                             uint64_t n, int all_chars, uint64_t* d_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lf_multi_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  lf_multi_kernel<<<gridFor(n * GCSA_B200_SIGMA, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_sp, (const u64*)d_ep, n, all_chars, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_lf_multi_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                            uint64_t n, int all_chars, uint64_t* out)
{
  HOST_PROLOGUE("lf_multi_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  u64* o = sc.alloc<u64>(n * GCSA_B200_SIGMA * 2);
  int rc = gcsa_b200_lf_multi_batch(index, a, b, n, all_chars, o, sc.stream);
  sc.out((u64*)out, o, n * GCSA_B200_SIGMA * 2);
  HOST_EPILOGUE("lf_multi_host", rc);
}

int gcsa_b200_count_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                          uint64_t n, uint64_t* d_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "count_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  count_kernel<<<gridFor(n, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_sp, (const u64*)d_ep, n, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_count_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n, uint64_t* out)
{
  HOST_PROLOGUE("count_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n); u64* o = sc.alloc<u64>(n);
  int rc = gcsa_b200_count_batch(index, a, b, n, o, sc.stream);
  sc.out((u64*)out, o, n);
  HOST_EPILOGUE("count_host", rc);
}

//------------------------------------------------------------------------------
// locate
//------------------------------------------------------------------------------

namespace {

template<class T> int scanExclusive(const T* in, T* out, u64 count, cudaStream_t st)
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

/*
  The whole locate pipeline on device buffers.  Outputs: d_out_offsets (n + 1).  If d_values is
  null or capacity is too small, only the sizes are computed and *needed is set.
  Temporaries are stream-ordered allocations.
*/
int locateGeneral(const gcsa_b200_index* index, const u64* d_sp, const u64* d_ep, u64 n,
                  u64* d_out_offsets, u64* d_values, u64 capacity, u64* needed, cudaStream_t st,
                  u64** d_values_alloc = nullptr, bool sorted_unique = true)
{
  const DevView& v = index->view;
  const int sm = index->sm_count;
  std::vector<void*> tmp;
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(engineMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { return nullptr; } tmp.push_back(p); return p; };
  auto cleanup = [&]() { for(void* p : tmp) { cudaFreeAsync(p, st); } tmp.clear(); };
  #define LOC_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { cleanup(); \
    return fail(GCSA_B200_ERR_CUDA, std::string("locate: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)
  #define LOC_RC(expr) do { int rc_ = (expr); if(rc_) { cleanup(); return rc_; } } while(0)

  // 1. nodes per range, exclusive scan
  u64* len = (u64*)alloc((n + 1) * sizeof(u64));
  u64* node_off = (u64*)alloc((n + 1) * sizeof(u64));
  if(!len || !node_off) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
  LOC_TRY(cudaMemsetAsync(len, 0, (n + 1) * sizeof(u64), st));
  locate_lengths_kernel<<<gridFor(n, sm), 256, 0, st>>>(v.path_nodes, d_sp, d_ep, n, len);
  LOC_RC(scanExclusive(len, node_off, n + 1, st));
  u64 items = 0;
  LOC_TRY(cudaMemcpyAsync(&items, node_off + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));

  if(items == 0)
  {
    LOC_TRY(cudaMemsetAsync(d_out_offsets, 0, (n + 1) * sizeof(u64), st));
    if(needed) { *needed = 0; }
    cleanup();
    return 0;
  }

  // 2. walk every node to its sample
  u64* first = (u64*)alloc(items * sizeof(u64));
  u32* steps = (u32*)alloc(items * sizeof(u32));
  u64* cnt = (u64*)alloc((items + 1) * sizeof(u64));
  u64* val_off = (u64*)alloc((items + 1) * sizeof(u64));
  if(!first || !steps || !cnt || !val_off) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
  LOC_TRY(cudaMemsetAsync(cnt + items, 0, sizeof(u64), st));
  locate_walk_kernel<<<gridFor(items, sm), 256, 0, st>>>(v, d_sp, node_off, n, items, first, steps, cnt);
  LOC_RC(scanExclusive(cnt, val_off, items + 1, st));
  u64 total = 0;
  LOC_TRY(cudaMemcpyAsync(&total, val_off + items, sizeof(u64), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));

  // 3. fill, segmented sort, unique
  u64* raw = (u64*)alloc(total * sizeof(u64));
  u64* sorted = (u64*)alloc(total * sizeof(u64));
  u64* seg = (u64*)alloc((n + 1) * sizeof(u64));
  u64* flag = (u64*)alloc((total + 1) * sizeof(u64));
  u64* flag_scan = (u64*)alloc((total + 1) * sizeof(u64));
  if(!raw || !sorted || !seg || !flag || !flag_scan) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
  locate_fill_kernel<<<gridFor(items, sm), 256, 0, st>>>(v, items, first, steps, val_off, raw);
  locate_segments_kernel<<<gridFor(n + 1, sm), 256, 0, st>>>(node_off, val_off, n, seg);
  if(!sorted_unique)
  {
    // sort = false (src/gcsa.cpp:840): the values in the order locateInternal() produces them
    if(needed) { *needed = total; }
    LOC_TRY(cudaMemcpyAsync(d_out_offsets, seg, (n + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    int rc0 = 0;
    if(d_values_alloc != nullptr)
    {
      void* p = nullptr;
      LOC_TRY(engineMallocAsync(&p, std::max<u64>(total, 1) * sizeof(u64), st));
      *d_values_alloc = (u64*)p; d_values = (u64*)p; capacity = total;
    }
    if(d_values == nullptr || capacity < total) { rc0 = GCSA_B200_ERR_CAPACITY; g_last_error = "locate: output capacity too small"; }
    else { LOC_TRY(cudaMemcpyAsync(d_values, raw, total * sizeof(u64), cudaMemcpyDeviceToDevice, st)); }
    cleanup();
    return rc0;
  }
  {
    size_t bytes = 0;
    LOC_TRY(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, raw, sorted, (long long)total, (long long)n, seg, seg + 1, st));
    void* t = alloc(bytes);
    if(!t) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
    LOC_TRY(cub::DeviceSegmentedSort::SortKeys(t, bytes, raw, sorted, (long long)total, (long long)n, seg, seg + 1, st));
  }
  LOC_TRY(cudaMemsetAsync(flag + total, 0, sizeof(u64), st));
  locate_flag_kernel<<<gridFor(total, sm), 256, 0, st>>>(sorted, seg, n, total, flag);
  LOC_RC(scanExclusive(flag, flag_scan, total + 1, st));
  u64 distinct = 0;
  LOC_TRY(cudaMemcpyAsync(&distinct, flag_scan + total, sizeof(u64), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));
  if(needed) { *needed = distinct; }
  locate_offsets_kernel<<<gridFor(n + 1, sm), 256, 0, st>>>(seg, flag_scan, n, total, distinct, d_out_offsets);
  int rc = 0;
  if(d_values_alloc != nullptr)
  {
    void* p = nullptr;
    LOC_TRY(engineMallocAsync(&p, std::max<u64>(distinct, 1) * sizeof(u64), st));
    *d_values_alloc = (u64*)p; d_values = (u64*)p; capacity = distinct;
  }
  if(d_values == nullptr || capacity < distinct) { rc = GCSA_B200_ERR_CAPACITY; g_last_error = "locate: output capacity too small"; }
  else { locate_compact_kernel<<<gridFor(total, sm), 256, 0, st>>>(sorted, flag, flag_scan, total, d_values, capacity); }
  LOC_TRY(cudaGetLastError());
  cleanup();
  return rc;
}

/*
  locate() of a batch of ranges as a CSR of sorted distinct positions.  With the locate table, short ranges are
  answered by the two register passes above (one thread per range) and only the others go through the general
  pipeline; without the table, for sort = false, or with GCSA_B200_LOCATE_SMALL=0 everything does.
*/
int locateDevice(const gcsa_b200_index* index, const u64* d_sp, const u64* d_ep, u64 n,
                 u64* d_out_offsets, u64* d_values, u64 capacity, u64* needed, cudaStream_t st,
                 u64** d_values_alloc = nullptr, bool sorted_unique = true)
{
  const DevView& v = index->view;
  const char* small_env = std::getenv("GCSA_B200_LOCATE_SMALL");
  const bool small_path = (small_env == nullptr || std::atoi(small_env) != 0);
  if(!sorted_unique || v.loc64 == nullptr || !small_path || n == 0)
  {
    return locateGeneral(index, d_sp, d_ep, n, d_out_offsets, d_values, capacity, needed, st, d_values_alloc, sorted_unique);
  }
  const int sm = index->sm_count;
  std::vector<void*> tmp;
  u64* gvals = nullptr;
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(engineMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { return nullptr; } tmp.push_back(p); return p; };
  auto cleanup = [&]() { for(void* p : tmp) { cudaFreeAsync(p, st); } tmp.clear(); if(gvals) { cudaFreeAsync(gvals, st); gvals = nullptr; } };

  u64* cnt = (u64*)alloc((n + 1) * sizeof(u64));
  u64* stash = (u64*)alloc(n * sizeof(u64));
  u64* glist = (u64*)alloc(n * sizeof(u64));
  ull* d_general = (ull*)alloc(sizeof(ull));
  if(!cnt || !stash || !glist || !d_general) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
  LOC_TRY(cudaMemsetAsync(cnt + n, 0, sizeof(u64), st));
  LOC_TRY(cudaMemsetAsync(d_general, 0, sizeof(ull), st));
  locate_small_count_kernel<<<gridFor(n, sm), 256, 0, st>>>(v, d_sp, d_ep, n, cnt, stash, glist, d_general);
  ull n_general = 0;
  LOC_TRY(cudaMemcpyAsync(&n_general, d_general, sizeof(ull), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));
  if(std::getenv("GCSA_B200_LOCATE_DEBUG") != nullptr) { std::fprintf(stderr, "locate: %llu of %llu ranges through the general pipeline\n", n_general, (ull)n); }

  u64* goffs = nullptr;
  if(n_general > 0)
  {
    u64* gsp = (u64*)alloc(n_general * sizeof(u64));
    u64* gep = (u64*)alloc(n_general * sizeof(u64));
    goffs = (u64*)alloc((n_general + 1) * sizeof(u64));
    if(!gsp || !gep || !goffs) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
    locate_general_gather_kernel<<<gridFor(n_general, sm), 256, 0, st>>>(d_sp, d_ep, glist, n_general, gsp, gep);
    u64 gneeded = 0;
    LOC_RC(locateGeneral(index, gsp, gep, n_general, goffs, nullptr, 0, &gneeded, st, &gvals, true));
    locate_general_counts_kernel<<<gridFor(n_general, sm), 256, 0, st>>>(glist, goffs, n_general, cnt);
  }
  LOC_RC(scanExclusive(cnt, d_out_offsets, n + 1, st));
  u64 distinct = 0;
  LOC_TRY(cudaMemcpyAsync(&distinct, d_out_offsets + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));
  if(needed) { *needed = distinct; }
  int rc = 0;
  if(d_values_alloc != nullptr)
  {
    void* p = nullptr;
    LOC_TRY(engineMallocAsync(&p, std::max<u64>(distinct, 1) * sizeof(u64), st));
    *d_values_alloc = (u64*)p; d_values = (u64*)p; capacity = distinct;
  }
  if(d_values == nullptr || capacity < distinct) { rc = GCSA_B200_ERR_CAPACITY; g_last_error = "locate: output capacity too small"; }
  else if(distinct > 0) { locate_small_fill_kernel<<<gridFor(n, sm), 256, 0, st>>>(v, d_sp, d_ep, n, d_out_offsets, stash, goffs, gvals, d_values); }
  LOC_TRY(cudaGetLastError());
  cleanup();
  #undef LOC_TRY
  #undef LOC_RC
  return rc;
}

} // namespace

int gcsa_b200_locate_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                           uint64_t* d_out_offsets, uint64_t* d_values, uint64_t capacity, uint64_t* needed, void* stream)
{
  if(index == nullptr || d_out_offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "locate_batch: null argument"); }
  DeviceGuard guard(index->device);
  if(n == 0)
  {
    CUDA_TRY(cudaMemsetAsync(d_out_offsets, 0, sizeof(u64), (cudaStream_t)stream));
    if(needed) { *needed = 0; }
    return 0;
  }
  return locateDevice(index, (const u64*)d_sp, (const u64*)d_ep, n, (u64*)d_out_offsets, (u64*)d_values, capacity, (u64*)needed, (cudaStream_t)stream);
}

static int locateHost(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                      uint64_t* out_offsets, uint64_t** values, bool sorted_unique);

int gcsa_b200_locate_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                          uint64_t* out_offsets, uint64_t** values)
{
  return locateHost(index, sp, ep, n, out_offsets, values, true);
}

int gcsa_b200_locate_raw_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                              uint64_t* out_offsets, uint64_t** values)
{
  return locateHost(index, sp, ep, n, out_offsets, values, false);
}

static int locateHost(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                      uint64_t* out_offsets, uint64_t** values, bool sorted_unique)
{
  if(out_offsets == nullptr || values == nullptr) { return fail(GCSA_B200_ERR_INVALID, "locate_host: null argument"); }
  *values = nullptr;
  HOST_PROLOGUE("locate_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  u64* offs = sc.alloc<u64>(n + 1);
  u64 needed = 0;
  u64* d_vals = nullptr;
  int rc = 0;
  if(n == 0) { out_offsets[0] = 0; *values = (uint64_t*)std::malloc(sizeof(u64)); }
  else
  {
    rc = locateDevice(index, a, b, n, offs, nullptr, 0, &needed, sc.stream, &d_vals, sorted_unique);
    if(rc == 0)
    {
      u64* vals = (u64*)std::malloc(std::max<u64>(needed, 1) * sizeof(u64));
      if(d_vals != nullptr) { sc.out(vals, d_vals, needed); sc.ptrs.push_back(d_vals); }
      sc.out((u64*)out_offsets, offs, n + 1);
      *values = (uint64_t*)vals;
    }
  }
  HOST_EPILOGUE("locate_host", rc);
}

namespace {
__global__ void __launch_bounds__(256)
add_base_kernel(u64* __restrict__ x, u64 n, u64 base)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) { x[i] += base; }
}
} // namespace

/*
  locate() into caller-owned host buffers (pinned memory makes the copies run at PCIe speed): the batch is cut
  into chunks on two streams -- the ranges of chunk i+1 go up and the values of chunk i-1 come down while
  chunk i is being located.  Same CSR as gcsa_b200_locate_host.
*/
int gcsa_b200_locate_into_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                               uint64_t* out_offsets, uint64_t* values, uint64_t capacity, uint64_t* needed)
{
  if(index == nullptr || out_offsets == nullptr || (n > 0 && (sp == nullptr || ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "locate_into_host: null argument");
  }
  if(needed) { *needed = 0; }
  out_offsets[0] = 0;
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  // Three streams: a stream's next upload queues behind its previous chunk's D2H, so with two streams the H2D engine
  // idles for the length of a locate + D2H every other chunk; with three the uploads run back to back.
  const int STREAMS = 3;
  const u64 CHUNK = std::max<u64>(1ull << 18, (n + 11) / 12);
  const u64 n_chunks = (n + CHUNK - 1) / CHUNK;
  cudaStream_t streams[STREAMS];
  for(int s = 0; s < STREAMS; s++) { CUDA_TRY(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking)); }
  struct Chunk { u64* d_sp = nullptr; u64* d_ep = nullptr; u64* d_offs = nullptr; };
  std::vector<Chunk> chunks(n_chunks);
  int rc = 0;
  bool overflow = false;
  u64 base = 0;
  auto upload = [&](u64 c) -> int
  {
    cudaStream_t st = streams[c % STREAMS];
    u64 q0 = c * CHUNK, m = std::min(n, q0 + CHUNK) - q0;
    Chunk& ch = chunks[c];
    if(engineMallocAsync((void**)&ch.d_sp, m * sizeof(u64), st) != cudaSuccess || engineMallocAsync((void**)&ch.d_ep, m * sizeof(u64), st) != cudaSuccess ||
       engineMallocAsync((void**)&ch.d_offs, (m + 1) * sizeof(u64), st) != cudaSuccess)
    {
      return fail(GCSA_B200_ERR_NOMEM, "locate_into_host: out of device memory");
    }
    cudaMemcpyAsync(ch.d_sp, sp + q0, m * sizeof(u64), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(ch.d_ep, ep + q0, m * sizeof(u64), cudaMemcpyHostToDevice, st);
    return 0;
  };
  rc = upload(0);
  for(u64 c = 0; c < n_chunks && rc == 0; c++)
  {
    cudaStream_t st = streams[c % STREAMS];
    u64 q0 = c * CHUNK, m = std::min(n, q0 + CHUNK) - q0;
    if(c + 1 < n_chunks) { rc = upload(c + 1); if(rc) { break; } }
    Chunk& ch = chunks[c];
    u64 need = 0; u64* d_vals = nullptr;
    rc = locateDevice(index, ch.d_sp, ch.d_ep, m, ch.d_offs, nullptr, 0, &need, st, &d_vals, true);
    if(rc) { break; }
    bool last = (c + 1 == n_chunks);
    add_base_kernel<<<gridFor(m + 1, index->sm_count), 256, 0, st>>>(ch.d_offs, m + 1, base);
    cudaMemcpyAsync(out_offsets + q0, ch.d_offs, (m + (last ? 1 : 0)) * sizeof(u64), cudaMemcpyDeviceToHost, st);
    if(values != nullptr && base + need <= capacity)
    {
      if(need > 0) { cudaMemcpyAsync(values + base, d_vals, need * sizeof(u64), cudaMemcpyDeviceToHost, st); }
    }
    else if(need > 0) { overflow = true; }
    if(d_vals) { cudaFreeAsync(d_vals, st); }
    cudaFreeAsync(ch.d_sp, st); cudaFreeAsync(ch.d_ep, st); cudaFreeAsync(ch.d_offs, st);
    ch = Chunk();
    base += need;
  }
  for(Chunk& ch : chunks)            // an upload that never ran (error path)
  {
    if(ch.d_sp) { cudaFree(ch.d_sp); } if(ch.d_ep) { cudaFree(ch.d_ep); } if(ch.d_offs) { cudaFree(ch.d_offs); }
  }
  cudaError_t err = cudaSuccess;
  for(int s = 0; s < STREAMS; s++)
  {
    cudaError_t e = cudaStreamSynchronize(streams[s]);
    if(e != cudaSuccess) { err = e; }
    cudaStreamDestroy(streams[s]);
  }
  if(rc) { return rc; }
  if(err != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string("locate_into_host: ") + cudaGetErrorString(err)); }
  if(needed) { *needed = base; }
  if(overflow) { return fail(GCSA_B200_ERR_CAPACITY, "locate_into_host: output capacity too small"); }
  return 0;
}

/*
  The same CSR from several GPUs (one handle per device, see gcsa_b200_find_fixed_host_multi).  The place of a block's
  values in the caller's buffer depends on the sizes of the blocks before it, so there are two rounds: count() of
  every range (GCSA::count is exactly the size of the sorted distinct locate() result, src/gcsa.cpp:802-809) into the
  offsets array, then locate() of every block straight into its final place.
*/
int gcsa_b200_locate_into_host_multi(const gcsa_b200_index* const* indexes, int count, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                                     uint64_t* out_offsets, uint64_t* values, uint64_t capacity, uint64_t* needed)
{
  int rc = checkHandles(indexes, count, "locate_into_host_multi");
  if(rc != 0) { return rc; }
  if(count == 1) { return gcsa_b200_locate_into_host(indexes[0], sp, ep, n, out_offsets, values, capacity, needed); }
  if(out_offsets == nullptr || (n > 0 && (sp == nullptr || ep == nullptr))) { return fail(GCSA_B200_ERR_INVALID, "locate_into_host_multi: null argument"); }
  if(needed) { *needed = 0; }
  out_offsets[0] = 0;
  if(n == 0) { return 0; }
  std::vector<u64> total(count, 0), base(count + 1, 0);
  rc = runPerHandle(count, "locate_into_host_multi", [&](int g) -> int
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    if(q0 == q1) { return 0; }
    int r = gcsa_b200_count_host(indexes[g], sp + q0, ep + q0, q1 - q0, out_offsets + q0 + 1);
    if(r != 0) { return r; }
    u64 sum = 0;
    for(u64 q = q0; q < q1; q++) { sum += out_offsets[q + 1]; }
    total[g] = sum;
    return 0;
  });
  if(rc != 0) { return rc; }
  for(int g = 0; g < count; g++) { base[g + 1] = base[g] + total[g]; }
  if(needed) { *needed = base[count]; }
  if(values == nullptr || base[count] > capacity)
  {
    // the offsets are complete either way: prefix sums of the counts
    u64 sum = 0;
    for(u64 q = 0; q < n; q++) { sum += out_offsets[q + 1]; out_offsets[q + 1] = sum; }
    return fail(GCSA_B200_ERR_CAPACITY, "locate_into_host_multi: output capacity too small");
  }
  std::vector<u64> got(count, 0);
  rc = runPerHandle(count, "locate_into_host_multi", [&](int g) -> int
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    if(q0 == q1) { return 0; }
    // block-local offsets into out_offsets[q0 .. q1]; the entry at q1 is also the first of the next block and is set below
    return gcsa_b200_locate_into_host(indexes[g], sp + q0, ep + q0, q1 - q0, out_offsets + q0, values + base[g], total[g], &got[g]);
  });
  if(rc != 0) { return rc; }
  for(int g = 0; g < count; g++)
  {
    if(got[g] != total[g]) { return fail(GCSA_B200_ERR_INCONSISTENT, "locate_into_host_multi: count() and locate() disagree on the size of a block"); }
  }
  #pragma omp parallel for schedule(static)
  for(int g = 0; g < count; g++)
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    out_offsets[q0] = base[g];
    for(u64 q = q0 + 1; q < q1; q++) { out_offsets[q] += base[g]; }
  }
  out_offsets[n] = base[count];
  return 0;
}

/*
  GCSA::locate(range, max_positions, results), src/gcsa.cpp:844-878, batched.  count() runs on the
  device; ranges with max >= total/2 are located in full on the device; the others draw positions
  with std::mt19937_64(sp ^ ep) exactly like the reference, one draw per unfinished range per
  round, and each round's nodes are located as one device batch.
*/
int gcsa_b200_locate_max_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                              uint64_t max_positions, uint64_t* out_offsets, uint64_t** values)
{
  if(index == nullptr || out_offsets == nullptr || values == nullptr) { return fail(GCSA_B200_ERR_INVALID, "locate_max_host: null argument"); }
  *values = nullptr;
  std::vector<u64> totals(n);
  int rc = gcsa_b200_count_host(index, sp, ep, n, (uint64_t*)totals.data());
  if(rc) { return rc; }

  // Only ranges that draw random positions or end up with more than max_positions results need the
  // reference's random machinery (rng(sp ^ ep), the draw loop, deterministicShuffle); everything else
  // is a plain locate().
  struct Special { std::mt19937_64 rng; std::unordered_set<u64> found; std::vector<u64> result; u64 draws = 0; };
  std::unordered_map<u64, Special> special;
  std::vector<u64> full_sp, full_ep, full_id, rnd_id;
  for(u64 i = 0; i < n; i++)
  {
    if(totals[i] == 0) { continue; }
    u64 max_i = std::min<u64>(max_positions, totals[i]);
    if(max_i >= totals[i] / 2) { full_sp.push_back(sp[i]); full_ep.push_back(ep[i]); full_id.push_back(i); }   // gcsa.cpp:860
    else { rnd_id.push_back(i); special[i].rng.seed(sp[i] ^ ep[i]); }             // gcsa.cpp:857
  }
  std::vector<u64> full_offs(full_id.size() + 1, 0);
  uint64_t* full_vals = nullptr;
  if(!full_id.empty())
  {
    rc = gcsa_b200_locate_host(index, (const uint64_t*)full_sp.data(), (const uint64_t*)full_ep.data(), full_id.size(), (uint64_t*)full_offs.data(), &full_vals);
    if(rc) { return rc; }
    // count() may be off for a range that is not a suffix-tree node, so "too many results" (gcsa.cpp:873)
    // is decided on what locate() returned; the generator is untouched until the shuffle on this path.
    for(u64 t = 0; t < full_id.size(); t++)
    {
      u64 i = full_id[t];
      if(full_offs[t + 1] - full_offs[t] > std::min<u64>(max_positions, totals[i]))
      {
        Special& state = special[i];
        state.rng.seed(sp[i] ^ ep[i]);
        state.result.assign(full_vals + full_offs[t], full_vals + full_offs[t + 1]);
      }
    }
  }
  // The reference's loop never ends when count() overestimates the distinct values of a range that
  // is not a suffix-tree node; after 16 * length + 1024 draws the whole range is located instead
  // (the CPU checker used by the tests does the same).
  std::vector<u64> giveup;
  while(!rnd_id.empty())
  {
    std::vector<u64> nodes, active;
    for(u64 t = 0; t < rnd_id.size(); t++)
    {
      u64 i = rnd_id[t];
      Special& state = special[i];
      if(state.draws++ >= 16 * (ep[i] + 1 - sp[i]) + 1024) { giveup.push_back(i); continue; }
      nodes.push_back(sp[i] + state.rng() % (ep[i] + 1 - sp[i]));                 // gcsa.cpp:866
      active.push_back(i);
    }
    if(active.empty()) { break; }
    std::vector<u64> offs(active.size() + 1); uint64_t* vals = nullptr;
    rc = gcsa_b200_locate_host(index, (const uint64_t*)nodes.data(), (const uint64_t*)nodes.data(), active.size(), (uint64_t*)offs.data(), &vals);
    if(rc) { std::free(full_vals); return rc; }
    std::vector<u64> still;
    for(u64 t = 0; t < active.size(); t++)
    {
      u64 i = active[t];
      Special& state = special[i];
      for(u64 j = offs[t]; j < offs[t + 1]; j++) { state.found.insert(vals[j]); }
      if(state.found.size() < std::min<u64>(max_positions, totals[i])) { still.push_back(i); }
      else { state.result.assign(state.found.begin(), state.found.end()); }
    }
    std::free(vals);
    rnd_id.swap(still);
  }
  if(!giveup.empty())
  {
    std::vector<u64> gsp, gep;
    for(u64 i : giveup) { gsp.push_back(sp[i]); gep.push_back(ep[i]); }
    std::vector<u64> offs(giveup.size() + 1); uint64_t* vals = nullptr;
    rc = gcsa_b200_locate_host(index, (const uint64_t*)gsp.data(), (const uint64_t*)gep.data(), giveup.size(), (uint64_t*)offs.data(), &vals);
    if(rc) { std::free(full_vals); return rc; }
    for(u64 t = 0; t < giveup.size(); t++)
    {
      Special& state = special[giveup[t]];
      for(u64 j = offs[t]; j < offs[t + 1]; j++) { state.found.insert(vals[j]); }
      state.result.assign(state.found.begin(), state.found.end());
    }
    std::free(vals);
  }
  for(auto& entry : special)
  {
    std::vector<u64>& r = entry.second.result;
    u64 max_i = std::min<u64>(max_positions, totals[entry.first]);
    if(r.size() > max_i)
    {
      std::sort(r.begin(), r.end());                        // deterministicShuffle, utils.h:359-370
      for(u64 j = r.size(); j > 0; j--) { std::swap(r[j - 1], r[entry.second.rng() % j]); }
      r.resize(max_i);
    }
    std::sort(r.begin(), r.end());
  }
  // assemble: plain ranges straight from the full locate, special ones from their state
  out_offsets[0] = 0;
  {
    u64 t = 0;
    for(u64 i = 0; i < n; i++)
    {
      while(t < full_id.size() && full_id[t] < i) { t++; }
      auto it = special.find(i);
      u64 size = 0;
      if(it != special.end()) { size = it->second.result.size(); }
      else if(t < full_id.size() && full_id[t] == i) { size = full_offs[t + 1] - full_offs[t]; }
      out_offsets[i + 1] = out_offsets[i] + size;
    }
  }
  u64* vals = (u64*)std::malloc(std::max<u64>(out_offsets[n], 1) * sizeof(u64));
  {
    u64 t = 0;
    for(u64 i = 0; i < n; i++)
    {
      while(t < full_id.size() && full_id[t] < i) { t++; }
      auto it = special.find(i);
      if(it != special.end()) { std::copy(it->second.result.begin(), it->second.result.end(), vals + out_offsets[i]); }
      else if(t < full_id.size() && full_id[t] == i) { std::copy(full_vals + full_offs[t], full_vals + full_offs[t + 1], vals + out_offsets[i]); }
    }
  }
  std::free(full_vals);
  *values = (uint64_t*)vals;
  return 0;
}

//------------------------------------------------------------------------------
// LCP
//------------------------------------------------------------------------------

int gcsa_b200_lcp_create(const gcsa_flat_lcp* host, int device, gcsa_b200_lcp** out)
{
  if(host == nullptr || out == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lcp_create: null argument"); }
  *out = nullptr;
  if(host->levels + 1 > 16 || host->levels == 0 || host->branching < 2) { return fail(GCSA_B200_ERR_INVALID, "lcp_create: bad tree shape"); }
  int n_dev = gcsa_b200_device_count();
  if(n_dev <= 0) { return fail(GCSA_B200_ERR_CUDA, "lcp_create: no CUDA device available (this engine has no CPU fallback)"); }
  if(device < 0 || device >= n_dev) { return fail(GCSA_B200_ERR_INVALID, "lcp_create: bad device ordinal"); }
  DeviceGuard guard(device);
  gcsa_b200_lcp* l = new gcsa_b200_lcp();
  l->device = device;
  cudaDeviceGetAttribute(&l->sm_count, cudaDevAttrMultiProcessorCount, device);
  LcpView& v = l->view;
  std::memset(&v, 0, sizeof(v));
  v.size = host->size; v.branching = host->branching; v.levels = host->levels;
  for(u64 i = 0; i <= host->levels; i++) { v.offsets[i] = host->offsets[i]; }
  for(u64 i = host->levels + 1; i < 16; i++) { v.offsets[i] = ~0ull; }
  v.values = host->offsets[host->levels];
  v.shift = -1;
  if((host->branching & (host->branching - 1)) == 0) { v.shift = 0; while((1ull << v.shift) < host->branching) { v.shift++; } }
  cudaError_t e = cudaMalloc(&l->data, ((std::max<u64>(v.values, 16) + 15) / 8) * 8);      // whole 8-byte words (the scans read words)
  if(e == cudaSuccess) { e = cudaMemset(l->data, 0xFF, ((std::max<u64>(v.values, 16) + 15) / 8) * 8); }
  if(e == cudaSuccess && v.values) { e = cudaMemcpy(l->data, host->data, v.values, cudaMemcpyHostToDevice); }
  if(e != cudaSuccess) { if(l->data) { cudaFree(l->data); } delete l; return fail(GCSA_B200_ERR_CUDA, std::string("lcp_create: ") + cudaGetErrorString(e)); }
  v.data = (const u8*)l->data;
  *out = l;
  return 0;
}

void gcsa_b200_lcp_destroy(gcsa_b200_lcp* lcp)
{
  if(lcp == nullptr) { return; }
  DeviceGuard guard(lcp->device);
  cudaFree(lcp->data);
  delete lcp;
}

int gcsa_b200_parent_batch(const gcsa_b200_lcp* lcp, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                           gcsa_b200_stnode* d_out, void* stream)
{
  if(lcp == nullptr) { return fail(GCSA_B200_ERR_INVALID, "parent_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(lcp->device);
  parent_kernel<<<gridFor(n, lcp->sm_count), 256, 0, (cudaStream_t)stream>>>(lcp->view, (const u64*)d_sp, (const u64*)d_ep, n, d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_parent_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n, gcsa_b200_stnode* out)
{
  HOST_PROLOGUE("parent_host", lcp);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  gcsa_b200_stnode* o = sc.alloc<gcsa_b200_stnode>(n);
  int rc = gcsa_b200_parent_batch(lcp, a, b, n, o, sc.stream);
  sc.out(out, o, n);
  HOST_EPILOGUE("parent_host", rc);
}

int gcsa_b200_depth_batch(const gcsa_b200_lcp* lcp, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                          uint64_t* d_out, void* stream)
{
  if(lcp == nullptr) { return fail(GCSA_B200_ERR_INVALID, "depth_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(lcp->device);
  depth_kernel<<<gridFor(n, lcp->sm_count), 256, 0, (cudaStream_t)stream>>>(lcp->view, (const u64*)d_sp, (const u64*)d_ep, n, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_depth_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n, uint64_t* out)
{
  HOST_PROLOGUE("depth_host", lcp);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n); u64* o = sc.alloc<u64>(n);
  int rc = gcsa_b200_depth_batch(lcp, a, b, n, o, sc.stream);
  sc.out((u64*)out, o, n);
  HOST_EPILOGUE("depth_host", rc);
}

int gcsa_b200_lcp_sv_host(const gcsa_b200_lcp* lcp, int which, const uint64_t* pos, uint64_t n,
                          uint64_t* out_pos, uint64_t* out_val)
{
  if(which < 0 || which > 3) { return fail(GCSA_B200_ERR_INVALID, "lcp_sv_host: which must be 0..3"); }
  HOST_PROLOGUE("lcp_sv_host", lcp);
  u64* a = sc.in((const u64*)pos, n); u64* op = sc.alloc<u64>(n); u64* ov = sc.alloc<u64>(n);
  int rc = 0;
  if(n) { lcp_sv_kernel<<<gridFor(n, lcp->sm_count), 256, 0, sc.stream>>>(lcp->view, which, a, n, op, ov); }
  sc.out((u64*)out_pos, op, n); sc.out((u64*)out_val, ov, n);
  HOST_EPILOGUE("lcp_sv_host", rc);
}

int gcsa_b200_lcp_rmq_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                           uint64_t* out_pos, uint64_t* out_val)
{
  HOST_PROLOGUE("lcp_rmq_host", lcp);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  u64* op = sc.alloc<u64>(n); u64* ov = sc.alloc<u64>(n);
  int rc = 0;
  if(n) { lcp_rmq_kernel<<<gridFor(n, lcp->sm_count), 256, 0, sc.stream>>>(lcp->view, a, b, n, op, ov); }
  sc.out((u64*)out_pos, op, n); sc.out((u64*)out_val, ov, n);
  HOST_EPILOGUE("lcp_rmq_host", rc);
}



//------------------------------------------------------------------------------
// countKMers
//------------------------------------------------------------------------------

/*
  countKMers(index, k, parameters), src/algorithms.cpp:387-421: the number of distinct k-mers over
  the bases (include_Ns: bases and N).  The reference walks the trie depth-first, one OpenMP task
  per 5-mer seed; here every level of the trie is one frontier expanded by one kernel launch.
  If ranges != NULL, *ranges receives the final frontier (malloc'ed sp[0..count) then ep[0..count)).
*/
int gcsa_b200_count_kmers(const gcsa_b200_index* index, uint64_t k, int include_Ns, uint64_t* result, uint64_t** ranges)
{
  if(index == nullptr || result == nullptr) { return fail(GCSA_B200_ERR_INVALID, "count_kmers: null argument"); }
  *result = 0;
  if(ranges) { *ranges = nullptr; }
  if(k == 0) { *result = 1; return 0; }
  if(index->header.path_nodes == 0) { return 0; }
  DeviceGuard guard(index->device);
  cudaStream_t st;
  CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const u32 chars = (include_Ns ? GCSA_B200_SIGMA - 2 : GCSA_B200_FAST_CHARS);
  u64 n = 1;
  u64 *sp = nullptr, *ep = nullptr;
  cudaError_t e = cudaSuccess;
  int rc = 0;
  #define KM_TRY(expr) do { e = (expr); if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("count_kmers: " #expr ": ") + cudaGetErrorString(e)); goto done; } } while(0)
  // Only the number is wanted: the frontier need not stay ordered, one kernel per level (kmer_level_kernel).  The
  // frontier of a level is at most `chars` times the one before, and the two buffers are sized for that.
  static const bool unordered_off = []() { const char* s_ = std::getenv("GCSA_B200_KMERS_ORDERED"); return (s_ != nullptr && std::atoi(s_) != 0); }();
  ulonglong2 *cur = nullptr, *next_buf = nullptr; unsigned long long* counter = nullptr;
  if(ranges == nullptr && !unordered_off)
  {
    u64 cur_capacity = 0, next_capacity = 0;
    {
      ulonglong2 root = make_ulonglong2(0, index->header.path_nodes - 1);
      KM_TRY(engineMallocAsync(&cur, sizeof(ulonglong2), st)); cur_capacity = 1;
      KM_TRY(engineMallocAsync(&counter, sizeof(unsigned long long), st));
      KM_TRY(cudaMemcpyAsync(cur, &root, sizeof(root), cudaMemcpyHostToDevice, st));
      for(u64 level = 0; level < k && n > 0; level++)
      {
        u64 want = n * chars;
        if(next_buf == nullptr || next_capacity < want)
        {
          if(next_buf) { cudaFreeAsync(next_buf, st); next_buf = nullptr; }
          KM_TRY(engineMallocAsync(&next_buf, want * sizeof(ulonglong2), st)); next_capacity = want;
        }
        KM_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
        kmer_level_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>(index->view, cur, n, chars, next_buf, counter, next_capacity);
        unsigned long long produced = 0;
        KM_TRY(cudaMemcpyAsync(&produced, counter, sizeof(produced), cudaMemcpyDeviceToHost, st));
        KM_TRY(cudaStreamSynchronize(st));
        std::swap(cur, next_buf); std::swap(cur_capacity, next_capacity);
        n = produced;
      }
      *result = n;
    }
    goto done;
  }
  {
    u64 root[2] = { 0, index->header.path_nodes - 1 };
    KM_TRY(engineMallocAsync(&sp, sizeof(u64), st)); KM_TRY(engineMallocAsync(&ep, sizeof(u64), st));
    KM_TRY(cudaMemcpyAsync(sp, &root[0], sizeof(u64), cudaMemcpyHostToDevice, st));
    KM_TRY(cudaMemcpyAsync(ep, &root[1], sizeof(u64), cudaMemcpyHostToDevice, st));
    for(u64 level = 0; level < k && n > 0; level++)
    {
      u64 total = n * chars;
      u64 *csp = nullptr, *cep = nullptr, *flag = nullptr, *pos = nullptr;
      KM_TRY(engineMallocAsync(&csp, total * sizeof(u64), st)); KM_TRY(engineMallocAsync(&cep, total * sizeof(u64), st));
      KM_TRY(engineMallocAsync(&flag, (total + 1) * sizeof(u64), st)); KM_TRY(engineMallocAsync(&pos, (total + 1) * sizeof(u64), st));
      KM_TRY(cudaMemsetAsync(flag + total, 0, sizeof(u64), st));
      kmer_expand_kernel<<<gridFor(total, index->sm_count), 256, 0, st>>>(index->view, sp, ep, n, chars, csp, cep, flag);
      rc = scanExclusive(flag, pos, total + 1, st);
      if(rc) { goto done; }
      u64 next = 0;
      KM_TRY(cudaMemcpyAsync(&next, pos + total, sizeof(u64), cudaMemcpyDeviceToHost, st));
      KM_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(sp, st); cudaFreeAsync(ep, st); sp = ep = nullptr;
      KM_TRY(engineMallocAsync(&sp, std::max<u64>(next, 1) * sizeof(u64), st)); KM_TRY(engineMallocAsync(&ep, std::max<u64>(next, 1) * sizeof(u64), st));
      kmer_compact_kernel<<<gridFor(total, index->sm_count), 256, 0, st>>>(csp, cep, flag, pos, total, sp, ep);
      cudaFreeAsync(csp, st); cudaFreeAsync(cep, st); cudaFreeAsync(flag, st); cudaFreeAsync(pos, st);
      n = next;
    }
    *result = n;
    if(ranges && n > 0)
    {
      u64* out = (u64*)std::malloc(2 * n * sizeof(u64));
      KM_TRY(cudaMemcpyAsync(out, sp, n * sizeof(u64), cudaMemcpyDeviceToHost, st));
      KM_TRY(cudaMemcpyAsync(out + n, ep, n * sizeof(u64), cudaMemcpyDeviceToHost, st));
      *ranges = (uint64_t*)out;
    }
  }
done:
  if(sp) { cudaFreeAsync(sp, st); }
  if(ep) { cudaFreeAsync(ep, st); }
  if(cur) { cudaFreeAsync(cur, st); }
  if(next_buf) { cudaFreeAsync(next_buf, st); }
  if(counter) { cudaFreeAsync(counter, st); }
  e = cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  #undef KM_TRY
  if(rc == 0 && e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("count_kmers: ") + cudaGetErrorString(e)); }
  return rc;
}

/*
  compareKMers(left, right, k, parameters), src/algorithms.cpp:535-616: result = (kmers in both, only
  in left, only in right).  The reference walks both tries depth-first in lockstep, one OpenMP task per
  5-mer seed; here every level is one frontier of (left range, right range) states expanded by one launch.
  Both indexes must live on the same device.
*/
int gcsa_b200_compare_kmers(const gcsa_b200_index* left, const gcsa_b200_index* right, uint64_t k, int include_Ns,
                            uint64_t* result, gcsa_b200_kmer_state** left_kmers, gcsa_b200_kmer_state** right_kmers)
{
  if(left == nullptr || right == nullptr || result == nullptr) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: null argument"); }
  result[0] = result[1] = result[2] = 0;
  if(left_kmers) { *left_kmers = nullptr; }
  if(right_kmers) { *right_kmers = nullptr; }
  if(k == 0) { result[0] = 1; return 0; }                                         // algorithms.cpp:540
  if(k > 64) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: comparison is only supported for k <= 64"); }   // KMerComparisonState::MAX_K
  if(left->device != right->device) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: the indexes live on different devices"); }
  if(left->header.path_nodes == 0 && right->header.path_nodes == 0) { return 0; }
  const bool want = (left_kmers != nullptr || right_kmers != nullptr);
  HOST_PROLOGUE("compare_kmers", left);
  cudaStream_t st = sc.stream;
  const u32 chars = (include_Ns ? GCSA_B200_SIGMA - 2 : GCSA_B200_FAST_CHARS);
  int rc = 0;
  u64 n = 1;
  u64 root[4] = { 0, left->header.path_nodes - 1, 0, right->header.path_nodes - 1 };
  u64 zero_kmer[3] = { 0, 0, 0 };
  u64* state = nullptr; u64* kmer = nullptr; unsigned long long* counter = nullptr;
  u64 stride = 1;                          // entries per array of `state`
  cudaError_t e = cudaSuccess;
  #define CK_TRY(expr) do { e = (expr); if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("compare_kmers: " #expr ": ") + cudaGetErrorString(e)); goto done; } } while(0)
  #define CK_ALLOC(ptr, count) do { CK_TRY(engineMallocAsync((void**)&(ptr), std::max<u64>((count), 1) * sizeof(u64), st)); } while(0)
  {
    CK_ALLOC(state, 4); CK_TRY(cudaMemcpyAsync(state, root, sizeof(root), cudaMemcpyHostToDevice, st));
    if(want) { CK_ALLOC(kmer, 3); CK_TRY(cudaMemcpyAsync(kmer, zero_kmer, sizeof(zero_kmer), cudaMemcpyHostToDevice, st)); }
    CK_TRY(engineMallocAsync((void**)&counter, sizeof(unsigned long long), st));
    for(u64 level = 0; level < k && n > 0; level++)
    {
      // one kernel per level: the next frontier holds at most `chars` children per state
      u64 capacity = n * chars;
      u64 *new_state = nullptr, *new_kmer = nullptr;
      CK_ALLOC(new_state, 4 * capacity);
      if(want) { CK_ALLOC(new_kmer, 3 * capacity); }
      CK_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
      compare_level_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(left->view, right->view, state, n, stride, kmer, chars, level,
                                                                        new_state, capacity, new_kmer, counter);
      unsigned long long produced = 0;
      CK_TRY(cudaMemcpyAsync(&produced, counter, sizeof(produced), cudaMemcpyDeviceToHost, st));
      CK_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(state, st);
      if(kmer) { cudaFreeAsync(kmer, st); }
      state = new_state; kmer = new_kmer; n = produced; stride = capacity;
    }
    if(n > 0)
    {
      ull* counts = nullptr; u64 *lflag = nullptr, *rflag = nullptr, *lpos = nullptr, *rpos = nullptr;
      CK_TRY(engineMallocAsync((void**)&counts, 3 * sizeof(ull), st));
      CK_TRY(cudaMemsetAsync(counts, 0, 3 * sizeof(ull), st));
      if(want)
      {
        CK_ALLOC(lflag, n + 1); CK_ALLOC(rflag, n + 1); CK_ALLOC(lpos, n + 1); CK_ALLOC(rpos, n + 1);
        CK_TRY(cudaMemsetAsync(lflag + n, 0, sizeof(u64), st)); CK_TRY(cudaMemsetAsync(rflag + n, 0, sizeof(u64), st));
      }
      compare_classify_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(state, n, stride, counts, lflag, rflag);
      ull host_counts[3] = { 0, 0, 0 };
      CK_TRY(cudaMemcpyAsync(host_counts, counts, sizeof(host_counts), cudaMemcpyDeviceToHost, st));
      CK_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(counts, st);
      for(int i = 0; i < 3; i++) { result[i] = host_counts[i]; }
      if(want)
      {
        rc = scanExclusive(lflag, lpos, n + 1, st); if(rc) { goto done; }
        rc = scanExclusive(rflag, rpos, n + 1, st); if(rc) { goto done; }
        for(int side = 0; side < 2; side++)
        {
          gcsa_b200_kmer_state** target = (side == 0 ? left_kmers : right_kmers);
          u64 count = result[1 + side];
          if(target == nullptr || count == 0) { continue; }
          u64* records = nullptr;
          CK_ALLOC(records, 8 * count);
          compare_emit_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(state, stride, kmer, n, k, side == 0 ? lflag : rflag, side == 0 ? lpos : rpos, records);
          gcsa_b200_kmer_state* host = (gcsa_b200_kmer_state*)std::malloc(count * sizeof(gcsa_b200_kmer_state));
          if(host == nullptr) { rc = fail(GCSA_B200_ERR_NOMEM, "compare_kmers: out of host memory"); cudaFreeAsync(records, st); goto done; }
          *target = host;
          CK_TRY(cudaMemcpyAsync(host, records, count * sizeof(gcsa_b200_kmer_state), cudaMemcpyDeviceToHost, st));
          CK_TRY(cudaStreamSynchronize(st));
          cudaFreeAsync(records, st);
        }
        cudaFreeAsync(lflag, st); cudaFreeAsync(rflag, st); cudaFreeAsync(lpos, st); cudaFreeAsync(rpos, st);
      }
    }
  }
done:
  if(state) { cudaFreeAsync(state, st); }
  if(kmer) { cudaFreeAsync(kmer, st); }
  if(counter) { cudaFreeAsync(counter, st); }
  #undef CK_TRY
  #undef CK_ALLOC
  if(rc != 0)
  {
    if(left_kmers && *left_kmers) { std::free(*left_kmers); *left_kmers = nullptr; }
    if(right_kmers && *right_kmers) { std::free(*right_kmers); *right_kmers = nullptr; }
  }
  HOST_EPILOGUE("compare_kmers", rc);
}

//------------------------------------------------------------------------------
// MEM-style scan
//------------------------------------------------------------------------------

/*
  One pass over the patterns: every lane counts its matches and writes the first `stride` of them into a
  scratch slot of its pattern; after the scan of the counts a gather kernel moves them into the CSR, and the
  few patterns with more matches are redone writing at their final positions.  (The first version ran the
  whole scan twice, once to count and once to write.)  d_matches_alloc != NULL: the values are allocated
  here (stream-ordered) instead of being written to d_matches.
*/
static int memDevice(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* d_chars, const uint64_t* d_offsets,
                     uint64_t n, uint64_t* d_out_offsets, uint64_t* d_matches, uint64_t capacity, uint64_t* needed, cudaStream_t st,
                     u64** d_matches_alloc)
{
  if(index == nullptr || lcp == nullptr || d_out_offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "mem_batch: null argument"); }
  if(index->device != lcp->device) { return fail(GCSA_B200_ERR_INVALID, "mem_batch: index and LCP array live on different devices"); }
  if(index->header.path_nodes != lcp->view.size) { return fail(GCSA_B200_ERR_INVALID, "mem_batch: index and LCP array have different sizes"); }
  DeviceGuard guard(index->device);
  if(needed) { *needed = 0; }
  if(d_matches_alloc) { *d_matches_alloc = nullptr; }
  CUDA_TRY(cudaMemsetAsync(d_out_offsets, 0, (n + 1) * sizeof(u64), st));
  if(n == 0 || index->header.path_nodes == 0) { return 0; }

  // scratch: up to 16 matches per pattern, fewer for huge batches, none (two full passes) if even 4 do not fit
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  u64 stride = std::min<u64>(16, (free_b / 8) / (n * 32));
  if(const char* e = std::getenv("GCSA_B200_MEM_STRIDE")) { stride = std::min<u64>(stride, (u64)std::atoi(e)); }   // tests: 0 = two passes
  if(stride < 4 && std::getenv("GCSA_B200_MEM_STRIDE") == nullptr) { stride = 0; }

  std::vector<void*> tmp;
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(engineMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { cudaGetLastError(); return nullptr; } tmp.push_back(p); return p; };
  auto cleanup = [&]() { for(void* p : tmp) { cudaFreeAsync(p, st); } tmp.clear(); };
  #define MEM_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { cleanup(); \
    return fail(GCSA_B200_ERR_CUDA, std::string("mem_batch: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)

  u64* counts = (u64*)alloc((n + 1) * sizeof(u64));
  ull* n_overflow = (ull*)alloc(sizeof(ull));
  u64* scratch = (stride > 0 ? (u64*)alloc(n * stride * 32) : nullptr);
  if(counts == nullptr || n_overflow == nullptr) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "mem_batch: out of device memory"); }
  if(scratch == nullptr) { stride = 0; }
  MEM_TRY(cudaMemsetAsync(counts, 0, (n + 1) * sizeof(u64), st));
  MEM_TRY(cudaMemsetAsync(n_overflow, 0, sizeof(ull), st));
  static const int mem_blocks = []() { const char* e_ = std::getenv("GCSA_B200_MEM_MINBLOCKS"); int m_ = (e_ != nullptr ? std::atoi(e_) : 5); return (m_ >= 6 ? 6 : (m_ == 5 ? 5 : 4)); }();   // 5: 23.3 ms against 24.2 (4) and 36.5 (6) per 4 M patterns
  int grid = gridFor(n, index->sm_count, mem_blocks);
  u32 parent_batch = 8;
  if(const char* e = std::getenv("GCSA_B200_MEM_PARENT_BATCH")) { parent_batch = (u32)std::max(1, std::atoi(e)); }
  // GCSA_B200_MEM_PACK=1: the 2-bit packed pattern window instead of byte loads (measured slower: mem.cuh); off by default.
  // GCSA_B200_MEM_MINBLOCKS=6: more resident warps at fewer registers each (experiments).
  bool pack = false;
  if(const char* e = std::getenv("GCSA_B200_MEM_PACK")) { pack = (index->view.default_alphabet != 0) && (std::atoi(e) != 0); }
  #define LAUNCH_MEM(M, G, ...) do { if(jump) { mem_kernel<M, true, false><<<G, 256, 0, st>>>(__VA_ARGS__); } \
    else if(pack) { mem_kernel<M, false, true><<<G, 256, 0, st>>>(__VA_ARGS__); } \
    else if(mem_blocks >= 6) { mem_kernel<M, false, false, 6><<<G, 256, 0, st>>>(__VA_ARGS__); } \
    else if(mem_blocks == 5) { mem_kernel<M, false, false, 5><<<G, 256, 0, st>>>(__VA_ARGS__); } \
    else { mem_kernel<M, false, false><<<G, 256, 0, st>>>(__VA_ARGS__); } } while(0)
  // GCSA_B200_MEM_JUMP=1: singleton ranges follow the jump tables (mem_kernel<.., JUMP>); off by default: measured slower
  bool jump = false;
  if(const char* e = std::getenv("GCSA_B200_MEM_JUMP")) { jump = (std::atoi(e) != 0 && index->view.jump != nullptr && index->view.default_alphabet != 0); }
  if(stride > 0)
  {
    LAUNCH_MEM(2, grid, index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, scratch, nullptr, stride, parent_batch);
  }
  else
  {
    LAUNCH_MEM(0, grid, index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, nullptr, nullptr, 0, parent_batch);
  }
  int rc = scanExclusive(counts, (u64*)d_out_offsets, n + 1, st);
  if(rc) { cleanup(); return rc; }
  if(stride > 0) { mem_count_overflow_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>(counts, n, stride, n_overflow); }
  u64 total = 0; ull overflowing = 0;
  MEM_TRY(cudaMemcpyAsync(&total, (u64*)d_out_offsets + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
  MEM_TRY(cudaMemcpyAsync(&overflowing, n_overflow, sizeof(ull), cudaMemcpyDeviceToHost, st));
  MEM_TRY(cudaStreamSynchronize(st));
  if(needed) { *needed = total; }
  if(d_matches_alloc != nullptr)
  {
    void* p = nullptr;
    MEM_TRY(engineMallocAsync(&p, std::max<u64>(total, 1) * 32, st));
    *d_matches_alloc = (u64*)p; d_matches = (u64*)p; capacity = total;
  }
  if(d_matches == nullptr || capacity < total) { cleanup(); return fail(GCSA_B200_ERR_CAPACITY, "mem_batch: output capacity too small"); }
  if(stride == 0)
  {
    LAUNCH_MEM(1, grid, index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, nullptr, (const u64*)d_out_offsets, (u64*)d_matches, nullptr, 0, parent_batch);
  }
  else
  {
    u64* overflow = (u64*)alloc(std::max<u64>(overflowing, 1) * sizeof(u64));
    if(overflow == nullptr) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "mem_batch: out of device memory"); }
    MEM_TRY(cudaMemsetAsync(n_overflow, 0, sizeof(ull), st));
    mem_gather_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>((const ulonglong4*)scratch, counts, (const u64*)d_out_offsets, n, stride,
                                                                   (ulonglong4*)d_matches, overflow, n_overflow);
    if(overflowing > 0)
    {
      LAUNCH_MEM(1, gridFor(overflowing, index->sm_count, 4), index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, overflowing,
                 nullptr, (const u64*)d_out_offsets, (u64*)d_matches, overflow, 0, parent_batch);
    }
  }
  MEM_TRY(cudaGetLastError());
  cleanup();
  #undef LAUNCH_MEM
  #undef MEM_TRY
  return 0;
}

int gcsa_b200_mem_batch(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* d_chars, const uint64_t* d_offsets,
                        uint64_t n, uint64_t* d_out_offsets, uint64_t* d_matches, uint64_t capacity, uint64_t* needed, void* stream)
{
  return memDevice(index, lcp, d_chars, d_offsets, n, d_out_offsets, d_matches, capacity, needed, (cudaStream_t)stream, nullptr);
}

int gcsa_b200_mem_host(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* chars, const uint64_t* offsets,
                       uint64_t n, uint64_t* out_offsets, uint64_t** matches)
{
  if(out_offsets == nullptr || matches == nullptr || (n > 0 && (chars == nullptr || offsets == nullptr))) { return fail(GCSA_B200_ERR_INVALID, "mem_host: null argument"); }
  *matches = nullptr;
  HOST_PROLOGUE("mem_host", index);
  u64 total_chars = (n ? offsets[n] : 0);
  u8* d_chars = sc.in(chars, total_chars + 1 > 1 ? total_chars : 1);
  u64* d_off = sc.in((const u64*)offsets, n + 1);
  u64* d_out = sc.alloc<u64>(n + 1);
  u64 needed = 0;
  u64* d_vals = nullptr;
  int rc = memDevice(index, lcp, d_chars, d_off, n, d_out, nullptr, 0, &needed, sc.stream, &d_vals);
  if(rc == 0)
  {
    u64* vals = (u64*)std::malloc(std::max<u64>(4 * needed, 1) * sizeof(u64));
    if(d_vals != nullptr) { sc.out(vals, d_vals, 4 * needed); sc.ptrs.push_back(d_vals); }
    sc.out((u64*)out_offsets, d_out, n + 1);
    *matches = (uint64_t*)vals;
  }
  HOST_EPILOGUE("mem_host", rc);
}
