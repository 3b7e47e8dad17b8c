/*
  cuda_emu.h -- TEST INFRASTRUCTURE.  A minimal host emulation of the CUDA constructs that
  gcsa2_b200/csrc/engine.cu uses, so that the kernel SOURCE (not a restatement of it) can be executed
  and checked against the oracle on a machine without a GPU (tests/test_emu_kernels.py).

  It is never part of the product: the shipped library is built by nvcc for sm_100a only
  (gcsa2_b200/build.py) and has no CPU path; this header is only ever seen by tests/emu/build_emu.py,
  which translates engine.cu (kernel launches `k<<<cfg>>>(args)` -> emu::launch, the one inline-PTX load)
  and compiles the result with g++ into tests/emu/_build/.

  Execution model: one launch runs on the calling OS thread, block after block.  Kernels without warp /
  block collectives run thread after thread as plain calls.  Kernels with collectives run every thread of a
  block as a fiber (own stack, hand-written context switch); __ballot_sync / __shfl_*_sync / __syncthreads
  suspend the fiber until all live participants have arrived, which is the lockstep the kernels rely on.
  Memory "on the device" is host memory; streams and events are ordered trivially (everything is
  synchronous).
*/
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <sys/mman.h>

//------------------------------------------------------------------------------
// Vector types, qualifiers
//------------------------------------------------------------------------------

struct uint3 { unsigned x, y, z; };
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
  dim3(int x_) : x((unsigned)x_), y(1), z(1) {}
  dim3(unsigned long x_) : x((unsigned)x_), y(1), z(1) {}
};
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(16) ulonglong4 { unsigned long long x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }
static inline ulonglong4 make_ulonglong4(unsigned long long x, unsigned long long y, unsigned long long z, unsigned long long w) { ulonglong4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static

//------------------------------------------------------------------------------
// Runtime API (the subset engine.cu uses)
//------------------------------------------------------------------------------

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
struct emu_stream_; typedef emu_stream_* cudaStream_t;
struct emu_event_;  typedef emu_event_* cudaEvent_t;
typedef void* cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxPersistingL2CacheSize = 108, cudaDevAttrMaxAccessPolicyWindowSize = 109 };
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5, cudaLimitPersistingL2CacheSize = 6 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4, cudaMemPoolAttrReservedMemCurrent = 5, cudaMemPoolAttrUsedMemCurrent = 7 };

namespace emu
{
inline size_t envBytes(const char* name, size_t def) { const char* s = std::getenv(name); return (s != nullptr && *s ? (size_t)std::strtoull(s, nullptr, 10) : def); }
// "Device" allocations end right before an inaccessible page (sizes rounded up to 32 bytes, the sector the
// kernels read): a kernel that reads or writes past the end of a buffer faults here instead of passing silently.
struct Allocations
{
  std::mutex mutex;
  std::unordered_map<void*, std::pair<void*, size_t>> live;     // pointer -> (mapping, length)
  std::map<uintptr_t, size_t> extent;                            // pointer -> usable bytes (range lookups)
};
inline Allocations& allocations() { static Allocations a; return a; }
// true if p points into (or one past the end of) memory obtained from cudaMalloc* / cudaHostAlloc
inline bool isDeviceVisible(const void* p);
inline void* alloc(size_t bytes)
{
  const size_t page = 4096;
  size_t total = (std::max<size_t>(bytes, 32) + 31) & ~(size_t)31;
  size_t pages = (total + page - 1) / page + 1;
  char* base = (char*)mmap(nullptr, pages * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if(base == MAP_FAILED) { return nullptr; }
  mprotect(base + (pages - 1) * page, page, PROT_NONE);
  void* p = base + (pages - 1) * page - total;
  // cudaMalloc does not clear memory: poison it, so that a kernel relying on zeroes it never wrote shows up here
  static const bool poison = (envBytes("GCSA_EMU_NO_POISON", 0) == 0);
  if(poison) { std::memset(p, 0xA5, total); }
  Allocations& a = allocations();
  std::lock_guard<std::mutex> lock(a.mutex);
  a.live[p] = std::make_pair((void*)base, pages * page);
  a.extent[(uintptr_t)p] = total;
  return p;
}
inline void release(void* p)
{
  if(p == nullptr) { return; }
  Allocations& a = allocations();
  std::lock_guard<std::mutex> lock(a.mutex);
  auto it = a.live.find(p);
  if(it == a.live.end()) { std::fprintf(stderr, "cuda_emu: free of a pointer that was not allocated here\n"); std::abort(); }
  munmap(it->second.first, it->second.second);
  a.live.erase(it);
  a.extent.erase((uintptr_t)p);
}
}

namespace emu
{
inline bool isDeviceVisible(const void* p)
{
  Allocations& a = allocations();
  std::lock_guard<std::mutex> lock(a.mutex);
  auto it = a.extent.upper_bound((uintptr_t)p);
  if(it == a.extent.begin()) { return false; }
  --it;
  return (uintptr_t)p <= it->first + it->second;
}
}

// Memory that did not come from cudaMalloc here but stands for device memory (the tests' torch tensors): the test
// registers it so that kernels may be handed pointers into it.
#ifdef EMU_DEFINE_SWITCH
extern "C" void emu_register_device_range(const void* p, size_t bytes)
{
  emu::Allocations& a = emu::allocations();
  std::lock_guard<std::mutex> lock(a.mutex);
  a.extent[(uintptr_t)p] = bytes;
}
extern "C" void emu_unregister_device_range(const void* p)
{
  emu::Allocations& a = emu::allocations();
  std::lock_guard<std::mutex> lock(a.mutex);
  a.extent.erase((uintptr_t)p);
}
#endif

static inline const char* cudaGetErrorString(cudaError_t e) { return (e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory (emulated)" : "error (emulated)")); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return (d == 0 ? cudaSuccess : cudaErrorInvalidValue); }
static inline cudaError_t cudaDeviceGetAttribute(int* value, cudaDeviceAttr attr, int) { *value = (attr == cudaDevAttrMultiProcessorCount ? (int)emu::envBytes("GCSA_EMU_SMS", 2) : 0); return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
static inline cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t, cudaMemPoolAttr, void* value) { *(uint64_t*)value = 0; return cudaSuccess; }
static inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
enum { cudaMemAllocationTypePinned = 1, cudaMemHandleTypeNone = 0, cudaMemLocationTypeDevice = 1 };
struct cudaMemPoolProps { int allocType, handleTypes; struct { int type, id; } location; };
static inline cudaError_t cudaMemPoolCreate(cudaMemPool_t* pool, const cudaMemPoolProps*) { static int token; *pool = &token; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = *total_b = emu::envBytes("GCSA_EMU_FREE_BYTES", (size_t)8 << 30); return cudaSuccess; }
template<class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { *p = (T*)emu::alloc(bytes); return (*p != nullptr ? cudaSuccess : cudaErrorMemoryAllocation); }
template<class T> static inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
template<class T> static inline cudaError_t cudaMallocFromPoolAsync(T** p, size_t bytes, cudaMemPool_t, cudaStream_t) { return cudaMalloc(p, bytes); }
template<class T> static inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFree(void* p) { emu::release(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { emu::release(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { emu::release(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
  if(bytes == 0) { return cudaSuccess; }
  // the device side of a copy must be device memory (the runtime returns cudaErrorInvalidValue otherwise)
  bool dst_ok = (kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice ? emu::isDeviceVisible(dst) && emu::isDeviceVisible((const char*)dst + bytes) : true);
  bool src_ok = (kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice ? emu::isDeviceVisible(src) && emu::isDeviceVisible((const char*)src + bytes) : true);
  if(!dst_ok || !src_ok) { std::fprintf(stderr, "cuda_emu: cudaMemcpy kind %d with a %s that is not device memory (or runs past its end)\n", (int)kind, dst_ok ? "source" : "destination"); std::abort(); }
  std::memmove(dst, src, bytes);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(dst, src, bytes, k); }
static inline cudaError_t cudaMemset(void* p, int value, size_t bytes) { if(bytes) { std::memset(p, value, bytes); } return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t = nullptr) { return cudaMemset(p, value, bytes); }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)emu::alloc(16); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { emu::release((void*)s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)emu::alloc(16); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { emu::release((void*)e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
// (two blocks per "SM": small grids, so that the grid-stride loops of the kernels sized this way are exercised)
template<class Kernel>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, Kernel, int, size_t) { *blocks = 2; return cudaSuccess; }

//------------------------------------------------------------------------------
// Thread coordinates and the fiber scheduler
//------------------------------------------------------------------------------

namespace emu
{

struct Coords { uint3 tid, bid; dim3 bdim, gdim; };
inline thread_local Coords coords;

#define threadIdx (emu::coords.tid)
#define blockIdx  (emu::coords.bid)
#define blockDim  (emu::coords.bdim)
#define gridDim   (emu::coords.gdim)

[[noreturn]] inline void die(const char* what) { std::fprintf(stderr, "cuda_emu: %s\n", what); std::abort(); }

// Context switch: callee-saved registers + stack pointer (System V x86-64).
extern "C" void emu_switch(void** save_sp, void* load_sp);
#ifdef EMU_DEFINE_SWITCH
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
#endif

// GCSA_EMU_SHUFFLE=seed: blocks of a launch, threads of a block (and the order in which waiting fibers are resumed)
// run in a pseudo-random order instead of ascending, so that results which depend on the order in which threads
// reach an atomic or a buffer -- arbitrary on the device -- show up as differences.
inline unsigned long long& shuffleState()
{
  static thread_local unsigned long long state = 0;
  return state;
}
inline bool shuffling()
{
  static const unsigned long long seed = envBytes("GCSA_EMU_SHUFFLE", 0);
  if(seed != 0 && shuffleState() == 0) { shuffleState() = seed * 0x9E3779B97F4A7C15ull + 1; }
  return seed != 0;
}
inline unsigned long long nextRandom()
{
  unsigned long long& s = shuffleState();
  s ^= s << 13; s ^= s >> 7; s ^= s << 17;
  return s;
}
// i-th element of a pseudo-random permutation of [0, n): an affine map with a multiplier coprime to n
struct Permutation
{
  unsigned long long n, mul, add;
  explicit Permutation(unsigned long long n_) : n(n_), mul(1), add(0)
  {
    if(!shuffling() || n < 2) { return; }
    add = nextRandom() % n;
    do { mul = nextRandom() % n; } while(mul == 0 || std::__gcd(mul, n) != 1);
  }
  unsigned operator()(unsigned long long i) const { return (unsigned)((i * mul + add) % n); }
};

struct Warp
{
  unsigned gen = 0, arrived = 0, alive = 0;
  int site = 0;                        // source line of the collective the current generation is gathering at
  unsigned pred_bits[2] = {0, 0};
  unsigned long long vals[2][32];
};

struct Block;
struct Fiber
{
  void* sp = nullptr; char* stack = nullptr;
  bool finished = true;
  unsigned tid = 0;
};

struct Block
{
  static constexpr size_t STACK = 256 << 10;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  void* sched_sp = nullptr;
  Fiber* current = nullptr;
  unsigned bar_gen = 0, bar_arrived = 0, alive = 0;
  unsigned long progress = 0;
  const void* body = nullptr; void (*invoke)(const void*) = nullptr;

  ~Block() { for(Fiber& f : fibers) { if(f.stack) { munmap(f.stack, STACK); } } }
  void yield() { Fiber* f = current; emu_switch(&f->sp, sched_sp); }
};
inline thread_local Block* block = nullptr;       // non-null while a fiber-mode kernel is running on this OS thread

inline void fiberMain()
{
  Block* b = block; Fiber* f = b->current;
  b->invoke(b->body);
  f->finished = true;
  Warp& w = b->warps[f->tid >> 5];
  b->alive--; w.alive--; b->progress++;
  // an exited thread no longer takes part: release collectives that were only waiting for it
  if(w.arrived > 0 && w.arrived >= w.alive) { w.arrived = 0; w.gen++; }
  if(b->bar_arrived > 0 && b->bar_arrived >= b->alive) { b->bar_arrived = 0; b->bar_gen++; }
  void* dummy; emu_switch(&dummy, b->sched_sp);
  die("finished fiber resumed");
}
extern "C" inline void emu_fiber_entry() { fiberMain(); }

template<class F> void runBlockFibers(Block& b, unsigned threads, const F& f)
{
  if(b.fibers.size() < threads) { b.fibers.resize(threads); }
  b.warps.assign((threads + 31) / 32, Warp());
  b.body = &f; b.invoke = [](const void* p) { (*(const F*)p)(); };
  b.bar_gen = 0; b.bar_arrived = 0; b.alive = threads;
  for(unsigned t = 0; t < threads; t++)
  {
    Fiber& fb = b.fibers[t];
    if(fb.stack == nullptr)
    {
      fb.stack = (char*)mmap(nullptr, Block::STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if(fb.stack == MAP_FAILED) { die("mmap of a fiber stack failed"); }
    }
    fb.finished = false; fb.tid = t;
    b.warps[t >> 5].alive++;
    // initial frame: six zeroed callee-saved registers, then the entry point as the return address;
    // after `ret` the stack pointer is 8 mod 16, as at any function entry
    uintptr_t top = ((uintptr_t)(fb.stack + Block::STACK) & ~(uintptr_t)15) - 8;
    void** sp = (void**)top;
    *(--sp) = (void*)&emu_fiber_entry;
    for(int i = 0; i < 6; i++) { *(--sp) = nullptr; }
    fb.sp = sp;
  }
  block = &b;
  while(b.alive > 0)
  {
    unsigned long before = b.progress;
    const Permutation order(threads);
    for(unsigned i = 0; i < threads; i++)
    {
      const unsigned t = order(i);
      Fiber& fb = b.fibers[t];
      if(fb.finished) { continue; }
      b.current = &fb; coords.tid.x = t;
      emu_switch(&b.sched_sp, fb.sp);
    }
    if(b.progress == before) { die("deadlock: every thread of the block waits at a collective that cannot complete"); }
  }
  block = nullptr;
}

struct Cfg
{
  dim3 grid, blockdim;
  Cfg(dim3 g, dim3 b, size_t = 0, cudaStream_t = nullptr) : grid(g), blockdim(b) {}
};

template<class F> void launch(const Cfg& cfg, const F& f, bool collectives)
{
  if(cfg.grid.y != 1 || cfg.grid.z != 1 || cfg.blockdim.y != 1 || cfg.blockdim.z != 1) { die("only 1-D launches are emulated"); }
  if(block != nullptr) { die("nested launch"); }
  Coords saved = coords;
  coords.gdim = cfg.grid; coords.bdim = cfg.blockdim;
  coords.tid = uint3{0, 0, 0}; coords.bid = uint3{0, 0, 0};
  static thread_local Block blk;
  const Permutation block_order(cfg.grid.x);
  for(unsigned i = 0; i < cfg.grid.x; i++)
  {
    coords.bid.x = block_order(i);
    if(collectives) { runBlockFibers(blk, cfg.blockdim.x, f); }
    else
    {
      const Permutation thread_order(cfg.blockdim.x);
      for(unsigned t = 0; t < cfg.blockdim.x; t++) { coords.tid.x = thread_order(t); f(); }
    }
  }
  coords = saved;
}

// A kernel launch with its arguments: every pointer argument must be null or point into memory the device can see
// (cudaMalloc*, cudaHostAlloc) -- a pointer to ordinary host memory works here but faults on the device.
template<class T> inline void checkKernelArgument(const T&, int) {}
template<class T> struct AlignOf { static constexpr size_t value = alignof(T); };
template<> struct AlignOf<void> { static constexpr size_t value = 1; };
template<> struct AlignOf<const void> { static constexpr size_t value = 1; };
template<class T> inline void checkKernelArgument(T* const& p, int position)
{
  if(p != nullptr && !isDeviceVisible((const void*)p))
  {
    std::fprintf(stderr, "cuda_emu: kernel argument %d (%p) is not a device pointer\n", position, (const void*)p);
    std::abort();
  }
  // the device faults on an access that is not aligned to its size: a pointer to 16-byte elements must be 16-byte aligned
  if(((uintptr_t)p % AlignOf<T>::value) != 0)
  {
    std::fprintf(stderr, "cuda_emu: kernel argument %d (%p) is not aligned to its element type (%zu bytes)\n", position, (const void*)p, AlignOf<T>::value);
    std::abort();
  }
}
template<class K, class... A> void launchChecked(const Cfg& cfg, bool collectives, K&& kernel, A&&... args)
{
  int position = 0;
  (checkKernelArgument(args, position++), ...);
  launch(cfg, [&] { kernel(args...); }, collectives);
}

// One warp-wide exchange: every live lane of the warp deposits (pred, value) and gets the generation's
// buffers back once all live lanes have arrived.
struct Exchange { unsigned ballot; const unsigned long long* vals; };
inline Exchange warpExchange(unsigned mask, bool pred, unsigned long long value, int site)
{
  Block* b = block;
  if(b == nullptr) { die("warp collective in a kernel that the translator classified as collective-free"); }
  unsigned tid = b->current->tid, lane = tid & 31;
  Warp& w = b->warps[tid >> 5];
  if(!((mask >> lane) & 1)) { die("calling lane is not in the mask of a *_sync collective"); }
  unsigned g = w.gen, slot = g & 1;
  // every lane of a generation must be at the same collective: lanes that meet at different call sites are
  // undefined behaviour on the device (and a wrong answer here), whatever their masks say
  if(w.arrived == 0) { w.pred_bits[slot] = 0; w.site = site; }
  else if(w.site != site)
  {
    std::fprintf(stderr, "cuda_emu: lanes of one warp wait at different collectives (lines %d and %d)\n", w.site, site);
    std::abort();
  }
  if(pred) { w.pred_bits[slot] |= (1u << lane); }
  w.vals[slot][lane] = value;
  w.arrived++; b->progress++;
  if(w.arrived >= w.alive) { w.arrived = 0; w.gen++; }
  else { while(w.gen == g) { b->yield(); } }
  return Exchange{ w.pred_bits[slot] & mask, w.vals[slot] };
}

inline void blockBarrier()
{
  Block* b = block;
  if(b == nullptr) { return; }      // direct mode: threads run one after another; a kernel that needs the barrier is run on fibers
  unsigned g = b->bar_gen;
  b->bar_arrived++; b->progress++;
  if(b->bar_arrived >= b->alive) { b->bar_arrived = 0; b->bar_gen++; }
  else { while(b->bar_gen == g) { b->yield(); } }
}

} // namespace emu

static inline unsigned emu_ballot(unsigned mask, int pred, int site) { return emu::warpExchange(mask, pred != 0, 0, site).ballot; }
static inline int emu_any(unsigned mask, int pred, int site) { return emu::warpExchange(mask, pred != 0, 0, site).ballot != 0; }
static inline int emu_all(unsigned mask, int pred, int site) { return emu::warpExchange(mask, pred == 0, 0, site).ballot == 0; }
static inline void emu_syncwarp(unsigned mask, int site) { emu::warpExchange(mask, false, 0, site); }
static inline void __syncthreads() { emu::blockBarrier(); }
// kind 0: absolute source lane, 1: down by delta, 2: up by delta, 3: xor
template<class T> static inline T emu_shfl(unsigned mask, T v, int arg, int width, int kind, int site)
{
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  unsigned long long raw = 0; std::memcpy(&raw, &v, sizeof(T));
  unsigned lane = emu::block->current->tid & 31, w = (unsigned)width, base = lane & ~(w - 1), rel = lane & (w - 1), from = lane;
  emu::Exchange x = emu::warpExchange(mask, false, raw, site);
  if(kind == 0) { from = base | ((unsigned)arg & (w - 1)); }
  else if(kind == 1) { from = (rel + (unsigned)arg < w ? lane + (unsigned)arg : lane); }
  else if(kind == 2) { from = (rel >= (unsigned)arg ? lane - (unsigned)arg : lane); }
  else { from = lane ^ (unsigned)arg; }
  T r; std::memcpy(&r, &x.vals[from & 31], sizeof(T)); return r;
}
#define __ballot_sync(mask, pred) emu_ballot((mask), (pred), __LINE__)
#define __any_sync(mask, pred) emu_any((mask), (pred), __LINE__)
#define __all_sync(mask, pred) emu_all((mask), (pred), __LINE__)
#define __syncwarp(...) emu_syncwarp(emu_first_or_full(__VA_ARGS__), __LINE__)
static inline unsigned emu_first_or_full(unsigned mask = 0xFFFFFFFFu) { return mask; }
#define EMU_SHFL_WIDTH(a, b, c, w, ...) w
#define __shfl_sync(...) emu_shfl(EMU_SHFL3(__VA_ARGS__), EMU_SHFL_WIDTH(__VA_ARGS__, 32, 32), 0, __LINE__)
#define __shfl_down_sync(...) emu_shfl(EMU_SHFL3(__VA_ARGS__), EMU_SHFL_WIDTH(__VA_ARGS__, 32, 32), 1, __LINE__)
#define __shfl_up_sync(...) emu_shfl(EMU_SHFL3(__VA_ARGS__), EMU_SHFL_WIDTH(__VA_ARGS__, 32, 32), 2, __LINE__)
#define __shfl_xor_sync(...) emu_shfl(EMU_SHFL3(__VA_ARGS__), EMU_SHFL_WIDTH(__VA_ARGS__, 32, 32), 3, __LINE__)
#define EMU_SHFL3(...) EMU_SHFL3_(__VA_ARGS__, 0)
#define EMU_SHFL3_(a, b, c, ...) (a), (b), (int)(c)

//------------------------------------------------------------------------------
// Device intrinsics
//------------------------------------------------------------------------------

// The device faults on an access that is not aligned to its size ("misaligned address"); x86 does not.
namespace emu
{
inline void checkAligned(const void* p, size_t size)
{
  if(((uintptr_t)p & (size - 1)) != 0) { std::fprintf(stderr, "cuda_emu: %zu-byte access at %p is misaligned\n", size, p); std::abort(); }
}
}
template<class T> static inline T __ldg(const T* p) { emu::checkAligned(p, sizeof(T)); return *p; }
template<class T> static inline T __ldcs(const T* p) { emu::checkAligned(p, sizeof(T)); return *p; }
template<class T> static inline T __ldcg(const T* p) { emu::checkAligned(p, sizeof(T)); return *p; }
template<class T> static inline void __stcs(T* p, T v) { emu::checkAligned(p, sizeof(T)); *p = v; }
template<class T> static inline void __stcg(T* p, T v) { emu::checkAligned(p, sizeof(T)); *p = v; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return (x == 0 ? 32 : __builtin_clz((unsigned)x)); }
static inline int __clzll(long long x) { return (x == 0 ? 64 : __builtin_clzll((unsigned long long)x)); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x)
{
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(x);
}
static inline unsigned long long __brevll(unsigned long long x) { return ((unsigned long long)__brev((unsigned)x) << 32) | __brev((unsigned)(x >> 32)); }
// position of the offset-th (1-based) set bit of mask at or above bit `base`; 0xFFFFFFFF if there is none
static inline unsigned __fns(unsigned mask, unsigned base, int offset)
{
  if(offset <= 0) { emu::die("__fns with offset <= 0 is not emulated"); }
  for(unsigned i = base; i < 32; i++) { if((mask >> i) & 1) { if(--offset == 0) { return i; } } }
  return 0xFFFFFFFFu;
}
static inline unsigned __vcmpltu4(unsigned a, unsigned b)
{
  unsigned r = 0;
  for(int i = 0; i < 4; i++) { if(((a >> (8 * i)) & 0xFF) < ((b >> (8 * i)) & 0xFF)) { r |= 0xFFu << (8 * i); } }
  return r;
}
static inline unsigned __vcmpleu4(unsigned a, unsigned b)
{
  unsigned r = 0;
  for(int i = 0; i < 4; i++) { if(((a >> (8 * i)) & 0xFF) <= ((b >> (8 * i)) & 0xFF)) { r |= 0xFFu << (8 * i); } }
  return r;
}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template<class T, class U> static inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
template<class T, class U> static inline T atomicOr(T* p, U v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_RELAXED); }
template<class T, class U> static inline T atomicMax(T* p, U v) { T old = *p; while(old < (T)v && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return old; }
template<class T, class U> static inline T atomicMin(T* p, U v) { T old = *p; while(old > (T)v && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return old; }
template<class T, class U> static inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_RELAXED); }
template<class T, class U> static inline T atomicCAS(T* p, U cmp, U v) { T expected = (T)cmp; __atomic_compare_exchange_n(p, &expected, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return expected; }

//------------------------------------------------------------------------------
// cub (the two device-wide primitives engine.cu calls)
//------------------------------------------------------------------------------

namespace cub
{
struct DeviceScan
{
  template<class In, class Out, class N>
  static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, In in, Out out, N n, cudaStream_t = nullptr)
  {
    if(tmp == nullptr) { bytes = 256; return cudaSuccess; }
    typedef typename std::remove_cv<typename std::remove_reference<decltype(out[0])>::type>::type T;
    T sum = 0;
    for(N i = 0; i < n; i++) { T x = (T)in[i]; out[i] = sum; sum += x; }
    return cudaSuccess;
  }
  template<class In, class Out, class N>
  static cudaError_t InclusiveSum(void* tmp, size_t& bytes, In in, Out out, N n, cudaStream_t = nullptr)
  {
    if(tmp == nullptr) { bytes = 256; return cudaSuccess; }
    typedef typename std::remove_cv<typename std::remove_reference<decltype(out[0])>::type>::type T;
    T sum = 0;
    for(N i = 0; i < n; i++) { sum += (T)in[i]; out[i] = sum; }
    return cudaSuccess;
  }
};
struct DeviceRadixSort
{
  // stable, by bits [begin_bit, end_bit) of the key
  template<class K, class V, class N>
  static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, N n,
                               int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr)
  {
    if(tmp == nullptr) { bytes = 256; return cudaSuccess; }
    if(keys_in == keys_out || (vals_in != nullptr && (const void*)vals_in == (const void*)vals_out)) { emu::die("DeviceRadixSort: in-place sorting is not allowed"); }
    const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8 ? ~(K)0 : (((K)1 << (end_bit - begin_bit)) - 1));
    std::vector<size_t> order((size_t)n);
    for(size_t i = 0; i < (size_t)n; i++) { order[i] = i; }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ((keys_in[a] >> begin_bit) & mask) < ((keys_in[b] >> begin_bit) & mask); });
    for(size_t i = 0; i < (size_t)n; i++) { keys_out[i] = keys_in[order[i]]; if(vals_in != nullptr) { vals_out[i] = vals_in[order[i]]; } }
    return cudaSuccess;
  }
  template<class K, class N>
  static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* keys_in, K* keys_out, N n,
                              int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t s = nullptr)
  {
    return SortPairs(tmp, bytes, keys_in, keys_out, (const char*)nullptr, (char*)nullptr, n, begin_bit, end_bit, s);
  }
};
struct DeviceSegmentedSort
{
  template<class K, class B, class E>
  static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* in, K* out, long long total, long long segments, B begins, E ends, cudaStream_t = nullptr)
  {
    if(tmp == nullptr) { bytes = 256; return cudaSuccess; }
    if(total > 0 && out != in) { std::memcpy(out, in, (size_t)total * sizeof(K)); }
    for(long long s = 0; s < segments; s++) { std::sort(out + begins[s], out + ends[s]); }
    return cudaSuccess;
  }
};
}
