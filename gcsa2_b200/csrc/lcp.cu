/*
  lcp.cu -- LCPArray: creation, parent / depth / psv / nsv / rmq entry points, and the MEM-style scan (LF + parent).
  One of the CUDA translation units of libgcsa2_b200.so (see engine.h); host side of the C ABI of include/gcsa2_b200.h,
  kernels in the device/*.cuh it includes.
*/
#include "engine.h"
#include "device/lcp.cuh"
#include "device/mem.cuh"
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

  // Scratch for the one-pass form: pattern q owns (len >> shift) + 4 entries of 32 bytes (device/mem.cuh), with the
  // smallest shift in 0 .. 2 that a quarter of the free memory holds.  A pattern reports at most one match per
  // character it consumes, so with shift = 0 (32 bytes per pattern character) no slot can overflow; with 1 or 2 the
  // overflowing patterns are rare and redone.  Else 16 entries each, fewer for huge batches, and none (two full passes)
  // if even 4 do not fit.  GCSA_B200_MEM_STRIDE (tests): that many entries each, 0 = two passes; GCSA_B200_MEM_SHIFT
  // (tests, experiments): that shift.
  size_t free_b = engineFreeMemory();
  u64 total_chars = 0;
  CUDA_TRY(cudaMemcpyAsync(&total_chars, d_offsets + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  u64 stride = 4; u32 shift = 0;
  if(const char* e = std::getenv("GCSA_B200_MEM_SHIFT")) { shift = (u32)std::min(std::max(std::atoi(e), 0), 2); }
  auto entriesFor = [&](u32 sh) -> u64 { return (total_chars >> sh) + stride * n + 1; };
  while(shift < 2 && entriesFor(shift) * 32 > free_b / 4) { shift++; }
  u64 scratch_entries = entriesFor(shift);
  const char* stride_env = std::getenv("GCSA_B200_MEM_STRIDE");
  if(stride_env != nullptr || scratch_entries * 32 > free_b / 4)
  {
    shift = 63;
    stride = std::min<u64>(16, (free_b / 8) / (n * 32));
    if(stride_env != nullptr) { stride = std::min<u64>(stride, (u64)std::atoi(stride_env)); }
    else if(stride < 4) { stride = 0; }
    scratch_entries = n * stride;
  }

  std::vector<void*> tmp;
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(engineMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { cudaGetLastError(); return nullptr; } tmp.push_back(p); return p; };
  auto cleanup = [&]() { for(void* p : tmp) { cudaFreeAsync(p, st); } tmp.clear(); };
  #define MEM_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { cleanup(); \
    return fail(GCSA_B200_ERR_CUDA, std::string("mem_batch: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)

  u64* counts = (u64*)alloc((n + 1) * sizeof(u64));
  ull* n_overflow = (ull*)alloc(sizeof(ull));
  u64* scratch = (stride > 0 ? (u64*)alloc(scratch_entries * 32) : nullptr);
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
    LAUNCH_MEM(2, grid, index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, scratch, nullptr, stride, parent_batch, shift);
  }
  else
  {
    LAUNCH_MEM(0, grid, index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, nullptr, nullptr, 0, parent_batch, shift);
  }
  int rc = scanExclusive(counts, (u64*)d_out_offsets, n + 1, st);
  if(rc) { cleanup(); return rc; }
  if(stride > 0 && shift != 0) { mem_count_overflow_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>(counts, (const u64*)d_offsets, n, stride, shift, n_overflow); }
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
    LAUNCH_MEM(1, grid, index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, nullptr, (const u64*)d_out_offsets, (u64*)d_matches, nullptr, 0, parent_batch, shift);
  }
  else
  {
    u64* overflow = (u64*)alloc(std::max<u64>(overflowing, 1) * sizeof(u64));
    if(overflow == nullptr) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "mem_batch: out of device memory"); }
    MEM_TRY(cudaMemsetAsync(n_overflow, 0, sizeof(ull), st));
    mem_gather_kernel<<<gridFor(8 * n, index->sm_count), 256, 0, st>>>((const ulonglong4*)scratch, counts, (const u64*)d_out_offsets, (const u64*)d_offsets, 0, n, stride, shift,
                                                                   (ulonglong4*)d_matches, overflow, n_overflow);
    if(overflowing > 0)
    {
      LAUNCH_MEM(1, gridFor(overflowing, index->sm_count, 4), index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, overflowing,
                 nullptr, (const u64*)d_out_offsets, (u64*)d_matches, overflow, 0, parent_batch, shift);
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

