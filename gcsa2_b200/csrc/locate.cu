/*
  locate.cu -- locate(): the general pipeline, the short-range path, the host-buffer pipelines, the multi-GPU form.
  One of the CUDA translation units of libgcsa2_b200.so (see engine.h); host side of the C ABI of include/gcsa2_b200.h,
  kernels in the device/*.cuh it includes.
*/
#include "engine.h"
#include "device/locate.cuh"
//------------------------------------------------------------------------------
// locate
//------------------------------------------------------------------------------

namespace {


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

// blocks of each locate_medium_kernel class that one SM holds at once
struct MediumGrids { int per_sm[4]; };
MediumGrids mediumGrids()
{
  MediumGrids g = { { 8, 3, 2, 1 } };
  int b = 0;
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, locate_medium_kernel<0>, 128, 0) == cudaSuccess && b > 0) { g.per_sm[0] = b; }
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, locate_medium_kernel<1>, 128, 0) == cudaSuccess && b > 0) { g.per_sm[1] = b; }
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, locate_medium_kernel<2>, 256, 0) == cudaSuccess && b > 0) { g.per_sm[2] = b; }
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, locate_medium_kernel<3>, 512, 0) == cudaSuccess && b > 0) { g.per_sm[3] = b; }
  return g;
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

  const char* med_env = std::getenv("GCSA_B200_LOCATE_MEDIUM");
  const bool medium = (med_env == nullptr || std::atoi(med_env) != 0);
  u64* cnt = (u64*)alloc((n + 1) * sizeof(u64));
  u64* stash = (u64*)alloc(n * sizeof(u64));
  u64* glist = (u64*)alloc(n * sizeof(u64));
  u64* mlist = (u64*)alloc(2 * n * sizeof(u64));
  ull* d_counters = (ull*)alloc(6 * sizeof(ull));
  if(!cnt || !stash || !glist || !mlist || !d_counters) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
  LOC_TRY(cudaMemsetAsync(cnt + n, 0, sizeof(u64), st));
  LOC_TRY(cudaMemsetAsync(d_counters, 0, 6 * sizeof(ull), st));
  locate_small_count_kernel<<<gridFor(n, sm), 256, 0, st>>>(v, d_sp, d_ep, n, cnt, stash, glist, mlist, d_counters, medium);
  ull counters[6] = { 0, 0, 0, 0, 0, 0 };
  LOC_TRY(cudaMemcpyAsync(counters, d_counters, 6 * sizeof(ull), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));
  const u64 n_short = counters[1], n_block = counters[2], n_warp = counters[4], n_huge = counters[5];
  const u64* short_list = mlist; const u64* block_list = mlist + (n - n_block); const u64* warp_list = mlist + n;
  const u64* huge_list = mlist + (2 * n - n_huge);
  u64* scratch = nullptr;
  if(n_short + n_warp + n_block + n_huge > 0)
  {
    // the medium ranges: sorted and deduplicated in registers, a warp or a block per range; the grids are what is
    // resident at once (the groups stride over their list, so no wave is left partly filled)
    scratch = (u64*)alloc(counters[3] * sizeof(u64));
    if(!scratch) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
    static const MediumGrids grids = mediumGrids();
    if(n_short > 0)
    {
      u64 blocks = std::min<u64>((n_short + 3) / 4, (u64)sm * grids.per_sm[0]);
      locate_medium_kernel<0><<<(unsigned)blocks, 128, 0, st>>>(v, d_sp, d_ep, short_list, n_short, cnt, stash, scratch, glist, d_counters);
    }
    if(n_warp > 0)
    {
      u64 blocks = std::min<u64>((n_warp + 3) / 4, (u64)sm * grids.per_sm[1]);
      locate_medium_kernel<1><<<(unsigned)blocks, 128, 0, st>>>(v, d_sp, d_ep, warp_list, n_warp, cnt, stash, scratch, glist, d_counters);
    }
    if(n_block > 0)
    {
      u64 blocks = std::min<u64>(n_block, (u64)sm * grids.per_sm[2]);
      locate_medium_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(v, d_sp, d_ep, block_list, n_block, cnt, stash, scratch, glist, d_counters);
    }
    if(n_huge > 0)
    {
      u64 blocks = std::min<u64>(n_huge, (u64)sm * grids.per_sm[3]);
      locate_medium_kernel<3><<<(unsigned)blocks, 512, 0, st>>>(v, d_sp, d_ep, huge_list, n_huge, cnt, stash, scratch, glist, d_counters);
    }
    LOC_TRY(cudaMemcpyAsync(counters, d_counters, sizeof(ull), cudaMemcpyDeviceToHost, st));
    LOC_TRY(cudaStreamSynchronize(st));
  }
  const ull n_general = counters[0];
  if(std::getenv("GCSA_B200_LOCATE_DEBUG") != nullptr)
  {
    std::fprintf(stderr, "locate: of %llu ranges %llu + %llu sorted by a warp, %llu + %llu by a block, %llu through the general pipeline\n", (ull)n, (ull)n_short, (ull)n_warp, (ull)n_block, (ull)n_huge, n_general);
  }

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
  else if(distinct > 0)
  {
    locate_small_fill_kernel<<<gridFor(n, sm), 256, 0, st>>>(v, d_sp, d_ep, n, d_out_offsets, stash, d_values);
    if(n_short > 0) { locate_copy_kernel<false><<<gridFor(n_short * 32, sm), 256, 0, st>>>(short_list, n_short, stash, d_out_offsets, nullptr, scratch, d_values); }
    if(n_warp > 0) { locate_copy_kernel<false><<<gridFor(n_warp * 32, sm), 256, 0, st>>>(warp_list, n_warp, stash, d_out_offsets, nullptr, scratch, d_values); }
    if(n_block > 0) { locate_copy_kernel<false><<<gridFor(n_block * 32, sm), 256, 0, st>>>(block_list, n_block, stash, d_out_offsets, nullptr, scratch, d_values); }
    if(n_huge > 0) { locate_copy_kernel<false><<<gridFor(n_huge * 32, sm), 256, 0, st>>>(huge_list, n_huge, stash, d_out_offsets, nullptr, scratch, d_values); }
    if(n_general > 0) { locate_copy_kernel<true><<<gridFor(n_general * 32, sm), 256, 0, st>>>(glist, n_general, stash, d_out_offsets, goffs, gvals, d_values); }
  }
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
  full_sp.reserve(n); full_ep.reserve(n); full_id.reserve(n);
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
  // (sizes first, one pass per source; the common case -- no special range at all -- is two parallel loops)
  std::vector<u64> slot_of(n, ~(u64)0);                                  // position in the full locate, or none
  #pragma omp parallel for schedule(static)
  for(u64 t = 0; t < full_id.size(); t++) { slot_of[full_id[t]] = t; }
  for(u64 i = 0; i <= n; i++) { out_offsets[i] = 0; }
  #pragma omp parallel for schedule(static)
  for(u64 t = 0; t < full_id.size(); t++) { out_offsets[full_id[t] + 1] = full_offs[t + 1] - full_offs[t]; }
  for(auto& entry : special) { out_offsets[entry.first + 1] = entry.second.result.size(); }
  for(u64 i = 0; i < n; i++) { out_offsets[i + 1] += out_offsets[i]; }
  u64* vals = (u64*)std::malloc(std::max<u64>(out_offsets[n], 1) * sizeof(u64));
  if(vals == nullptr) { std::free(full_vals); return fail(GCSA_B200_ERR_NOMEM, "locate_max_host: out of host memory"); }
  #pragma omp parallel for schedule(static)
  for(u64 t = 0; t < full_id.size(); t++)
  {
    u64 i = full_id[t];
    if(out_offsets[i + 1] - out_offsets[i] == full_offs[t + 1] - full_offs[t])        // (a special range has its own, shorter result)
    {
      std::copy(full_vals + full_offs[t], full_vals + full_offs[t + 1], vals + out_offsets[i]);
    }
  }
  for(auto& entry : special) { std::copy(entry.second.result.begin(), entry.second.result.end(), vals + out_offsets[entry.first]); }
  std::free(full_vals);
  *values = (uint64_t*)vals;
  return 0;
}

