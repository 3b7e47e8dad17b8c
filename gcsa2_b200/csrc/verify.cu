/*
  verify.cu -- verifyIndex() of the reference (src/algorithms.cpp:101-295), device-resident.

  The reference queries the index with every distinct kmer label of the construction input, one OpenMP thread per
  chunk of labels, one query at a time.  Here the kmer records are uploaded once and everything up to and including
  locate() happens on the device, stage by stage over whole batches of labels:

    records -> (label, start node) pairs, sorted (two stable radix sorts)             parallelQuickSort, :106
    label groups, distinct start nodes per label = the expected occurrences             :117-125, :184-187
    patterns (Key::decode, cut after the first endmarker)                                :127-129
    find()                                                                               :131-143
    parent() == the first different range found by dropping characters from the right end, depth() == its lcp   :145-181
    count() == number of distinct start nodes                                            :183-200
    locate() == those nodes                                                              :202-234
    locate(range, 10): min(10, n) of them                                                :236-274

  Every stage calls the same entry points a caller would (gcsa_b200_find_batch, _parent_batch, _depth_batch,
  _count_batch, _locate_batch: device pointers, one stream), the comparisons are small kernels that count failures per
  stage, and a label that fails a stage is not looked at by the later ones, as in the reference.  Only the last stage
  goes through the host, and only for the labels with more than 10 occurrences: locate(range, max_positions) draws with
  std::mt19937_64 (gcsa_b200_locate_max_host); for a range with at most max_positions occurrences it is by definition
  locate(range) (src/gcsa.cpp:860-875), which the stage before has just compared.  The first version did every stage through the host entry points
  and spent its time sorting, building patterns and filtering on the host (20.5 s for the 58 M labels of cfg3; the
  device was busy for a fraction of a second).

  With a NodeMapping (the `mapping` argument of verifyIndex) the expected occurrences are the mapped start nodes
  (Node::map, src/support.cpp:604-612).
*/
#include "engine.h"

#include <chrono>

namespace
{

constexpr u64 RANDOM_LOCATE_SIZE = 10;        // algorithms.cpp:92
constexpr u64 CHUNK_LABELS = 16u << 20;       // labels per batch (bounds the device buffers: ~150 bytes per label)

enum { FAIL_FIND = 0, FAIL_PARENT, FAIL_DEPTH, FAIL_COUNT, FAIL_LOCATE, FAIL_RANDOM, TODO_LEFT, N_COUNTERS };

// (label, start node) of every record; the start node through the NodeMapping, if any
__global__ void __launch_bounds__(256)
vf_records_kernel(const u64* __restrict__ keys, const u64* __restrict__ from, u64 n, u64 map_first, const u64* __restrict__ map_ids, u64 map_size,
                  u64* __restrict__ labels, u64* __restrict__ nodes)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    labels[i] = keys[i] >> 16;                                          // Key::label, support.h:403
    u64 value = from[i], id = value >> 11;
    if(map_size > 0 && id >= map_first && id - map_first < map_size) { value = (map_ids[id - map_first] << 11) | (value & 0x7FF); }
    nodes[i] = value;
  }
}

// head[i]: first record of its label; distinct[i]: first record of its (label, start node)
__global__ void __launch_bounds__(256)
vf_flags_kernel(const u64* __restrict__ labels, const u64* __restrict__ nodes, u64 n, u64* __restrict__ head, u64* __restrict__ distinct)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    bool h = (i == 0 || labels[i] != labels[i - 1]);
    head[i] = (h ? 1 : 0);
    distinct[i] = (h || nodes[i] != nodes[i - 1] ? 1 : 0);
  }
}

// head_pos / distinct_pos: exclusive scans of the flags (n + 1 entries)
__global__ void __launch_bounds__(256)
vf_groups_kernel(const u64* __restrict__ labels, const u64* __restrict__ nodes, const u64* __restrict__ head_pos, const u64* __restrict__ distinct_pos,
                 u64 n, u64* __restrict__ group_label, u64* __restrict__ exp_offsets, u64* __restrict__ expected)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    if(head_pos[i + 1] != head_pos[i]) { u64 g = head_pos[i]; group_label[g] = labels[i]; exp_offsets[g] = distinct_pos[i]; }
    if(distinct_pos[i + 1] != distinct_pos[i]) { expected[distinct_pos[i]] = nodes[i]; }
    if(i == n - 1) { exp_offsets[head_pos[n]] = distinct_pos[n]; }
  }
}

// Key::decode (support.cpp:539-553), cut after the first endmarker (algorithms.cpp:127-129): the length
__global__ void __launch_bounds__(256)
vf_lengths_kernel(const u64* __restrict__ group_label, u64 m, u32 k, u64* __restrict__ lengths)
{
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g <= m; g += (u64)gridDim.x * blockDim.x)
  {
    u64 len = 0;
    if(g < m)
    {
      u64 label = group_label[g];
      for(u32 i = 0; i < k; i++) { len++; if(((label >> (3 * (k - 1 - i))) & 7) == 0) { break; } }
    }
    lengths[g] = len;
  }
}

__global__ void __launch_bounds__(256)
vf_chars_kernel(const u64* __restrict__ group_label, const u64* __restrict__ offsets, u64 m, u32 k, u8* __restrict__ chars)
{
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    const u64 label = group_label[g], start = offsets[g], len = offsets[g + 1] - start;
    for(u64 i = 0; i < len; i++)
    {
      u32 comp = (u32)((label >> (3 * (k - 1 - i))) & 7);
      chars[start + i] = (u8)("$ACGTN#N"[comp]);                          // src/support.cpp:92
    }
  }
}

// find(): an empty range is a failure (algorithms.cpp:132-143)
__global__ void __launch_bounds__(256)
vf_find_check_kernel(const u64* __restrict__ sp, const u64* __restrict__ ep, u64 m, u8* __restrict__ alive, ull* __restrict__ counters)
{
  ull fails = 0;
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    bool ok = !range_empty(sp[g], ep[g]);
    alive[g] = (ok ? 1 : 0);
    fails += (ok ? 0 : 1);
  }
  for(int d = 16; d > 0; d >>= 1) { fails += __shfl_down_sync(0xFFFFFFFFu, fails, d); }
  if((threadIdx.x & 31) == 0 && fails > 0) { atomicAdd(counters + FAIL_FIND, fails); }
}

// One round of "drop a character from the right end until the range changes": the length of the next shorter pattern
// of every label that is still looking (0 for the others)
__global__ void __launch_bounds__(256)
vf_shorter_kernel(const u8* __restrict__ todo, u64* __restrict__ q_len, u64 m, u64* __restrict__ sub_lengths)
{
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g <= m; g += (u64)gridDim.x * blockDim.x)
  {
    u64 len = 0;
    if(g < m && todo[g]) { len = --q_len[g]; }
    sub_lengths[g] = len;
  }
}

__global__ void __launch_bounds__(256)
vf_prefix_kernel(const u8* __restrict__ chars, const u64* __restrict__ offsets, const u64* __restrict__ sub_offsets, u64 m, u8* __restrict__ sub)
{
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    const u64 from = offsets[g], to = sub_offsets[g], len = sub_offsets[g + 1] - to;
    for(u64 i = 0; i < len; i++) { sub[to + i] = chars[from + i]; }
  }
}

__global__ void __launch_bounds__(256)
vf_shorter_update_kernel(const u64* __restrict__ sp, const u64* __restrict__ ep, const u64* __restrict__ s, const u64* __restrict__ e,
                         const u64* __restrict__ q_len, u64 m, u8* __restrict__ todo, u64* __restrict__ q_sp, u64* __restrict__ q_ep, ull* __restrict__ counters)
{
  ull left = 0;
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    if(!todo[g]) { continue; }
    q_sp[g] = s[g]; q_ep[g] = e[g];
    bool again = (s[g] == sp[g] && e[g] == ep[g] && q_len[g] > 0);
    todo[g] = (again ? 1 : 0);
    left += (again ? 1 : 0);
  }
  for(int d = 16; d > 0; d >>= 1) { left += __shfl_down_sync(0xFFFFFFFFu, left, d); }
  if((threadIdx.x & 31) == 0 && left > 0) { atomicAdd(counters + TODO_LEFT, left); }
}

// parent(range) must be that shorter range at that depth (algorithms.cpp:146-166); the ranges of the parents for depth()
__global__ void __launch_bounds__(256)
vf_parent_check_kernel(const gcsa_b200_stnode* __restrict__ parents, const u64* __restrict__ q_sp, const u64* __restrict__ q_ep,
                       const u64* __restrict__ q_len, u64 m, u8* __restrict__ alive, u64* __restrict__ psp, u64* __restrict__ pep, ull* __restrict__ counters)
{
  ull fails = 0;
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    const gcsa_b200_stnode p = parents[g];
    psp[g] = p.sp; pep[g] = p.ep;
    if(!alive[g]) { continue; }
    if(p.sp != q_sp[g] || p.ep != q_ep[g] || p.node_lcp != q_len[g]) { alive[g] = 0; fails++; }
  }
  for(int d = 16; d > 0; d >>= 1) { fails += __shfl_down_sync(0xFFFFFFFFu, fails, d); }
  if((threadIdx.x & 31) == 0 && fails > 0) { atomicAdd(counters + FAIL_PARENT, fails); }
}

// depth(parent range) == parent.lcp() (algorithms.cpp:167-180)
__global__ void __launch_bounds__(256)
vf_depth_check_kernel(const gcsa_b200_stnode* __restrict__ parents, const u64* __restrict__ depth, u64 m, u8* __restrict__ alive, ull* __restrict__ counters)
{
  ull fails = 0;
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    if(alive[g] && depth[g] != parents[g].node_lcp) { alive[g] = 0; fails++; }
  }
  for(int d = 16; d > 0; d >>= 1) { fails += __shfl_down_sync(0xFFFFFFFFu, fails, d); }
  if((threadIdx.x & 31) == 0 && fails > 0) { atomicAdd(counters + FAIL_DEPTH, fails); }
}

// count(range) == number of distinct start nodes (algorithms.cpp:183-200); flags the labels that go on to locate()
__global__ void __launch_bounds__(256)
vf_count_check_kernel(const u64* __restrict__ counts, const u64* __restrict__ exp_offsets, u64 m, u8* __restrict__ alive, u64* __restrict__ keep, ull* __restrict__ counters)
{
  ull fails = 0;
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g <= m; g += (u64)gridDim.x * blockDim.x)
  {
    u64 flag = 0;
    if(g < m && alive[g])
    {
      if(counts[g] != exp_offsets[g + 1] - exp_offsets[g]) { alive[g] = 0; fails++; } else { flag = 1; }
    }
    keep[g] = flag;
  }
  for(int d = 16; d > 0; d >>= 1) { fails += __shfl_down_sync(0xFFFFFFFFu, fails, d); }
  if((threadIdx.x & 31) == 0 && fails > 0) { atomicAdd(counters + FAIL_COUNT, fails); }
}

// the labels that passed count(): their number in the chunk, their range; keep_pos = exclusive scan of keep
__global__ void __launch_bounds__(256)
vf_select_kernel(const u64* __restrict__ keep_pos, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 m,
                 u64* __restrict__ ids, u64* __restrict__ a, u64* __restrict__ b)
{
  for(u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (u64)gridDim.x * blockDim.x)
  {
    if(keep_pos[g + 1] != keep_pos[g]) { u64 d = keep_pos[g]; ids[d] = g; a[d] = sp[g]; b[d] = ep[g]; }
  }
}

// locate(range) == the distinct start nodes (algorithms.cpp:202-234); random_flag: the size was right (the reference tries
// the random locate after a value mismatch too) and the label has more than 10 occurrences
__global__ void __launch_bounds__(256)
vf_locate_check_kernel(const u64* __restrict__ ids, u64 n_ids, const u64* __restrict__ loc_offsets, const u64* __restrict__ located,
                       const u64* __restrict__ exp_offsets, const u64* __restrict__ expected, u64* __restrict__ random_flag, ull* __restrict__ counters)
{
  ull fails = 0;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= n_ids; i += (u64)gridDim.x * blockDim.x)
  {
    u64 flag = 0;
    if(i < n_ids)
    {
      const u64 g = ids[i], got = loc_offsets[i + 1] - loc_offsets[i], want = exp_offsets[g + 1] - exp_offsets[g];
      bool same = (got == want);
      for(u64 j = 0; same && j < want; j++) { same = (located[loc_offsets[i] + j] == expected[exp_offsets[g] + j]); }
      fails += (same ? 0 : 1);
      // locate(range, 10) of a range with at most 10 occurrences IS locate(range) (gcsa.cpp:860-875: everything is located,
      // and nothing is dropped because nothing exceeds max_positions): the reference compares locate() with itself there.
      // The labels with more occurrences are the ones that draw random numbers; they go through the real entry point.
      flag = (got == want && want > RANDOM_LOCATE_SIZE ? 1 : 0);
    }
    random_flag[i] = flag;
  }
  for(int d = 16; d > 0; d >>= 1) { fails += __shfl_down_sync(0xFFFFFFFFu, fails, d); }
  if((threadIdx.x & 31) == 0 && fails > 0) { atomicAdd(counters + FAIL_LOCATE, fails); }
}

// Device buffers of one verification, freed together.
struct Buffers
{
  cudaStream_t stream = nullptr;
  std::vector<void*> all;
  template<class T> T* get(u64 count)
  {
    T* p = nullptr;
    if(engineMallocAsync(&p, std::max<u64>(count, 1) * sizeof(T) + 64, stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    all.push_back(p);
    return p;
  }
  void release(void* p)
  {
    for(size_t i = 0; i < all.size(); i++) { if(all[i] == p) { all.erase(all.begin() + i); cudaFreeAsync(p, stream); return; } }
  }
  ~Buffers()
  {
    for(void* p : all) { cudaFreeAsync(p, stream); }
    if(stream != nullptr) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
  }
};

int sortPairs(Buffers& buf, const u64* keys_in, u64* keys_out, const u64* vals_in, u64* vals_out, u64 n, int end_bit)
{
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, buf.stream));
  u8* tmp = buf.get<u8>(bytes);
  if(tmp == nullptr) { return fail(GCSA_B200_ERR_NOMEM, "verify_index: out of device memory"); }
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, buf.stream);
  buf.release(tmp);
  CUDA_TRY(e);
  return 0;
}

} // namespace

extern "C" int gcsa_b200_verify_index(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint64_t* keys,
                                      const uint64_t* from, uint64_t n, int kmer_length, gcsa_b200_verify_report* report)
{
  return gcsa_b200_verify_index_mapped(index, lcp, keys, from, n, kmer_length, 0, nullptr, 0, report);
}

extern "C" int gcsa_b200_verify_index_mapped(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint64_t* keys,
                                             const uint64_t* from, uint64_t n, int kmer_length, uint64_t mapping_first_node,
                                             const uint64_t* mapping_ids, uint64_t mapping_size, gcsa_b200_verify_report* report)
{
  if(index == nullptr || report == nullptr || (n > 0 && (keys == nullptr || from == nullptr)) || kmer_length < 1 || kmer_length > 16 ||
     (mapping_size > 0 && mapping_ids == nullptr))
  {
    return fail(GCSA_B200_ERR_INVALID, "verify_index: bad argument");
  }
  if(lcp != nullptr && lcp->device != index->device) { return fail(GCSA_B200_ERR_INVALID, "verify_index: the index and the LCP array live on different devices"); }
  std::memset(report, 0, sizeof(*report));
  if(n == 0) { return 0; }
  const auto started = std::chrono::steady_clock::now();
  double host_seconds = 0.0;                                              // the last stage's host part
  const bool debug = (std::getenv("GCSA_B200_VERIFY_DEBUG") != nullptr);
  DeviceGuard guard(index->device);
  Buffers buf;
  CUDA_TRY(cudaStreamCreateWithFlags(&buf.stream, cudaStreamNonBlocking));
  cudaStream_t st = buf.stream;
  const int sm = index->sm_count;
  const u32 k = (u32)kmer_length;
  auto lap_start = started;
  auto lap = [&](const char* what)
  {
    if(!debug) { return; }
    cudaStreamSynchronize(st);
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "verify: %-28s %.3f s\n", what, std::chrono::duration<double>(now - lap_start).count());
    lap_start = now;
  };
  #define VF_ALLOC(var, type, count) type* var = buf.get<type>(count); if(var == nullptr) { return fail(GCSA_B200_ERR_NOMEM, "verify_index: out of device memory"); }
  #define VF_RC(expr) do { int rc_ = (expr); if(rc_ != 0) { return rc_; } } while(0)

  // ---- the records, sorted by (label, start node) ----
  VF_ALLOC(labels, u64, n); VF_ALLOC(nodes, u64, n);
  {
    VF_ALLOC(d_keys, u64, n); VF_ALLOC(d_from, u64, n);
    u64* d_map = nullptr;
    if(mapping_size > 0)
    {
      d_map = buf.get<u64>(mapping_size);
      if(d_map == nullptr) { return fail(GCSA_B200_ERR_NOMEM, "verify_index: out of device memory"); }
      CUDA_TRY(cudaMemcpyAsync(d_map, mapping_ids, mapping_size * sizeof(u64), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(d_keys, keys, n * sizeof(u64), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_from, from, n * sizeof(u64), cudaMemcpyHostToDevice, st));
    vf_records_kernel<<<gridFor(n, sm), 256, 0, st>>>(d_keys, d_from, n, mapping_first_node, d_map, mapping_size, labels, nodes);
    // stable sorts: by start node, then by label (3 bits per character)
    VF_RC(sortPairs(buf, nodes, d_from, labels, d_keys, n, 64));          // -> (d_from = nodes, d_keys = labels) by node
    VF_RC(sortPairs(buf, d_keys, labels, d_from, nodes, n, 3 * (int)k));  // -> (labels, nodes) by label, nodes ascending within
    buf.release(d_keys); buf.release(d_from);
    if(d_map) { buf.release(d_map); }
  }
  lap("upload + sort");

  // ---- label groups and expected occurrences ----
  u64 unique = 0, total_expected = 0;
  u64 *group_label = nullptr, *exp_offsets = nullptr, *expected = nullptr;
  {
    VF_ALLOC(head, u64, n + 1); VF_ALLOC(distinct, u64, n + 1);
    vf_flags_kernel<<<gridFor(n, sm), 256, 0, st>>>(labels, nodes, n, head, distinct);
    CUDA_TRY(cudaMemsetAsync(head + n, 0, sizeof(u64), st)); CUDA_TRY(cudaMemsetAsync(distinct + n, 0, sizeof(u64), st));
    VF_RC(scanExclusive(head, head, n + 1, st));
    VF_RC(scanExclusive(distinct, distinct, n + 1, st));
    CUDA_TRY(cudaMemcpyAsync(&unique, head + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&total_expected, distinct + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    group_label = buf.get<u64>(unique); exp_offsets = buf.get<u64>(unique + 1); expected = buf.get<u64>(total_expected);
    if(group_label == nullptr || exp_offsets == nullptr || expected == nullptr) { return fail(GCSA_B200_ERR_NOMEM, "verify_index: out of device memory"); }
    vf_groups_kernel<<<gridFor(n, sm), 256, 0, st>>>(labels, nodes, head, distinct, n, group_label, exp_offsets, expected);
    CUDA_TRY(cudaGetLastError());
    buf.release(head); buf.release(distinct); buf.release(labels); buf.release(nodes);
  }
  report->unique = unique;
  lap("groups + expected");

  VF_ALLOC(counters, ull, N_COUNTERS);
  CUDA_TRY(cudaMemsetAsync(counters, 0, N_COUNTERS * sizeof(ull), st));

  for(u64 base = 0; base < unique; base += CHUNK_LABELS)
  {
    const u64 m = std::min<u64>(CHUNK_LABELS, unique - base);
    const u64* labels_c = group_label + base;
    const u64* exp_c = exp_offsets + base;                              // (offsets into `expected`, batch-wide)

    // patterns
    VF_ALLOC(offsets, u64, m + 1);
    vf_lengths_kernel<<<gridFor(m + 1, sm), 256, 0, st>>>(labels_c, m, k, offsets);
    VF_RC(scanExclusive(offsets, offsets, m + 1, st));
    VF_ALLOC(chars, u8, m * k + 16);
    vf_chars_kernel<<<gridFor(m, sm), 256, 0, st>>>(labels_c, offsets, m, k, chars);

    // find()
    VF_ALLOC(sp, u64, m); VF_ALLOC(ep, u64, m); VF_ALLOC(alive, u8, m);
    VF_RC(gcsa_b200_find_batch(index, chars, offsets, m, sp, ep, st));
    vf_find_check_kernel<<<gridFor(m, sm), 256, 0, st>>>(sp, ep, m, alive, counters);
    lap("patterns + find");

    // parent() and depth()
    if(lcp != nullptr)
    {
      VF_ALLOC(parents, gcsa_b200_stnode, m);
      VF_RC(gcsa_b200_parent_batch(lcp, sp, ep, m, parents, st));
      VF_ALLOC(todo, u8, m); VF_ALLOC(q_sp, u64, m); VF_ALLOC(q_ep, u64, m); VF_ALLOC(q_len, u64, m);
      VF_ALLOC(sub_offsets, u64, m + 1); VF_ALLOC(sub, u8, m * k + 16); VF_ALLOC(s, u64, m); VF_ALLOC(e, u64, m);
      CUDA_TRY(cudaMemcpyAsync(todo, alive, m, cudaMemcpyDeviceToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(q_sp, sp, m * sizeof(u64), cudaMemcpyDeviceToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(q_ep, ep, m * sizeof(u64), cudaMemcpyDeviceToDevice, st));
      // q_len = the pattern lengths
      vf_lengths_kernel<<<gridFor(m + 1, sm), 256, 0, st>>>(labels_c, m, k, sub_offsets);
      CUDA_TRY(cudaMemcpyAsync(q_len, sub_offsets, m * sizeof(u64), cudaMemcpyDeviceToDevice, st));
      for(u32 round = 0; round <= k; round++)
      {
        CUDA_TRY(cudaMemsetAsync(counters + TODO_LEFT, 0, sizeof(ull), st));
        vf_shorter_kernel<<<gridFor(m + 1, sm), 256, 0, st>>>(todo, q_len, m, sub_offsets);
        VF_RC(scanExclusive(sub_offsets, sub_offsets, m + 1, st));
        vf_prefix_kernel<<<gridFor(m, sm), 256, 0, st>>>(chars, offsets, sub_offsets, m, sub);
        VF_RC(gcsa_b200_find_batch(index, sub, sub_offsets, m, s, e, st));
        vf_shorter_update_kernel<<<gridFor(m, sm), 256, 0, st>>>(sp, ep, s, e, q_len, m, todo, q_sp, q_ep, counters);
        ull left = 0;
        CUDA_TRY(cudaMemcpyAsync(&left, counters + TODO_LEFT, sizeof(ull), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if(left == 0) { break; }
      }
      VF_ALLOC(psp, u64, m); VF_ALLOC(pep, u64, m); VF_ALLOC(depth, u64, m);
      vf_parent_check_kernel<<<gridFor(m, sm), 256, 0, st>>>(parents, q_sp, q_ep, q_len, m, alive, psp, pep, counters);
      VF_RC(gcsa_b200_depth_batch(lcp, psp, pep, m, depth, st));
      vf_depth_check_kernel<<<gridFor(m, sm), 256, 0, st>>>(parents, depth, m, alive, counters);
      for(void* p : { (void*)parents, (void*)todo, (void*)q_sp, (void*)q_ep, (void*)q_len, (void*)sub_offsets, (void*)sub, (void*)s, (void*)e, (void*)psp, (void*)pep, (void*)depth }) { buf.release(p); }
      lap("parent + depth");
    }

    // count()
    VF_ALLOC(counts, u64, m); VF_ALLOC(keep, u64, m + 1);
    VF_RC(gcsa_b200_count_batch(index, sp, ep, m, counts, st));
    vf_count_check_kernel<<<gridFor(m + 1, sm), 256, 0, st>>>(counts, exp_c, m, alive, keep, counters);
    VF_RC(scanExclusive(keep, keep, m + 1, st));
    u64 n_ids = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_ids, keep + m, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    lap("count");

    // locate()
    if(n_ids > 0)
    {
      VF_ALLOC(ids, u64, n_ids); VF_ALLOC(a, u64, n_ids); VF_ALLOC(b, u64, n_ids);
      vf_select_kernel<<<gridFor(m, sm), 256, 0, st>>>(keep, sp, ep, m, ids, a, b);
      VF_ALLOC(loc_offsets, u64, n_ids + 1);
      // count() of a label that got here equals its expected number of nodes: at most that many values in this chunk
      u64 exp_first = 0, exp_last = 0;
      CUDA_TRY(cudaMemcpyAsync(&exp_first, exp_c, sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(&exp_last, exp_c + m, sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      u64 capacity = exp_last - exp_first + 16, needed = 0;
      u64* located = buf.get<u64>(capacity);
      if(located == nullptr) { return fail(GCSA_B200_ERR_NOMEM, "verify_index: out of device memory"); }
      int rc = gcsa_b200_locate_batch(index, a, b, n_ids, loc_offsets, located, capacity, &needed, st);
      if(rc == GCSA_B200_ERR_CAPACITY)                                   // locate() returned more than count() promised: the check below will say so
      {
        buf.release(located); capacity = needed + 16;
        located = buf.get<u64>(capacity);
        if(located == nullptr) { return fail(GCSA_B200_ERR_NOMEM, "verify_index: out of device memory"); }
        rc = gcsa_b200_locate_batch(index, a, b, n_ids, loc_offsets, located, capacity, &needed, st);
      }
      if(rc != 0) { return rc; }
      VF_ALLOC(random_flag, u64, n_ids + 1);
      vf_locate_check_kernel<<<gridFor(n_ids + 1, sm), 256, 0, st>>>(ids, n_ids, loc_offsets, located, exp_c, expected, random_flag, counters);
      CUDA_TRY(cudaGetLastError());
      lap("locate");

      // locate(range, 10) -- algorithms.cpp:236-274.  The random draws are the host's (std::mt19937_64): if any label of
      // the chunk has more than 10 occurrences, the ranges and the located values come down for this stage.
      VF_RC(scanExclusive(random_flag, random_flag, n_ids + 1, st));
      u64 n_random = 0;
      CUDA_TRY(cudaMemcpyAsync(&n_random, random_flag + n_ids, sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if(n_random == 0)
      {
        for(void* p : { (void*)ids, (void*)a, (void*)b, (void*)loc_offsets, (void*)located, (void*)random_flag }) { buf.release(p); }
        lap("locate(range, 10): none");
      }
      else
      {
      std::vector<u64> h_a(n_ids), h_b(n_ids), h_offs(n_ids + 1), h_flag(n_ids + 1), h_located(std::max<u64>(needed, 1));
      CUDA_TRY(cudaMemcpyAsync(h_a.data(), a, n_ids * sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(h_b.data(), b, n_ids * sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(h_offs.data(), loc_offsets, (n_ids + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(h_flag.data(), random_flag, (n_ids + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
      if(needed > 0) { CUDA_TRY(cudaMemcpyAsync(h_located.data(), located, needed * sizeof(u64), cudaMemcpyDeviceToHost, st)); }
      CUDA_TRY(cudaStreamSynchronize(st));
      for(void* p : { (void*)ids, (void*)a, (void*)b, (void*)loc_offsets, (void*)located, (void*)random_flag }) { buf.release(p); }
      const auto host_start = std::chrono::steady_clock::now();
      std::vector<u64> pick;                                             // positions in ids whose locate() had the right size
      for(u64 i = 0; i < n_ids; i++) { if(h_flag[i + 1] != h_flag[i]) { pick.push_back(i); } }       // (h_flag holds the exclusive scan)
      std::vector<u64> ra(pick.size() + 1), rb(pick.size() + 1), rnd_offsets(pick.size() + 1, 0);
      #pragma omp parallel for schedule(static)
      for(u64 i = 0; i < pick.size(); i++) { ra[i] = h_a[pick[i]]; rb[i] = h_b[pick[i]]; }
      uint64_t* randoms = nullptr;
      rc = gcsa_b200_locate_max_host(index, ra.data(), rb.data(), pick.size(), RANDOM_LOCATE_SIZE, rnd_offsets.data(), &randoms);
      if(rc != 0) { return rc; }
      u64 random_fails = 0;
      #pragma omp parallel for schedule(static) reduction(+:random_fails)
      for(u64 i = 0; i < pick.size(); i++)
      {
        const u64 q = pick[i];
        const u64* occs = h_located.data() + h_offs[q]; const u64 n_occs = h_offs[q + 1] - h_offs[q];
        const u64* rnd = randoms + rnd_offsets[i]; const u64 n_rnd = rnd_offsets[i + 1] - rnd_offsets[i];
        if(n_rnd != std::min(RANDOM_LOCATE_SIZE, n_occs)) { random_fails++; continue; }
        bool subset = true;
        for(u64 x = 0, y = 0; x < n_rnd; x++)
        {
          while(y + 1 < n_occs && occs[y] < rnd[x]) { y++; }
          if(y >= n_occs || rnd[x] != occs[y]) { subset = false; break; }
          y++;
        }
        if(!subset) { random_fails++; }
      }
      gcsa_b200_free(randoms);
      report->random_locate_failures += random_fails;
      host_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - host_start).count();
      lap("locate(range, 10)");
      }
    }
    for(void* p : { (void*)offsets, (void*)chars, (void*)sp, (void*)ep, (void*)alive, (void*)counts, (void*)keep }) { buf.release(p); }
  }

  ull h_counters[N_COUNTERS];
  CUDA_TRY(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  #undef VF_ALLOC
  #undef VF_RC
  report->find_failures = h_counters[FAIL_FIND]; report->parent_failures = h_counters[FAIL_PARENT];
  report->depth_failures = h_counters[FAIL_DEPTH]; report->count_failures = h_counters[FAIL_COUNT];
  report->locate_failures = h_counters[FAIL_LOCATE];
  report->failures = report->find_failures + report->parent_failures + report->depth_failures + report->count_failures +
                     report->locate_failures + report->random_locate_failures;
  report->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - started).count();
  report->engine_seconds = report->seconds - host_seconds;
  return GCSA_B200_OK;
}
