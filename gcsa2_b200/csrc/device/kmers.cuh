/* device/kmers.cuh -- countKMers / compareKMers as breadth-first frontier expansion.
   Included by kmers.cu only; sm_100a only. */
#ifndef GCSA2_B200_DEVICE_KMERS_CUH
#define GCSA2_B200_DEVICE_KMERS_CUH

//------------------------------------------------------------------------------
// Kernels: countKMers (src/algorithms.cpp:364-421) as breadth-first frontier expansion
//------------------------------------------------------------------------------

// One thread per (frontier range, comp): the child range of processSubtree()'s expansion
// (LF_fast for bases, LF_all with N), and whether it survives (non-empty).
__global__ void __launch_bounds__(256)
kmer_expand_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, u32 chars,
                   u64* __restrict__ csp, u64* __restrict__ cep, u64* __restrict__ flag)
{
  u64 total = n * chars;
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    u64 i = t / chars; u32 c = (u32)(t - i * chars) + 1;
    u64 s = sp[i], e = ep[i], a = 1, b = 0;
    if(s == e)                                                                // gcsa.cpp:748-756, 774-789
    {
      // a single path node: the bit test and the step read the same sector -- one load, not two
      if(c <= GCSA_B200_FAST_CHARS) { u64 p; if(pred_fast(v, s, c - 1, p)) { a = b = p; } }
      else if(bwt_bit(v, s, c)) { lf_range(v, s, e, c, a, b); }
    }
    else { lf_range(v, s, e, c, a, b); }
    csp[t] = a; cep[t] = b; flag[t] = (range_empty(a, b) ? 0 : 1);
  }
}

__global__ void __launch_bounds__(256)
kmer_compact_kernel(const u64* __restrict__ csp, const u64* __restrict__ cep, const u64* __restrict__ flag,
                    const u64* __restrict__ pos, u64 total, u64* __restrict__ sp, u64* __restrict__ ep)
{
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    if(flag[t]) { sp[pos[t]] = csp[t]; ep[pos[t]] = cep[t]; }
  }
}

__device__ __forceinline__ void trie_child(const DevView& v, u64 s, u64 e, u32 c, u64& a, u64& b);

// The same level in ONE kernel when only the NUMBER of k-mers is wanted (the order of the frontier is then free):
// one thread per frontier range computes all its children -- for a single path node its four fast sectors are one
// 128-byte line, read once -- and the surviving ones are appended to the next frontier at a position taken from a
// counter (one atomic per warp).  No flags, no scan, no second pass, no host round trip inside a level.
__global__ void __launch_bounds__(256)
kmer_level_kernel(const DevView v, const ulonglong2* __restrict__ in, u64 n, u32 chars, ulonglong2* __restrict__ out,
                  unsigned long long* __restrict__ out_count, u64 capacity)
{
  const u32 lane = threadIdx.x & 31;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 rounds = (n + stride - 1) / stride;
  for(u64 r = 0; r < rounds; r++)
  {
    const u64 i = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 a[6], b[6]; u32 alive = 0;
    if(i < n)
    {
      const ulonglong2 range = __ldcs(in + i);
      #pragma unroll
      for(u32 c = 1; c <= 5; c++)
      {
        a[c] = 1; b[c] = 0;
        if(c > chars) { continue; }
        trie_child(v, range.x, range.y, c, a[c], b[c]);
        if(!range_empty(a[c], b[c])) { alive |= 1u << c; }
      }
    }
    // positions in the next frontier: exclusive prefix of the survivor counts over the warp
    const u32 mine = (u32)__popc(alive);
    u32 before = mine;
    #pragma unroll
    for(int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, before, d); if(lane >= (u32)d) { before += o; } }
    const u32 warp_total = __shfl_sync(0xFFFFFFFFu, before, 31);
    before -= mine;
    unsigned long long base = 0;
    if(lane == 0 && warp_total > 0) { base = atomicAdd(out_count, (unsigned long long)warp_total); }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    u64 at = base + before;
    #pragma unroll
    for(u32 c = 1; c <= 5; c++)
    {
      if((alive >> c) & 1) { if(at < capacity) { __stcs((ulonglong2*)out + at, make_ulonglong2(a[c], b[c])); } at++; }
    }
  }
}

//------------------------------------------------------------------------------
// Kernels: compareKMers (src/algorithms.cpp:505-616) -- the tries of two indexes in lockstep
//------------------------------------------------------------------------------

// One child range of LF_fast / LF_all (src/gcsa.cpp:742-798): empty input and a single path node
// without the predecessor give Range::empty_range(); the general case gives LF() uncanonicalised.
__device__ __forceinline__ void trie_child(const DevView& v, u64 s, u64 e, u32 c, u64& a, u64& b)
{
  a = 1; b = 0;
  if(range_empty(s, e)) { return; }
  if(s == e)
  {
    if(c <= GCSA_B200_FAST_CHARS) { u64 p; if(pred_fast(v, s, c - 1, p)) { a = b = p; } }      // one sector: bit and step
    else if(bwt_bit(v, s, c)) { lf_range(v, s, e, c, a, b); }
  }
  else { lf_range(v, s, e, c, a, b); }
}

// One level of the two tries in lockstep, in one kernel: one thread per state (a pair of ranges, left and right index)
// computes its children in both indexes and appends those that are alive in either (algorithms.cpp:514) to the next
// frontier at a position taken from a counter (one atomic per warp).  The order of the frontier is free: the counts
// do not depend on it and the reference's own output order depends on thread scheduling.
// states: 4 arrays (left sp, left ep, right sp, right ep) of `stride` entries each; kmers: 3 words per state or null.
__global__ void __launch_bounds__(256)
compare_level_kernel(const DevView vl, const DevView vr, const u64* __restrict__ in, u64 n, u64 in_stride, const u64* __restrict__ in_kmer,
                     u32 chars, u64 level, u64* __restrict__ out, u64 out_stride, u64* __restrict__ out_kmer,
                     unsigned long long* __restrict__ out_count)
{
  const u32 lane = threadIdx.x & 31;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 rounds = (n + stride - 1) / stride;
  for(u64 r = 0; r < rounds; r++)
  {
    const u64 i = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 la[6], lb[6], ra[6], rb[6]; u32 alive = 0;
    if(i < n)
    {
      const u64 ls = in[i], le = in[in_stride + i], rs = in[2 * in_stride + i], re = in[3 * in_stride + i];
      #pragma unroll
      for(u32 c = 1; c <= 5; c++)
      {
        la[c] = 1; lb[c] = 0; ra[c] = 1; rb[c] = 0;
        if(c > chars) { continue; }
        trie_child(vl, ls, le, c, la[c], lb[c]);
        trie_child(vr, rs, re, c, ra[c], rb[c]);
        if(!(range_empty(la[c], lb[c]) && range_empty(ra[c], rb[c]))) { alive |= 1u << c; }
      }
    }
    const u32 mine = (u32)__popc(alive);
    u32 before = mine;
    #pragma unroll
    for(int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, before, d); if(lane >= (u32)d) { before += o; } }
    const u32 warp_total = __shfl_sync(0xFFFFFFFFu, before, 31);
    before -= mine;
    unsigned long long base = 0;
    if(lane == 0 && warp_total > 0) { base = atomicAdd(out_count, (unsigned long long)warp_total); }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    u64 at = base + before;
    #pragma unroll
    for(u32 c = 1; c <= 5; c++)
    {
      if(!((alive >> c) & 1)) { continue; }
      if(at < out_stride)
      {
        out[at] = la[c]; out[out_stride + at] = lb[c]; out[2 * out_stride + at] = ra[c]; out[3 * out_stride + at] = rb[c];
        if(out_kmer != nullptr)
        {
          u64 w0 = in_kmer[3 * i], w1 = in_kmer[3 * i + 1], w2 = in_kmer[3 * i + 2];
          u64 bit = level * 3, word = bit >> 6, off = bit & 63, x = (u64)c << off, y = (off > 61 ? (u64)c >> (64 - off) : 0);   // KMerComparisonState::set, algorithms.cpp:451-457
          if(word == 0) { w0 |= x; w1 |= y; } else if(word == 1) { w1 |= x; w2 |= y; } else { w2 |= x; }
          out_kmer[3 * at] = w0; out_kmer[3 * at + 1] = w1; out_kmer[3 * at + 2] = w2;
        }
      }
      at++;
    }
  }
}

// KMerSymmetricDifference::report, algorithms.cpp:488-500: side[i] = 0 shared, 1 left only, 2 right only.
__global__ void __launch_bounds__(256)
compare_classify_kernel(const u64* __restrict__ st, u64 n, u64 st_stride, ull* __restrict__ counts, u64* __restrict__ left_flag, u64* __restrict__ right_flag)
{
  ull shared_n = 0, left_n = 0, right_n = 0;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 llen = st[st_stride + i] + 1 - st[i], rlen = st[3 * st_stride + i] + 1 - st[2 * st_stride + i];
    u32 side = (llen > 0 && rlen > 0 ? 0 : (llen > 0 ? 1 : 2));
    shared_n += (side == 0); left_n += (side == 1); right_n += (side == 2);
    if(left_flag != nullptr) { left_flag[i] = (side == 1); right_flag[i] = (side == 2); }
  }
  #pragma unroll
  for(int k = 0; k < 3; k++)
  {
    ull x = (k == 0 ? shared_n : (k == 1 ? left_n : right_n));
    for(int d = 16; d > 0; d >>= 1) { x += __shfl_down_sync(0xFFFFFFFFu, x, d); }
    if((threadIdx.x & 31) == 0 && x > 0) { atomicAdd(counts + k, x); }
  }
}

// Unique kmers as gcsa_b200_kmer_state records (8 words each).
__global__ void __launch_bounds__(256)
compare_emit_kernel(const u64* __restrict__ st, u64 st_stride, const u64* __restrict__ kmer, u64 n, u64 k, const u64* __restrict__ flag,
                    const u64* __restrict__ pos, u64* __restrict__ records)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    if(!flag[i]) { continue; }
    u64* r = records + 8 * pos[i];
    r[0] = st[i]; r[1] = st[st_stride + i]; r[2] = st[2 * st_stride + i]; r[3] = st[3 * st_stride + i]; r[4] = k;
    r[5] = kmer[3 * i]; r[6] = kmer[3 * i + 1]; r[7] = kmer[3 * i + 2];
  }
}

#endif
