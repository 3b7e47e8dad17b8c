/* device/kmers.cuh -- countKMers / compareKMers as breadth-first frontier expansion.
   Part of the single translation unit engine.cu (included there in order); sm_100a only. */
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

// states: 4 arrays (left sp, left ep, right sp, right ep) of `stride` entries each; kmers: 3 words per state or null.
__global__ void __launch_bounds__(256)
compare_expand_kernel(const DevView vl, const DevView vr, const u64* __restrict__ in, u64 n, const u64* __restrict__ in_kmer,
                      u32 chars, u64 level, u64* __restrict__ out, u64* __restrict__ out_kmer, u64* __restrict__ flag)
{
  u64 total = n * chars;
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    u64 i = t / chars; u32 c = (u32)(t - i * chars) + 1;
    u64 la, lb, ra, rb;
    trie_child(vl, in[i], in[n + i], c, la, lb);
    trie_child(vr, in[2 * n + i], in[3 * n + i], c, ra, rb);
    out[t] = la; out[total + t] = lb; out[2 * total + t] = ra; out[3 * total + t] = rb;
    flag[t] = ((range_empty(la, lb) && range_empty(ra, rb)) ? 0 : 1);          // algorithms.cpp:514
    if(out_kmer != nullptr)
    {
      u64 w0 = in_kmer[3 * i], w1 = in_kmer[3 * i + 1], w2 = in_kmer[3 * i + 2];
      u64 bit = level * 3, word = bit >> 6, off = bit & 63, x = (u64)c << off, y = (off > 61 ? (u64)c >> (64 - off) : 0);   // KMerComparisonState::set, algorithms.cpp:451-457
      if(word == 0) { w0 |= x; w1 |= y; } else if(word == 1) { w1 |= x; w2 |= y; } else { w2 |= x; }
      out_kmer[3 * t] = w0; out_kmer[3 * t + 1] = w1; out_kmer[3 * t + 2] = w2;
    }
  }
}

__global__ void __launch_bounds__(256)
compare_compact_kernel(const u64* __restrict__ child, const u64* __restrict__ child_kmer, const u64* __restrict__ flag,
                       const u64* __restrict__ pos, u64 total, u64 next, u64* __restrict__ out, u64* __restrict__ out_kmer)
{
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    if(!flag[t]) { continue; }
    u64 d = pos[t];
    for(int f = 0; f < 4; f++) { out[f * next + d] = child[f * total + t]; }
    if(out_kmer != nullptr) { for(int w = 0; w < 3; w++) { out_kmer[3 * d + w] = child_kmer[3 * t + w]; } }
  }
}

// KMerSymmetricDifference::report, algorithms.cpp:488-500: side[i] = 0 shared, 1 left only, 2 right only.
__global__ void __launch_bounds__(256)
compare_classify_kernel(const u64* __restrict__ st, u64 n, ull* __restrict__ counts, u64* __restrict__ left_flag, u64* __restrict__ right_flag)
{
  ull shared_n = 0, left_n = 0, right_n = 0;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 llen = st[n + i] + 1 - st[i], rlen = st[3 * n + i] + 1 - st[2 * n + i];
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
compare_emit_kernel(const u64* __restrict__ st, const u64* __restrict__ kmer, u64 n, u64 k, const u64* __restrict__ flag,
                    const u64* __restrict__ pos, u64* __restrict__ records)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    if(!flag[i]) { continue; }
    u64* r = records + 8 * pos[i];
    r[0] = st[i]; r[1] = st[n + i]; r[2] = st[2 * n + i]; r[3] = st[3 * n + i]; r[4] = k;
    r[5] = kmer[3 * i]; r[6] = kmer[3 * i + 1]; r[7] = kmer[3 * i + 2];
  }
}

#endif
