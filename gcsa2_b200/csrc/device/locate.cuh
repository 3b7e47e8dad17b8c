/* device/locate.cuh -- locate(): the general pipeline, the short-range register path, the walk / locate / jump tables.
   Included by locate.cu only; sm_100a only. */
#ifndef GCSA2_B200_DEVICE_LOCATE_CUH
#define GCSA2_B200_DEVICE_LOCATE_CUH

//------------------------------------------------------------------------------
// Kernels: locate (src/gcsa.cpp:827-842, 880-896)
//------------------------------------------------------------------------------

// number of path nodes each range contributes (0 for empty / out-of-range ranges, gcsa.cpp:831)
__global__ void __launch_bounds__(256)
locate_lengths_kernel(u64 path_nodes, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, u64* __restrict__ len)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 s = sp[i], e = ep[i];
    len[i] = ((range_empty(s, e) || e >= path_nodes) ? 0 : e + 1 - s);
  }
}

// Last r in [0, n) with off[r] <= t (off[0] = 0), starting from a guess: gallop, then bisect.  With
// ranges of similar length the guess is off by a few entries and the search costs 2-3 loads
// instead of log2(n).
__device__ __forceinline__ u64 owner_of(const u64* __restrict__ off, u64 n, u64 t, u64 guess)
{
  u64 g = (guess < n ? guess : n - 1), lo, hi;
  if(__ldg(off + g) <= t)
  {
    lo = g;
    u64 step = 1;
    while(true)
    {
      u64 nxt = lo + step;
      if(nxt > n - 1) { hi = n - 1; break; }
      if(__ldg(off + nxt) <= t) { lo = nxt; step <<= 1; } else { hi = nxt - 1; break; }
    }
  }
  else
  {
    u64 cur = g, step = 1;
    while(true)
    {
      u64 nxt = (cur >= step ? cur - step : 0);
      if(__ldg(off + nxt) <= t) { lo = nxt; hi = cur - 1; break; }
      cur = nxt; step <<= 1;
    }
  }
  while(lo < hi)
  {
    u64 mid = lo + (hi - lo + 1) / 2;
    if(__ldg(off + mid) <= t) { lo = mid; } else { hi = mid - 1; }
  }
  return lo;
}

#define LOC_DIRECT 0xFFFFFFFFu          // steps marker: `first` holds the value itself (locate table, single-valued node)

/*
  One thread per (range, node): walk LF until a sampled node (locateInternal, gcsa.cpp:882-887),
  remember (first sample, steps) and how many values the node stores (firstSample, gcsa.h:202-206;
  the select on `samples` is an explicit offset array here).
*/
__global__ void __launch_bounds__(256)
locate_walk_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ node_off, u64 n, u64 items,
                   u64* __restrict__ first, u32* __restrict__ steps_out, u64* __restrict__ cnt)
{
  const double ratio = (double)n / (double)items;
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (u64)gridDim.x * blockDim.x)
  {
    // range owning item t: last r with node_off[r] <= t
    u64 lo = owner_of(node_off, n, t, (u64)((double)t * ratio));
    u64 node = sp[lo] + (t - node_off[lo]);
    u32 steps = 0;
    u64 r;
    if(v.loc64 != nullptr)
    {
      u64 e = __ldg(v.loc64 + node);
      if(e >> 63) { first[t] = e & ~(1ull << 63); steps_out[t] = LOC_DIRECT; cnt[t] = 1; continue; }
      r = e >> 24; steps = (u32)(e & 0xFFFFFFu);
    }
    else if(v.walk32 != nullptr)
    {
      u32 e = __ldg(v.walk32 + node);
      while(!(e & 1)) { e = __ldg(v.walk32 + (e >> 1)); steps++; }
      r = e >> 1;
    }
    else if(v.walk64 != nullptr)
    {
      u64 e = __ldg(v.walk64 + node);
      while(!(e & 1)) { e = __ldg(v.walk64 + (e >> 1)); steps++; }
      r = e >> 1;
    }
    else
    {
      while(!rv_get_rank(v.sampled, node, r)) { node = lf_node(v, node); steps++; }
    }
    u64 s0 = v.sample_start[r], s1 = v.sample_start[r + 1];
    first[t] = s0; steps_out[t] = steps; cnt[t] = s1 - s0;
  }
}

__global__ void __launch_bounds__(256)
locate_fill_kernel(const DevView v, u64 items, const u64* __restrict__ first, const u32* __restrict__ steps,
                   const u64* __restrict__ val_off, u64* __restrict__ raw)
{
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (u64)gridDim.x * blockDim.x)
  {
    u64 s0 = first[t], o0 = val_off[t], c = val_off[t + 1] - o0;
    if(steps[t] == LOC_DIRECT) { raw[o0] = s0; continue; }
    for(u64 j = 0; j < c; j++) { raw[o0 + j] = v.stored_samples[s0 + j] + steps[t]; }   // gcsa.cpp:893
  }
}

// segment boundaries of the raw values, per range: seg[r] = val_off[node_off[r]]
__global__ void __launch_bounds__(256)
locate_segments_kernel(const u64* __restrict__ node_off, const u64* __restrict__ val_off, u64 n, u64* __restrict__ seg)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (u64)gridDim.x * blockDim.x)
  {
    seg[i] = val_off[node_off[i]];
  }
}

/*
  Short ranges through the locate table, without the general pipeline: a range of at most LOC_SMALL path nodes
  whose table entries all hold their start position directly (the usual case: a k-mer that occurs a few times)
  is gathered, sorted and deduplicated in registers by one thread -- locate(range) of src/gcsa.cpp:827-842 with
  removeDuplicates (utils.h:350-357) on up to eight values.  Pass 1 counts (and keeps the value of single-valued
  ranges), an exclusive scan gives the CSR offsets, pass 2 writes.  Every other range (longer, or with a node whose
  sampled ancestor stores several positions) is appended to a list and goes through the general pipeline below.
*/
#define LOC_SMALL 8
#define LOC_TOP (1ull << 63)
#define LOC_MED (1ull << 62)            // stash marker of a medium range (below LOC_TOP: start positions are < 2^62)
#define LOC_MED_SHORT 128               // nodes a warp sorts with at most 4 values per thread
#define LOC_MED_WARP 1024               // nodes a warp sorts (32 values per thread)
#define LOC_MED_BLOCK 4096              // nodes a block of 256 threads sorts (16 values per thread)
#define LOC_MED_HUGE 16384              // nodes a block of 512 threads sorts (32 values per thread)

// start positions of the nodes [s, s + len), len <= LOC_SMALL, padded with ~0; false if an entry is not direct
__device__ __forceinline__ bool locate_small_values(const DevView& v, u64 s, u32 len, u64 (&a)[LOC_SMALL])
{
  bool direct = true;
  #pragma unroll
  for(u32 j = 0; j < LOC_SMALL; j++)
  {
    u64 e = (j < len ? __ldg(v.loc64 + s + j) : ~0ull);
    direct = direct && ((e >> 63) != 0);
    a[j] = (j < len ? (e & ~LOC_TOP) : ~0ull);
  }
  return direct;
}

// odd-even transposition network over LOC_SMALL registers (the padding sorts to the end)
__device__ __forceinline__ void locate_small_sort(u64 (&a)[LOC_SMALL])
{
  #pragma unroll
  for(int r = 0; r < LOC_SMALL; r++)
  {
    #pragma unroll
    for(int j = (r & 1); j + 1 < LOC_SMALL; j += 2)
    {
      u64 x = a[j], y = a[j + 1];
      a[j] = (x < y ? x : y); a[j + 1] = (x < y ? y : x);
    }
  }
}

// Pass 1.  cnt[i] = number of distinct positions of range i (0 for the general ranges, which are appended to
// glist); stash[i] = the position itself when there is exactly one, LOC_TOP | list slot for a general range.
// Ranges of LOC_SMALL + 1 .. LOC_MED_HUGE nodes are appended to `mlist` (2 n entries) instead: those of at most
// LOC_MED_SHORT nodes from its front, those of LOC_MED_WARP + 1 .. LOC_MED_BLOCK from the back of its first half, those of
// at most LOC_MED_WARP from the front of its second half, the largest from its back; they reserve their nodes' worth of the
// medium scratch: stash[i] = LOC_MED | offset into the scratch.  counters: [0] general ranges, [1] short, [2] block-sized,
// [3] scratch entries reserved, [4] warp-sized, [5] large-block-sized.
__global__ void __launch_bounds__(256)
locate_small_count_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n,
                          u64* __restrict__ cnt, u64* __restrict__ stash, u64* __restrict__ glist, u64* __restrict__ mlist,
                          ull* __restrict__ counters, bool medium)
{
  // (the lists and the scratch are claimed once per warp: one atomic per counter and warp instead of one per range)
  const u32 lane = threadIdx.x & 31;
  for(u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < n; i0 += (u64)gridDim.x * blockDim.x)
  {
    const u64 i = i0 + threadIdx.x;
    const bool active = (i < n);
    u64 s = (active ? sp[i] : 1), e = (active ? ep[i] : 0);
    u64 c = 0, keep = 0, med_len = 0;
    u32 kind = 0;                                                    // 1 general, 2 medium (short), 3 medium (block), 4 medium (warp), 5 medium (large block)
    if(!(range_empty(s, e) || e >= v.path_nodes))                    // gcsa.cpp:831
    {
      u64 len = e + 1 - s;
      if(len == 1)
      {
        u64 x = __ldg(v.loc64 + s);
        if(x >> 63) { c = 1; keep = x & ~LOC_TOP; } else { kind = 1; }
      }
      else if(len <= LOC_SMALL)
      {
        u64 a[LOC_SMALL];
        if(locate_small_values(v, s, (u32)len, a))
        {
          locate_small_sort(a);
          c = 1;
          #pragma unroll
          for(u32 j = 1; j < LOC_SMALL; j++) { c += ((j < len && a[j] != a[j - 1]) ? 1 : 0); }
          keep = a[0];
        }
        else { kind = 1; }
      }
      else if(medium && len <= LOC_MED_HUGE) { kind = (len <= LOC_MED_SHORT ? 2 : (len <= LOC_MED_WARP ? 4 : (len <= LOC_MED_BLOCK ? 3 : 5))); med_len = len; }
      else { kind = 1; }
    }
    const u32 general = __ballot_sync(0xFFFFFFFFu, kind == 1);
    const u32 by_short = __ballot_sync(0xFFFFFFFFu, kind == 2), by_block = __ballot_sync(0xFFFFFFFFu, kind == 3);
    const u32 by_warp = __ballot_sync(0xFFFFFFFFu, kind == 4), by_huge = __ballot_sync(0xFFFFFFFFu, kind == 5);
    const u32 below = (1u << lane) - 1;
    if(general != 0)
    {
      u64 first = (lane == 0 ? atomicAdd(counters, (ull)__popc(general)) : 0);
      first = __shfl_sync(0xFFFFFFFFu, first, 0);
      if(kind == 1) { u64 slot = first + __popc(general & below); glist[slot] = i; keep = LOC_TOP | slot; }
    }
    if((by_short | by_warp | by_block | by_huge) != 0)
    {
      u64 incl = med_len;
      #pragma unroll
      for(int d = 1; d < 32; d <<= 1) { u64 y = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= (u32)d) { incl += y; } }
      u64 all = __shfl_sync(0xFFFFFFFFu, incl, 31);
      u64 at = (lane == 0 ? atomicAdd(counters + 3, (ull)all) : 0);
      u64 first_s = (lane == 0 && by_short != 0 ? atomicAdd(counters + 1, (ull)__popc(by_short)) : 0);
      u64 first_b = (lane == 0 && by_block != 0 ? atomicAdd(counters + 2, (ull)__popc(by_block)) : 0);
      u64 first_w = (lane == 0 && by_warp != 0 ? atomicAdd(counters + 4, (ull)__popc(by_warp)) : 0);
      u64 first_h = (lane == 0 && by_huge != 0 ? atomicAdd(counters + 5, (ull)__popc(by_huge)) : 0);
      first_h = __shfl_sync(0xFFFFFFFFu, first_h, 0);
      at = __shfl_sync(0xFFFFFFFFu, at, 0); first_s = __shfl_sync(0xFFFFFFFFu, first_s, 0);
      first_b = __shfl_sync(0xFFFFFFFFu, first_b, 0); first_w = __shfl_sync(0xFFFFFFFFu, first_w, 0);
      if(kind == 2) { mlist[first_s + __popc(by_short & below)] = i; }
      if(kind == 3) { mlist[n - 1 - (first_b + __popc(by_block & below))] = i; }
      if(kind == 4) { mlist[n + first_w + __popc(by_warp & below)] = i; }
      if(kind == 5) { mlist[2 * n - 1 - (first_h + __popc(by_huge & below))] = i; }
      if(kind >= 2) { keep = LOC_MED | (at + incl - med_len); }
    }
    if(active) { cnt[i] = c; stash[i] = keep; }
  }
}

// the general ranges, in list order
__global__ void __launch_bounds__(256)
locate_general_gather_kernel(const u64* __restrict__ sp, const u64* __restrict__ ep, const u64* __restrict__ glist, u64 m,
                             u64* __restrict__ gsp, u64* __restrict__ gep)
{
  for(u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (u64)gridDim.x * blockDim.x)
  {
    u64 i = glist[k];
    gsp[k] = sp[i]; gep[k] = ep[i];
  }
}

// their counts, once the general pipeline has answered
__global__ void __launch_bounds__(256)
locate_general_counts_kernel(const u64* __restrict__ glist, const u64* __restrict__ goffs, u64 m, u64* __restrict__ cnt)
{
  for(u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (u64)gridDim.x * blockDim.x)
  {
    cnt[glist[k]] = goffs[k + 1] - goffs[k];
  }
}

// Pass 2: values[off[i], off[i + 1]) of every range.
__global__ void __launch_bounds__(256)
locate_small_fill_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n,
                         const u64* __restrict__ off, const u64* __restrict__ stash, u64* __restrict__ values)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 o = off[i], c = off[i + 1] - o;
    if(c == 0) { continue; }
    u64 keep = stash[i];
    if(keep >> 62) { continue; }                         // medium or general: locate_copy_kernel
    else if(c == 1) { values[o] = keep; }
    else
    {
      u64 s = sp[i], len = ep[i] + 1 - s;
      u64 a[LOC_SMALL];
      locate_small_values(v, s, (u32)len, a);
      locate_small_sort(a);
      values[o] = a[0];
      u64 w = 1;
      #pragma unroll
      for(u32 j = 1; j < LOC_SMALL; j++)
      {
        if(j < len && a[j] != a[j - 1]) { values[o + w] = a[j]; w++; }
      }
    }
  }
}

/*
  Medium ranges through the locate table: one group of threads (a warp, or a block of 256) per range.  The group
  loads the range's table entries (consecutive nodes: coalesced), and if all of them are direct sorts the start
  positions, drops the duplicates (removeDuplicates, utils.h:350-357) and writes the distinct positions to the
  range's place in the scratch; cnt[i] = their number.  A range with a node whose sampled ancestor stores several
  positions is handed to the general pipeline after all (appended to glist, its stash entry rewritten).
  locate(range) of src/gcsa.cpp:827-842 for the ranges of a short pattern: tens to thousands of occurrences, where the
  segmented sort of the general pipeline spends most of its time on bookkeeping.

  The sort is a bitonic network over E values per thread held in REGISTERS (E = 1 .. 32, a power of two; the sequence
  is padded with ~0): thread t owns elements t * E .. t * E + E - 1, so a compare-exchange at distance j is register
  to register for j < E, a warp shuffle for E <= j < 32 E and an exchange through shared memory only between the
  warps of a block (j >= 32 E).  Of the 55 stages of a 1024-value sort by one warp 40 are register-only and 15 are
  shuffles; the same network run out of shared memory was bound by its bandwidth (profiles/r02_locate_medium.txt).
*/
// one stage: compare-exchange at distance J inside bitonic runs of length KK (both compile-time, so that the values
// never leave the registers)
template<int E, int WARPS, u32 KK, u32 J>
__device__ __forceinline__ void locate_bitonic_stage(u64 (&a)[E], u32 tid, u64* exchange)
{
  constexpr u32 T = 32u * WARPS;                         // threads of the group
  if constexpr (J >= 32u * E)
  {
    // partner in another warp: through shared memory, element r of thread t at [r * T + t] (conflict-free)
    // (LOC_MED_BLOCK values at a time: the 512-thread class holds four times that)
    constexpr u32 tj = J / E;
    constexpr int R = ((u32)E * T > LOC_MED_BLOCK ? (int)(LOC_MED_BLOCK / T) : E);
    const bool keep_min = (((tid & tj) == 0) == ((tid & (KK / E)) == 0));
    #pragma unroll
    for(int r0 = 0; r0 < E; r0 += R)
    {
      __syncthreads();
      #pragma unroll
      for(int r = 0; r < R; r++) { exchange[r * T + tid] = a[r0 + r]; }
      __syncthreads();
      #pragma unroll
      for(int r = 0; r < R; r++)
      {
        u64 other = exchange[r * T + (tid ^ tj)];
        a[r0 + r] = (keep_min ? (a[r0 + r] < other ? a[r0 + r] : other) : (a[r0 + r] < other ? other : a[r0 + r]));
      }
    }
  }
  else if constexpr (J >= (u32)E)
  {
    constexpr u32 tj = J / E;
    const bool keep_min = (((tid & tj) == 0) == ((tid & (KK / E)) == 0));
    #pragma unroll
    for(int r = 0; r < E; r++)
    {
      u64 other = __shfl_xor_sync(0xFFFFFFFFu, a[r], tj);
      a[r] = (keep_min ? (a[r] < other ? a[r] : other) : (a[r] < other ? other : a[r]));
    }
  }
  else
  {
    #pragma unroll
    for(int r = 0; r < E; r++)
    {
      if((r & J) == 0)
      {
        const bool up = (KK < (u32)E ? ((r & KK) == 0) : ((tid & (KK / E)) == 0));
        u64 x = a[r], y = a[r | J];
        bool swap = ((x > y) == up);
        a[r] = (swap ? y : x); a[r | J] = (swap ? x : y);
      }
    }
  }
}

template<int E, int WARPS, u32 KK, u32 J>
__device__ __forceinline__ void locate_bitonic_merge(u64 (&a)[E], u32 tid, u64* exchange)
{
  locate_bitonic_stage<E, WARPS, KK, J>(a, tid, exchange);
  if constexpr (J > 1) { locate_bitonic_merge<E, WARPS, KK, J / 2>(a, tid, exchange); }
}

template<int E, int WARPS, u32 KK = 2>
__device__ __forceinline__ void locate_bitonic_sort(u64 (&a)[E], u32 tid, u64* exchange)
{
  locate_bitonic_merge<E, WARPS, KK, KK / 2>(a, tid, exchange);
  if constexpr (KK < 32u * WARPS * E) { locate_bitonic_sort<E, WARPS, KK * 2>(a, tid, exchange); }
}

// One range: `len` nodes from s, len <= 32 * WARPS * E.  Returns the number of distinct positions written to `out`
// (valid in every thread), or ~0 when an entry of the range is not direct.
template<int E, int WARPS>
__device__ __forceinline__ u32 locate_medium_range(const DevView& v, u64 s, u32 len, u32 tid, u64* __restrict__ out, u64* exchange, u32* warp_total)
{
  constexpr u32 T = 32u * WARPS;
  const u32 lane = tid & 31;
  u64 a[E];
  bool indirect = false;
  #pragma unroll
  for(int r = 0; r < E; r++)
  {
    u32 idx = (u32)r * T + tid;                         // coalesced; which thread sorts which value does not matter
    u64 e = (idx < len ? __ldg(v.loc64 + s + idx) : ~0ull);
    indirect = indirect || ((e >> 63) == 0);
    a[r] = (e == ~0ull ? ~0ull : (e & ~LOC_TOP));       // (a direct entry is LOC_TOP | position: never all ones)
  }
  if(WARPS == 1) { indirect = (__any_sync(0xFFFFFFFFu, indirect) != 0); }
  else
  {
    __syncthreads();
    if(tid == 0) { warp_total[WARPS] = 0; }
    __syncthreads();
    if(indirect) { warp_total[WARPS] = 1; }
    __syncthreads();
    indirect = (warp_total[WARPS] != 0);
  }
  if(indirect) { return ~0u; }

  locate_bitonic_sort<E, WARPS>(a, tid, exchange);

  // first copies: element 0 of a thread is compared with the last element of the thread before it
  u64 before = __shfl_up_sync(0xFFFFFFFFu, a[E - 1], 1);
  if(WARPS > 1)
  {
    __syncthreads();
    if(lane == 31) { exchange[tid / 32] = a[E - 1]; }
    __syncthreads();
    if(lane == 0 && tid > 0) { before = exchange[tid / 32 - 1]; }
  }
  u32 mine = 0;
  #pragma unroll
  for(int r = 0; r < E; r++)
  {
    bool first = (a[r] != ~0ull) && (r == 0 ? (tid == 0 || a[0] != before) : (a[r] != a[r > 0 ? r - 1 : 0]));
    mine += (first ? 1u : 0u);
  }
  u32 incl = mine;
  #pragma unroll
  for(int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, incl, d); if(lane >= (u32)d) { incl += y; } }
  u32 pos = incl - mine, total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  if(WARPS > 1)
  {
    __syncthreads();
    if(lane == 31) { warp_total[tid / 32] = incl; }
    __syncthreads();
    total = 0;
    #pragma unroll
    for(int w = 0; w < WARPS; w++) { u32 x = warp_total[w]; total += x; if(w < (int)(tid / 32)) { pos += x; } }
  }
  #pragma unroll
  for(int r = 0; r < E; r++)
  {
    bool first = (a[r] != ~0ull) && (r == 0 ? (tid == 0 || a[0] != before) : (a[r] != a[r > 0 ? r - 1 : 0]));
    if(first) { out[pos] = a[r]; pos++; }
  }
  return total;
}

// CLASS 0: a warp per range of at most LOC_MED_SHORT nodes (few registers, many warps in flight: these are bound by
// the latency of their loads); CLASS 1: a warp per range of at most LOC_MED_WARP nodes (up to 32 values per thread;
// bound by the integer pipe); CLASS 2: a block of 256 threads per range of at most LOC_MED_BLOCK nodes;
// CLASS 3: a block of 512 threads per range of at most LOC_MED_HUGE nodes (32 values per thread).
template<int CLASS>
__global__ void __launch_bounds__(CLASS == 3 ? 512 : (CLASS == 2 ? 256 : 128))
locate_medium_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, const u64* __restrict__ mlist, u64 m,
                     u64* __restrict__ cnt, u64* __restrict__ stash, u64* __restrict__ scratch,
                     u64* __restrict__ glist, ull* __restrict__ n_general)
{
  constexpr int WARPS = (CLASS == 3 ? 16 : (CLASS == 2 ? 8 : 1));
  __shared__ u64 exchange[WARPS == 1 ? 1 : LOC_MED_BLOCK];
  __shared__ u32 warp_total[WARPS + 1];
  const u32 tid = (WARPS == 1 ? threadIdx.x & 31 : threadIdx.x);
  const u64 group = (WARPS == 1 ? ((u64)blockIdx.x * blockDim.x + threadIdx.x) / 32 : (u64)blockIdx.x);
  const u64 n_groups = (WARPS == 1 ? ((u64)gridDim.x * blockDim.x) / 32 : (u64)gridDim.x);
  for(u64 k = group; k < m; k += n_groups)
  {
    const u64 i = mlist[k];
    const u64 s = sp[i];
    const u32 len = (u32)(ep[i] + 1 - s);
    u64* out = scratch + (stash[i] & ~(LOC_TOP | LOC_MED));
    const u32 per = (len + 32 * WARPS - 1) / (32 * WARPS);
    u32 c;
    if(CLASS == 0)
    {
      if(per <= 1) { c = locate_medium_range<1, WARPS>(v, s, len, tid, out, exchange, warp_total); }
      else if(per <= 2) { c = locate_medium_range<2, WARPS>(v, s, len, tid, out, exchange, warp_total); }
      else { c = locate_medium_range<4, WARPS>(v, s, len, tid, out, exchange, warp_total); }
    }
    else if(CLASS == 1)
    {
      if(per <= 8) { c = locate_medium_range<8, WARPS>(v, s, len, tid, out, exchange, warp_total); }
      else if(per <= 16) { c = locate_medium_range<16, WARPS>(v, s, len, tid, out, exchange, warp_total); }
      else { c = locate_medium_range<32, WARPS>(v, s, len, tid, out, exchange, warp_total); }
    }
    else if(CLASS == 2)
    {
      if(per <= 8) { c = locate_medium_range<8, WARPS>(v, s, len, tid, out, exchange, warp_total); }
      else { c = locate_medium_range<16, WARPS>(v, s, len, tid, out, exchange, warp_total); }
    }
    else
    {
      if(per <= 16) { c = locate_medium_range<16, WARPS>(v, s, len, tid, out, exchange, warp_total); }
      else { c = locate_medium_range<32, WARPS>(v, s, len, tid, out, exchange, warp_total); }
    }
    if(tid == 0)
    {
      if(c == ~0u)
      {
        u64 slot = atomicAdd(n_general, 1ull);
        glist[slot] = i; stash[i] = LOC_TOP | slot; cnt[i] = 0;
      }
      else { cnt[i] = c; }
    }
  }
}

// values of the medium (from the scratch) or general (from the general pipeline's CSR) ranges into their final place,
// one warp per range
template<bool GENERAL>
__global__ void __launch_bounds__(256)
locate_copy_kernel(const u64* __restrict__ list, u64 m, const u64* __restrict__ stash, const u64* __restrict__ off,
                   const u64* __restrict__ src_offs, const u64* __restrict__ src, u64* __restrict__ values)
{
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for(u64 k = warp; k < m; k += n_warps)
  {
    u64 i = list[k], keep = stash[i];
    if(!GENERAL && (keep >> 63)) { continue; }             // handed to the general pipeline by the medium kernel
    u64 o = off[i], c = off[i + 1] - o;
    const u64* from = (GENERAL ? src + src_offs[k] : src + (keep & ~(LOC_TOP | LOC_MED)));
    for(u64 j = lane; j < c; j += 32) { values[o + j] = from[j]; }
  }
}

// removeDuplicates (utils.h:350-357) after the segmented sort: flag the first copy of each value
__global__ void __launch_bounds__(256)
locate_flag_kernel(const u64* __restrict__ sorted, const u64* __restrict__ seg, u64 n, u64 total, u64* __restrict__ flag)
{
  const double ratio = (double)n / (double)total;
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    // segment of t: last r with seg[r] <= t
    u64 lo = owner_of(seg, n, t, (u64)((double)t * ratio));
    flag[t] = (t == seg[lo] || sorted[t] != sorted[t - 1]) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256)
locate_compact_kernel(const u64* __restrict__ sorted, const u64* __restrict__ flag, const u64* __restrict__ flag_scan,
                      u64 total, u64* __restrict__ values, u64 capacity)
{
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    if(flag[t] && flag_scan[t] < capacity) { values[flag_scan[t]] = sorted[t]; }
  }
}

__global__ void __launch_bounds__(256)
locate_offsets_kernel(const u64* __restrict__ seg, const u64* __restrict__ flag_scan, u64 n, u64 total, u64 distinct,
                      u64* __restrict__ out_offsets)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 s = seg[i];
    out_offsets[i] = (s >= total ? distinct : flag_scan[s]);
  }
}

#endif
