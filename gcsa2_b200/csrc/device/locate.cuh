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
__global__ void __launch_bounds__(256)
locate_small_count_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n,
                          u64* __restrict__ cnt, u64* __restrict__ stash, u64* __restrict__ glist, ull* __restrict__ n_general)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 s = sp[i], e = ep[i];
    u64 c = 0, keep = 0;
    bool general = false;
    if(!(range_empty(s, e) || e >= v.path_nodes))                    // gcsa.cpp:831
    {
      u64 len = e + 1 - s;
      if(len == 1)
      {
        u64 x = __ldg(v.loc64 + s);
        if(x >> 63) { c = 1; keep = x & ~LOC_TOP; } else { general = true; }
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
        else { general = true; }
      }
      else { general = true; }
    }
    if(general)
    {
      u64 slot = atomicAdd(n_general, 1ull);
      glist[slot] = i;
      keep = LOC_TOP | slot;
    }
    cnt[i] = c; stash[i] = keep;
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
                         const u64* __restrict__ off, const u64* __restrict__ stash,
                         const u64* __restrict__ goffs, const u64* __restrict__ gvals, u64* __restrict__ values)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 o = off[i], c = off[i + 1] - o;
    if(c == 0) { continue; }
    u64 keep = stash[i];
    if(keep >> 63)
    {
      u64 g = goffs[keep & ~LOC_TOP];
      for(u64 j = 0; j < c; j++) { values[o + j] = gvals[g + j]; }
    }
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
