/* device/lf_count.cuh -- LF(range, c), LF(node), LF_fast / LF_all and count() kernels.
   Included by ops.cu only; sm_100a only. */
#ifndef GCSA2_B200_DEVICE_LF_COUNT_CUH
#define GCSA2_B200_DEVICE_LF_COUNT_CUH

//------------------------------------------------------------------------------
// Kernels: LF, count
//------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
lf_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, const u8* __restrict__ comp,
          u64 n, u64* __restrict__ osp, u64* __restrict__ oep)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    // The reference does no sanity checks here (gcsa.h:132-135) and a rank past the end of a vector is undefined
    // there; on the device it would be a read outside the index that poisons the whole context, so positions past
    // the end are clamped to the end (rank(size) is defined: the number of ones).
    u64 s = sp[i], e = ep[i], a, b;
    if(s > v.path_nodes) { s = v.path_nodes; }
    if(e != ~0ull && e >= v.path_nodes) { e = v.path_nodes - 1; }
    lf_range(v, s, e, comp[i], a, b);
    osp[i] = a; oep[i] = b;
  }
}

__global__ void __launch_bounds__(256)
lf_node_kernel(const DevView v, const u64* __restrict__ nodes, u64 n, u64* __restrict__ out)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 node = nodes[i];
    out[i] = (node < v.path_nodes ? lf_node(v, node) : ~0ull);      // no such node (the reference reads past its vectors here)
  }
}

// GCSA::LF_fast / LF_all, src/gcsa.cpp:742-798.  One thread per (range, comp).
__global__ void __launch_bounds__(256)
lf_multi_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, int all_chars,
                u64* __restrict__ out)
{
  u64 total = n * GCSA_B200_SIGMA;
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    u64 i = t / GCSA_B200_SIGMA; u32 c = (u32)(t - i * GCSA_B200_SIGMA);
    u64 a = 1, b = 0;                                        // Range::empty_range()
    u32 last = (all_chars ? GCSA_B200_SIGMA - 2 : GCSA_B200_FAST_CHARS);
    u64 s = sp[i], e = ep[i];
    if(e != ~0ull && e >= v.path_nodes) { e = v.path_nodes - 1; }     // clamped like in lf_kernel
    if(c >= 1 && c <= last && !range_empty(s, e) && s < v.path_nodes)
    {
      if(s == e)                                             // single path node: follow set bits only
      {
        if(c <= GCSA_B200_FAST_CHARS) { u64 p; if(pred_fast(v, s, c - 1, p)) { a = b = p; } }   // bit and step from one sector
        else if(bwt_bit(v, s, c)) { lf_range(v, s, e, c, a, b); }
      }
      else { lf_range(v, s, e, c, a, b); }
    }
    out[t * 2] = a; out[t * 2 + 1] = b;
  }
}

// SadaSparse::count, support.h:329-335
__device__ __forceinline__ u64 sada_sparse_count(const DevView& v, u64 sp, u64 ep)
{
  u64 a = rv_rank(v.extra_filter, sp), b = rv_rank(v.extra_filter, ep + 1);
  if(b <= a) { return 0; }
  return (sv_select(v.extra_values, b) + 1) - (a > 0 ? sv_select(v.extra_values, a) + 1 : 0);
}

// SadaCount::count, support.h:255-258
__device__ __forceinline__ u64 sada_count(const DevView& v, u64 sp, u64 ep)
{
  return (sv_select(v.redundant, ep + 1) - ep) - (sp > 0 ? sv_select(v.redundant, sp) + 1 - sp : 0);
}

// GCSA::count, src/gcsa.cpp:802-809
__device__ __forceinline__ u64 count_range(const DevView& v, u64 sp, u64 ep)
{
  if(range_empty(sp, ep) || ep >= v.path_nodes) { return 0; }
  u64 res = sada_sparse_count(v, sp, ep) + (ep + 1 - sp);
  if(ep > sp) { res -= sada_count(v, sp, ep - 1); }
  return res;
}

__global__ void __launch_bounds__(256)
count_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, u64* __restrict__ out)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    out[i] = count_range(v, sp[i], ep[i]);
  }
}

#endif
