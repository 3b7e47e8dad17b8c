/* device/lcp.cuh -- LCPArray: parent, depth, psv / nsv / rmq over the k-ary minimum tree.
   Included by lcp.cu only; sm_100a only. */
#ifndef GCSA2_B200_DEVICE_LCP_CUH
#define GCSA2_B200_DEVICE_LCP_CUH

//------------------------------------------------------------------------------
// Kernels: LCP (src/lcp.cpp:152-200, 276-519)
//------------------------------------------------------------------------------

struct Pair64 { u64 first, second; };

__device__ __forceinline__ u64 rmt_parent(const LcpView& l, u64 node, u64 level)
{
  u64 rel = node - l.offsets[level];
  return l.offsets[level + 1] + (l.shift >= 0 ? rel >> l.shift : rel / l.branching);
}
__device__ __forceinline__ u64 rmt_first_sibling(const LcpView& l, u64 node, u64 level)
{
  u64 rel = node - l.offsets[level];
  return node - (l.shift >= 0 ? rel & (l.branching - 1) : rel % l.branching);
}
__device__ __forceinline__ u64 rmt_last_sibling(const LcpView& l, u64 first_child, u64 level)
{ u64 a = l.offsets[level + 1], b = first_child + l.branching; return (a < b ? a : b) - 1; }
__device__ __forceinline__ u64 rmt_first_child(const LcpView& l, u64 node, u64 level) { return l.offsets[level - 1] + (node - l.offsets[level]) * l.branching; }
__device__ __forceinline__ u64 rmt_last_child(const LcpView& l, u64 node, u64 level) { return rmt_last_sibling(l, rmt_first_child(l, node, level), level - 1); }
__device__ __forceinline__ u64 rmt_level(const LcpView& l, u64 node) { u64 level = 0; while(l.offsets[level + 1] <= node) { level++; } return level; }

template<bool OR_EQUAL> __device__ __forceinline__ bool sv_less(u64 a, u64 b) { return (OR_EQUAL ? a <= b : a < b); }

// The sibling scans of psv / nsv (lcp.cpp:354-367, 410-423) read the one-byte values eight at a time:
// 0x80 in every byte of x that is < thr (1 <= thr <= 256).
__device__ __forceinline__ u64 bytes_below(u64 x, u32 thr)
{
  if(thr >= 256) { return 0x8080808080808080ull; }
  u32 t = thr * 0x01010101u;
  u32 lo = __vcmpltu4((u32)x, t), hi = __vcmpltu4((u32)(x >> 32), t);
  return (((u64)hi << 32) | lo) & 0x8080808080808080ull;
}

// first / last index in [a, b] (a <= b) whose value is < thr; ~0 if there is none
__device__ __forceinline__ u64 scan_up(const u8* __restrict__ data, u64 a, u64 b, u32 thr)
{
  if(thr == 0) { return ~0ull; }
  const u64 w0 = a >> 3, w1 = b >> 3;
  for(u64 w = w0; w <= w1; w++)
  {
    u64 m = bytes_below(__ldg((const unsigned long long*)data + w), thr);
    if(w == w0) { m &= ~0ull << ((a & 7) * 8); }
    if(w == w1) { m &= ~0ull >> ((7 - (b & 7)) * 8); }
    if(m) { return w * 8 + ((u64)(__ffsll((long long)m) - 1) >> 3); }
  }
  return ~0ull;
}

__device__ __forceinline__ u64 scan_down(const u8* __restrict__ data, u64 a, u64 b, u32 thr)
{
  if(thr == 0) { return ~0ull; }
  const u64 w0 = a >> 3, w1 = b >> 3;
  for(u64 w = w1; ; w--)
  {
    u64 m = bytes_below(__ldg((const unsigned long long*)data + w), thr);
    if(w == w0) { m &= ~0ull << ((a & 7) * 8); }
    if(w == w1) { m &= ~0ull >> ((7 - (b & 7)) * 8); }
    if(m) { return w * 8 + ((u64)(63 - __clzll((long long)m)) >> 3); }
    if(w == w0) { break; }
  }
  return ~0ull;
}

// lcp.cpp:333-370
template<bool OR_EQUAL>
__device__ Pair64 lcp_psv(const LcpView& l, u64 to)
{
  Pair64 nf = { l.values, l.values };
  if(to == 0 || to >= l.size) { return nf; }
  u64 level = 0;
  const u32 thr = (u32)l.data[to] + (OR_EQUAL ? 1 : 0);
  u64 found = ~0ull;
  while(to != l.values - 1)
  {
    u64 from = rmt_first_sibling(l, to, level);
    found = (to > from ? scan_down(l.data, from, to - 1, thr) : ~0ull);
    if(found != ~0ull) { break; }
    to = rmt_parent(l, to, level); level++;
  }
  if(found == ~0ull) { return nf; }
  while(level > 0)
  {
    u64 from = rmt_first_child(l, found, level); level--;
    found = scan_down(l.data, from, rmt_last_sibling(l, from, level), thr);
  }
  Pair64 res = { found, l.data[found] };
  return res;
}

// lcp.cpp:389-426
template<bool OR_EQUAL>
__device__ Pair64 lcp_nsv(const LcpView& l, u64 from)
{
  Pair64 nf = { l.values, l.values };
  if(from + 1 >= l.size) { return nf; }
  u64 level = 0;
  const u32 thr = (u32)l.data[from] + (OR_EQUAL ? 1 : 0);
  u64 found = ~0ull;
  while(from != l.values - 1)
  {
    u64 to = rmt_last_sibling(l, from, level);
    found = (from + 1 <= to ? scan_up(l.data, from + 1, to, thr) : ~0ull);
    if(found != ~0ull) { break; }
    from = rmt_parent(l, from, level); level++;
  }
  if(found == ~0ull) { return nf; }
  while(level > 0)
  {
    from = rmt_first_child(l, found, level); level--;
    found = scan_up(l.data, from, rmt_last_sibling(l, from, level), thr);
  }
  Pair64 res = { found, l.data[found] };
  return res;
}

/*
  lcp.cpp:448-513 rmq(sp, ep): leftmost minimum.  The reference collects the right-hand partial
  sibling groups on a stack and pops them afterwards so that positions are visited left to right;
  here the right-hand side keeps its own running minimum with "<=" (a later, more-left group wins
  ties), which yields the same leftmost minimum without a stack.
*/
__device__ Pair64 lcp_rmq(const LcpView& l, u64 sp, u64 ep)
{
  Pair64 nf = { l.values, l.values };
  if(sp > ep || ep >= l.size) { return nf; }
  if(sp == ep) { Pair64 r = { sp, l.data[sp] }; return r; }

  Pair64 res = { l.values, l.size };
  Pair64 tail = { l.values, ~0ull };
  u64 level = 0, left = sp, right = ep;
  while(true)
  {
    u64 left_par = rmt_parent(l, left, level), right_par = rmt_parent(l, right, level);
    if(left_par == right_par)
    {
      for(u64 i = left; i <= right; i++) { u64 x = l.data[i]; if(x < res.second) { res.first = i; res.second = x; } }
      break;
    }
    u64 left_child = rmt_first_child(l, left_par, level + 1);
    if(left != left_child)
    {
      u64 last_child = rmt_last_sibling(l, left_child, level);
      for(u64 i = left; i <= last_child; i++) { u64 x = l.data[i]; if(x < res.second) { res.first = i; res.second = x; } }
      left_par++;
    }
    u64 right_child = rmt_last_child(l, right_par, level + 1);
    if(right != right_child)
    {
      u64 first_child = rmt_first_sibling(l, right_child, level);
      // this group lies to the LEFT of everything already in tail: it wins ties; inside the group
      // the leftmost minimum wins.
      Pair64 grp = { l.values, ~0ull };
      for(u64 i = first_child; i <= right; i++) { u64 x = l.data[i]; if(x < grp.second) { grp.first = i; grp.second = x; } }
      if(grp.second <= tail.second) { tail = grp; }
      right_par--;
    }
    if(left_par >= right_par)
    {
      if(left_par == right_par) { u64 x = l.data[left_par]; if(x < res.second) { res.first = left_par; res.second = x; } }
      break;
    }
    left = left_par; right = right_par; level++;
  }
  if(tail.first < l.values && tail.second < res.second) { res = tail; }
  if(res.first >= l.values) { return res; }

  level = rmt_level(l, res.first);
  while(level > 0)
  {
    res.first = rmt_first_child(l, res.first, level); level--;
    while(l.data[res.first] != res.second) { res.first++; }
  }
  return res;
}

// LCPArray::parent(range), lcp.cpp:276-301 with nodeFor (lcp.h:163-175) and root (lcp.h:137)
__device__ gcsa_b200_stnode lcp_parent(const LcpView& l, u64 sp, u64 ep)
{
  gcsa_b200_stnode out;
  // a start outside the array (the reference would read past its vector): the root, like a range that covers everything
  if(sp >= l.size) { out.sp = 0; out.ep = l.size - 1; out.left_lcp = 0; out.right_lcp = 0; out.node_lcp = 0; return out; }
  if(sp == 0 && ep == l.size - 1) { out.sp = 0; out.ep = l.size - 1; out.left_lcp = 0; out.right_lcp = 0; out.node_lcp = 0; return out; }
  u64 left_lcp = l.data[sp];
  u64 right_lcp = (ep + 1 < l.size ? l.data[ep + 1] : 0);
  u64 node_lcp = (left_lcp > right_lcp ? left_lcp : right_lcp);
  Pair64 left = { sp, left_lcp }, right = { ep + 1, right_lcp };
  if(left_lcp == node_lcp)
  {
    left = lcp_psv<false>(l, sp);
    if(left.first == l.values && left.second == l.values) { left.first = 0; left.second = 0; }
  }
  if(right_lcp == node_lcp)
  {
    right = lcp_nsv<false>(l, ep + 1);
    if(right.first == l.values && right.second == l.values) { right.first = l.size; right.second = 0; }
  }
  out.sp = left.first; out.ep = right.first - 1; out.left_lcp = left.second; out.right_lcp = right.second; out.node_lcp = node_lcp;
  return out;
}

__global__ void __launch_bounds__(256)
parent_kernel(const LcpView l, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, gcsa_b200_stnode* __restrict__ out)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    out[i] = lcp_parent(l, sp[i], ep[i]);
  }
}

__global__ void __launch_bounds__(256)
depth_kernel(const LcpView l, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, u64* __restrict__ out)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 s = sp[i], e = ep[i];
    u64 res = GCSA_B200_UNKNOWN;
    if(e + 1 - s > 1)                                                  // lcp.cpp:321
    {
      Pair64 r = lcp_rmq(l, s + 1, e);
      if(!(r.first == l.values && r.second == l.values)) { res = r.second; }
    }
    out[i] = res;
  }
}

__global__ void __launch_bounds__(256)
lcp_sv_kernel(const LcpView l, int which, const u64* __restrict__ pos, u64 n, u64* __restrict__ opos, u64* __restrict__ oval)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    Pair64 r;
    if(which == 0) { r = lcp_psv<false>(l, pos[i]); }
    else if(which == 1) { r = lcp_psv<true>(l, pos[i]); }
    else if(which == 2) { r = lcp_nsv<false>(l, pos[i]); }
    else { r = lcp_nsv<true>(l, pos[i]); }
    opos[i] = r.first; oval[i] = r.second;
  }
}

__global__ void __launch_bounds__(256)
lcp_rmq_kernel(const LcpView l, const u64* __restrict__ sp, const u64* __restrict__ ep, u64 n, u64* __restrict__ opos, u64* __restrict__ oval)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    Pair64 r = lcp_rmq(l, sp[i], ep[i]);
    opos[i] = r.first; oval[i] = r.second;
  }
}

#endif
