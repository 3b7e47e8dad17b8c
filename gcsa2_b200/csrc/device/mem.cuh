/* device/mem.cuh -- MEM-style scan (LF + parent), BASELINE.json configs[4].
   Included by lcp.cu only; sm_100a only. */
#ifndef GCSA2_B200_DEVICE_MEM_CUH
#define GCSA2_B200_DEVICE_MEM_CUH

//------------------------------------------------------------------------------
// Kernel: MEM-style scan (LF + parent), BASELINE.json configs[4]
//------------------------------------------------------------------------------

/*
  The driver loop over GCSA::LF (gcsa.h:155-162) and LCPArray::parent (lcp.cpp:276-301): extend the
  match to the left while possible; when it cannot be extended, report it (if it grew since the last
  report) and shorten it from the right by moving to the suffix-tree parent.  One pattern per lane;
  patterns of very different lengths share a warp, so finished lanes are refilled from the warp's
  slice exactly as in find_kernel.  WRITE = false counts the matches, WRITE = true stores them at
  the offsets computed from the counts.
*/
// MODE 0: count the matches of each pattern.  MODE 1: write them at out_offsets (exact positions known).
// MODE 2: count AND write the first matches of pattern q into its scratch slot (one pass; the few patterns with more
// matches than their slot holds are redone in MODE 1 over the id list `ids`).  A slot grows with the pattern: pattern q of
// length len, starting at character `begin`, owns (len >> shift) + base entries from (begin >> shift) + q * base
// (mem_slot_start / mem_slot_size; shift = 63 gives every pattern `base` entries).  The number of matches grows with the
// length -- one per mismatch, roughly -- so equal slots either overflow for the long patterns, whose second pass was a third
// of the time of configs[4] (profiles/r02_mem_scan_variants.txt), or waste memory on the short ones.
// JUMP: singleton ranges advance along the unary backward path of their node with one load (the jump tables of
// find_kernel): the pattern is kept 2-bit packed, 32 characters at a time, and a path of up to 16 steps is one XOR
// against it.  A path that the pattern leaves after t characters is followed by t + 1 single steps (the last of
// which fails, as it must), so matches, depths and ranges are those of the single-step loop.
// PACK: the pattern is kept 2-bit packed as for JUMP (one 8-byte streaming load per 8 characters instead of a byte
// load and a table lookup in front of every step), without the jump-table probes.  Measured slower than the byte
// loads (34.2 vs 27.6 ms per 4 M patterns, profiles/r02_mem_scan_variants.txt): the bytes hit the L1, the packing costs
// registers and instructions in a kernel that is short of both; opt-in (GCSA_B200_MEM_PACK=1).
__device__ __forceinline__ u64 mem_slot_start(u64 begin, u64 q, u64 base, u32 shift) { return (begin >> shift) + q * base; }
__device__ __forceinline__ u64 mem_slot_size(u64 len, u64 base, u32 shift) { return (len >> shift) + base; }

// A match record (start, length, sp, ep)
__device__ __forceinline__ void store_match(u64* m, u64 start, u64 length, u64 sp, u64 ep)
{
  m[0] = start; m[1] = length; m[2] = sp; m[3] = ep;
}

template<int MODE, bool JUMP = false, bool PACK = false, int MIN_BLOCKS = 4>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
mem_kernel(const DevView v, const LcpView l, const u8* __restrict__ chars, const u64* __restrict__ offsets, u64 char_base,
           u64 n, u64* __restrict__ counts, const u64* __restrict__ out_offsets, u64* __restrict__ matches,
           const u64* __restrict__ ids, u64 stride, u32 parent_batch, u32 shift)
{
  constexpr bool WRITE = (MODE == 1);
  constexpr bool TAIL = (JUMP || PACK);
  __shared__ u8 c2c[256];
  for(int i = threadIdx.x; i < 256; i += blockDim.x) { c2c[i] = v.char2comp[i]; }
  __syncthreads();

  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u64 per = (n + n_warps - 1) / n_warps;
  u64 next = warp * per;
  const u64 slice_end = (next + per < n ? next + per : n);
  if(next >= n || v.path_nodes == 0) { return; }

  u64 q = 0, sp = 0, ep = 0, depth = 0, pos = 0, begin = 0, emitted = 0, out_at = 0;
  bool live = false, extended = false, need_parent = false;
  u64 tail = 0, tail_end = 0, next_pack = 0; u32 tail_n = 0, skip = 0;     // TAIL: characters [tail_end - tail_n, tail_end) packed as in find_kernel

  // pack the (up to) 32 characters that end at `end_pos` (exclusive), eight at a time, stopping at a non-base
  auto pack_tail = [&](u64 end_pos)
  {
    tail = 0; tail_n = 0; tail_end = end_pos;
    for(u32 w = 0; w < 4; w++)
    {
      u64 pe = end_pos - 8 * w;
      if(pe - begin < 8) { break; }
      u64 addr = (u64)(chars + pe - 8); u32 a = (u32)(addr & 7);
      const unsigned long long* base = (const unsigned long long*)(addr - a);
      u64 word = __ldcs(base);
      if(a != 0) { word = (word >> (8 * a)) | ((u64)__ldcs(base + 1) << (64 - 8 * a)); }
      u32 good;
      u32 r = pack8_reversed(word, &good);
      tail |= (u64)r << (16 * w);
      tail_n += good;
      if(good < 8) { break; }
    }
  };
  auto comp_at = [&](u64 p) -> u32
  {
    if(TAIL)
    {
      u64 off = tail_end - 1 - p;
      if(off < (u64)tail_n) { return (u32)((tail >> (2 * off)) & 3) + 1; }
    }
    return c2c[chars[p]];      // (plain loads: the 32-byte sector stays in the L1 for the following steps; evict-first loads measured 13 % slower)
  };

  while(true)
  {
    u32 dead = __ballot_sync(0xFFFFFFFFu, !live);
    if(dead)
    {
      u32 my = __popc(dead & ((1u << lane) - 1));
      if(!live)
      {
        u64 cand = next + my;
        if(cand < slice_end)
        {
          q = (ids != nullptr ? ids[cand] : cand); live = true; need_parent = false;
          begin = offsets[q] - char_base; pos = offsets[q + 1] - char_base;
          sp = 0; ep = v.path_nodes - 1; depth = 0; extended = false; emitted = 0;
          if(TAIL) { tail = 0; tail_n = 0; tail_end = pos; skip = 0; next_pack = pos; }
          if(WRITE) { out_at = out_offsets[q]; }
          if(MODE == 2) { out_at = mem_slot_start(begin, q, stride, shift); }
        }
      }
      next += __popc(dead);
      if(next > slice_end) { next = slice_end; }
    }
    if(__ballot_sync(0xFFFFFFFFu, live) == 0) { break; }

    // Two phases, chosen per warp: backward steps for the lanes that can take one, or parent() for the lanes
    // whose step failed.  parent() is several times longer than a step, so lanes waiting for it are held back
    // until `parent_batch` of them wait (or nobody can step): the long path then runs with many lanes active
    // instead of one or two.
    u32 waiting = __ballot_sync(0xFFFFFFFFu, live && need_parent);
    u32 stepping = __ballot_sync(0xFFFFFFFFu, live && !need_parent);
    if(waiting != 0 && ((u32)__popc(waiting) >= parent_batch || stepping == 0))
    {
      if(live && need_parent)
      {
        gcsa_b200_stnode node = lcp_parent(l, sp, ep);
        sp = node.sp; ep = node.ep; depth = node.node_lcp;
        need_parent = false;
      }
      continue;
    }
    if(!live || need_parent) { continue; }

    if(pos == begin)
    {
      if(depth > 0 && extended)
      {
        if(WRITE || (MODE == 2 && emitted < mem_slot_size(offsets[q + 1] - offsets[q], stride, shift))) { store_match(matches + 4 * (out_at + emitted), 0, depth, sp, ep); }
        emitted++;
      }
      if(!WRITE) { counts[q] = emitted; }
      live = false;
      continue;
    }
    if(JUMP)
    {
      u64 left = pos - begin;
      if(sp == ep && skip == 0 && left >= 4)
      {
        const u64* from = (left >= (u64)v.jump_k ? v.jump : v.jump_short);
        u64 e = (from != nullptr ? __ldg(from + sp) : 0);
        u32 len = (u32)(e >> 59);
        if(len >= 2)
        {
          u64 off = tail_end - pos;
          if(off + len > (u64)tail_n) { pack_tail(pos); off = 0; }
          if(len <= tail_n)
          {
            u64 stored = ((e << 5) >> 5) >> v.jump_tbits;
            u64 diff = ((tail >> (2 * off)) ^ stored) & ((1ull << (2 * len)) - 1);
            if(diff == 0)
            {
              sp = ep = (e & ((1ull << v.jump_tbits) - 1));
              depth += len; pos -= len; extended = true;
              continue;
            }
            skip = ((u32)(__ffsll((long long)diff) - 1) >> 1) + 2;       // single steps up to and including the one that fails
          }
          else { skip = 9; }                                             // a non-base or the start of the pattern is near: eight single steps
        }
      }
      if(skip > 0) { skip--; }
    }
    if(PACK && tail_end - pos >= (u64)tail_n && pos <= next_pack)
    {
      // the packed window is used up: the next 32 characters (fewer in front of a non-base or of the start of the
      // pattern: those go through the byte path, and packing is tried again eight characters further on)
      pack_tail(pos);
      next_pack = pos - (tail_n > 0 ? tail_n : (pos - begin < 8 ? pos - begin : 8));
    }
    u64 nsp, nep;
    lf_range(v, sp, ep, comp_at(pos - 1), nsp, nep);
    if(!range_empty(nsp, nep)) { sp = nsp; ep = nep; depth++; pos--; extended = true; continue; }
    if(depth == 0) { pos--; continue; }
    if(extended)
    {
      if(WRITE || (MODE == 2 && emitted < mem_slot_size(offsets[q + 1] - offsets[q], stride, shift))) { store_match(matches + 4 * (out_at + emitted), pos - begin, depth, sp, ep); }
      emitted++; extended = false;
    }
    need_parent = true;
  }
}

// scratch slots -> CSR, eight lanes per pattern (a match record is 32 bytes: the eight copy 256 contiguous bytes at a
// time); patterns with more matches than their slot holds are listed in `overflow`
__global__ void __launch_bounds__(256)
mem_gather_kernel(const ulonglong4* __restrict__ scratch, const u64* __restrict__ counts, const u64* __restrict__ out_offsets,
                  const u64* __restrict__ offsets, u64 char_base, u64 n, u64 stride, u32 shift,
                  ulonglong4* __restrict__ matches, u64* __restrict__ overflow, ull* __restrict__ n_overflow)
{
  const u32 sub = threadIdx.x & 7;
  const u64 groups = ((u64)gridDim.x * blockDim.x) >> 3;
  for(u64 q = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 3; q < n; q += groups)
  {
    u64 c = counts[q], begin = offsets[q] - char_base, len = offsets[q + 1] - offsets[q];
    if(c > mem_slot_size(len, stride, shift)) { if(sub == 0) { overflow[atomicAdd(n_overflow, 1ull)] = q; } continue; }
    const ulonglong4* src = scratch + mem_slot_start(begin, q, stride, shift);
    ulonglong4* dst = matches + out_offsets[q];
    for(u64 e = sub; e < c; e += 8) { dst[e] = src[e]; }
  }
}

__global__ void __launch_bounds__(256)
mem_count_overflow_kernel(const u64* __restrict__ counts, const u64* __restrict__ offsets, u64 n, u64 stride, u32 shift, ull* __restrict__ n_overflow)
{
  ull mine = 0;
  for(u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x)
  {
    mine += (counts[q] > mem_slot_size(offsets[q + 1] - offsets[q], stride, shift) ? 1 : 0);
  }
  for(int d = 16; d > 0; d >>= 1) { mine += __shfl_down_sync(0xFFFFFFFFu, mine, d); }
  if((threadIdx.x & 31) == 0 && mine > 0) { atomicAdd(n_overflow, mine); }
}

#endif
