/* device/layout.cuh -- device views of the index (DevView, LcpView) and the rank / select / LF primitives over the fused sectors.
   Included by every CUDA translation unit of the engine (device functions only, no kernels); sm_100a only. */
#ifndef GCSA2_B200_DEVICE_LAYOUT_CUH
#define GCSA2_B200_DEVICE_LAYOUT_CUH

#include <cuda_runtime.h>
#include <cstdint>

#include "../../../include/gcsa2_b200.h"

typedef uint64_t u64;
typedef unsigned long long ull;
typedef unsigned int u32;
typedef unsigned char u8;

#define BWT_W 87u
#define RV_W 192u
#define SEL_HINT 512u
#define M40 ((1ull << 40) - 1)
#define TABLE_ESCAPE 0xFFFFFFull

//------------------------------------------------------------------------------
// Device views
//------------------------------------------------------------------------------

struct RankVecDev { const ulonglong4* sec; u64 n_bits; u64 n_sec; };
struct SelVecDev  { RankVecDev rv; const u32* hints; u64 ones; };

struct DevView
{
  u64 path_nodes, edge_count;
  u64 C[GCSA_B200_SIGMA + 1];
  u64 char_sp[GCSA_B200_SIGMA], char_ep[GCSA_B200_SIGMA];
  const ulonglong4* bwt;
  const ulonglong4* bwt2;              // two-step blocks (16 sectors per block), or nullptr
  RankVecDev edges, sampled, extra_filter;
  SelVecDev extra_values, redundant;
  const u64* sparse_pos[3]; u64 sparse_n[3];       // comps 0, 5, 6
  const u64* stored_samples; const u64* sample_start; u64 sample_count;
  const u64* table; int table_k;      // entry = sp | length << 40; length 0xFFFFFF = not tabulated
  const ulonglong2* table2;            // fused form (replaces `table`): { that entry, the jump entry of sp if the range is a singleton, else 0 }
  const u32* walk32; const u64* walk64; // locate walk table: LF(i) << 1, or rank(sampled, i) << 1 | 1 for sampled nodes
  u32 default_alphabet;                // char2comp is exactly ACGT / acgt -> 1..4 for the bases (enables the SWAR pattern packing)
  const u64* jump; u32 jump_k, jump_tbits;   // jump table: len << 59 | 2-bit chars << jump_tbits | target (see jump_extend_kernel)
  const u64* jump_short;               // the same table cut at 4 steps: for the tail of a pattern that is shorter than the long path
  const ulonglong2* jump_wide;         // the long table with 16-byte entries { target | len << 40, characters } for indexes whose node numbers
                                       // leave fewer than 16 characters in an 8-byte entry (replaces `jump`; jump_k stays the longest path)
  const u64* loc64;                    // locate table: bit 63 | value for nodes with one start position, else rank of the sampled node << 24 | steps
  u8 char2comp[256];
};

struct LcpView
{
  u64 size, branching, levels, values;
  int shift;                           // log2(branching) if it is a power of two, else -1
  u64 offsets[16];
  const u8* data;
};

//------------------------------------------------------------------------------
// Device primitives
//------------------------------------------------------------------------------

__device__ __forceinline__ ulonglong4 ld256(const ulonglong4* p)
{
  ulonglong4 r;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ bool range_empty(u64 sp, u64 ep) { return (sp + 1 > ep + 1); }   // utils.h:93-101

// A jump-table entry, whichever table it came from: the unary backward path of a node.
struct JumpPath { u32 len; u64 chars, target; };      // chars: comp - 1, 2 bits each, first step lowest
__device__ __forceinline__ JumpPath jump_decode(u64 e, u32 tbits)
{
  JumpPath p; p.len = (u32)(e >> 59); p.chars = ((e << 5) >> 5) >> tbits; p.target = e & ((1ull << tbits) - 1);
  return p;
}
__device__ __forceinline__ JumpPath jump_decode_wide(ulonglong2 e)
{
  JumpPath p; p.len = (u32)(e.x >> 40) & 63u; p.chars = e.y; p.target = e.x & M40;
  return p;
}

// ones among the low k bits of w, 0 <= k <= 64
__device__ __forceinline__ u32 popc_low(u64 w, u32 k)
{
  u64 m = (k >= 64 ? ~0ull : ((1ull << k) - 1));
  return (u32)__popcll(w & m);
}

// ones among the low k bits of the (up to) 88-bit string hi:lo, 0 <= k <= 88
__device__ __forceinline__ u32 popc_low88(u64 lo, u32 hi, u32 k)
{
  u32 klo = (k < 64 ? k : 64), khi = k - klo;
  return popc_low(lo, klo) + (u32)__popc(hi & ((1u << khi) - 1));
}

__device__ __forceinline__ u64 rv_rank(const RankVecDev& v, u64 i)
{
  u64 s = i / RV_W; u32 off = (u32)(i - s * RV_W);
  ulonglong4 q = ld256(v.sec + s);
  u32 w = off >> 6, r = off & 63;
  u64 res = q.x;
  if(w > 0) { res += __popcll(q.y); }
  if(w > 1) { res += __popcll(q.z); }
  u64 word = (w == 0 ? q.y : (w == 1 ? q.z : q.w));
  return res + popc_low(word, r);
}

// bit i and rank(i) from one sector
__device__ __forceinline__ bool rv_get_rank(const RankVecDev& v, u64 i, u64& rank)
{
  u64 s = i / RV_W; u32 off = (u32)(i - s * RV_W);
  ulonglong4 q = ld256(v.sec + s);
  u32 w = off >> 6, r = off & 63;
  u64 res = q.x;
  if(w > 0) { res += __popcll(q.y); }
  if(w > 1) { res += __popcll(q.z); }
  u64 word = (w == 0 ? q.y : (w == 1 ? q.z : q.w));
  rank = res + popc_low(word, r);
  return (word >> r) & 1;
}

// position of the j-th (1-based) set bit of w; w has at least j set bits
__device__ __forceinline__ u32 select_in_word(u64 w, u32 j)
{
  u32 lo = (u32)w, c = __popc(lo);
  if(j <= c) { return __fns(lo, 0, j); }
  return 32 + __fns((u32)(w >> 32), 0, j - c);
}

// select1(k), k >= 1 (SadaCount / SadaSparse selects, support.h:253, 324)
__device__ __forceinline__ u64 sv_select(const SelVecDev& v, u64 k)
{
  u64 h = (k - 1) / SEL_HINT;
  u64 lo = v.hints[h], hi = v.hints[h + 1];
  // last sector in [lo, hi] whose cumulative count is < k
  while(lo < hi)
  {
    u64 mid = lo + (hi - lo + 1) / 2;
    u64 cum = __ldg(&(v.rv.sec[mid].x));
    if(cum < k) { lo = mid; } else { hi = mid - 1; }
  }
  ulonglong4 q = ld256(v.rv.sec + lo);
  u32 need = (u32)(k - q.x);
  u32 c0 = __popcll(q.y), c1 = __popcll(q.z);
  u64 base = lo * RV_W;
  if(need <= c0) { return base + select_in_word(q.y, need); }
  need -= c0;
  if(need <= c1) { return base + 64 + select_in_word(q.z, need); }
  need -= c1;
  return base + 128 + select_in_word(q.w, need);
}

// number of list entries < i
__device__ __forceinline__ u64 sparse_rank(const u64* pos, u64 n, u64 i)
{
  u64 lo = 0, hi = n;
  while(lo < hi)
  {
    u64 mid = (lo + hi) >> 1;
    if(__ldg(pos + mid) < i) { lo = mid + 1; } else { hi = mid; }
  }
  return lo;
}

__device__ __forceinline__ int sparse_slot(u32 c) { return (c == 0 ? 0 : (int)c - 4); }   // 0,5,6 -> 0,1,2

/*
  GCSA::LF(range, comp), include/gcsa/gcsa.h:155-162 with 262-274 and pathNodeRange 253-258.
  Fast characters: one fused sector per endpoint.  Sparse characters: list rank + edges rank.
*/
__device__ __forceinline__ void lf_range(const DevView& v, u64 sp, u64 ep, u32 c, u64& osp, u64& oep, u32* sectors = nullptr)
{
  if(c >= 1 && c <= GCSA_B200_FAST_CHARS)
  {
    u64 e1 = ep + 1;
    u64 bs = sp / BWT_W, be = e1 / BWT_W;
    u32 os = (u32)(sp - bs * BWT_W), oe = (u32)(e1 - be * BWT_W);
    ulonglong4 a = ld256(v.bwt + bs * 4 + (c - 1));
    ulonglong4 b = a;
    if(be != bs) { b = ld256(v.bwt + be * 4 + (c - 1)); }
    if(sectors) { *sectors += (be != bs ? 2 : 1); }
    u32 js = popc_low88(a.y, (u32)(a.x >> 40), os);
    u32 je = popc_low88(b.y, (u32)(b.x >> 40), oe);
    u64 f = (a.x & M40) + js;
    u64 s = (b.x & M40) + je - 1;
    if(range_empty(f, s)) { osp = f; oep = s; return; }
    osp = (a.z & M40) + popc_low88(a.w, (u32)(a.z >> 40), js + 1);
    oep = (b.z & M40) + popc_low88(b.w, (u32)(b.z >> 40), je);
  }
  else if(c < GCSA_B200_SIGMA)
  {
    int slot = sparse_slot(c);
    u64 f = v.C[c] + sparse_rank(v.sparse_pos[slot], v.sparse_n[slot], sp);
    u64 s = v.C[c] + sparse_rank(v.sparse_pos[slot], v.sparse_n[slot], ep + 1) - 1;
    if(range_empty(f, s)) { osp = f; oep = s; return; }
    osp = rv_rank(v.edges, f);
    oep = rv_rank(v.edges, s);
    if(sectors) { *sectors += 2; }
  }
  else { osp = 1; oep = 0; }    // not a comp value: Range::empty_range()
}

/*
  Two backward steps in one probe.  For a pair of fast characters (c1, c2) the block holds the
  same sector format over the "squared" graph: B2[i] = 1 iff node i has a predecessor j by c2 that
  itself has a predecessor h by c1; the 2-paths of one label, ordered by target, are ordered by
  source as well and consecutive sources differ by at most one node, so the source of the x-th
  2-path is H0 + popcount(boundary bits), exactly like rank(edges, .) in the one-step sector.
  Equivalent to LF(LF(range, c2), c1) whenever that is non-empty; returns false otherwise (the
  caller then takes the two single steps, which produce the reference's uncanonicalised pair).
*/
__device__ __forceinline__ bool lf2_range(const DevView& v, u64 sp, u64 ep, u32 c1, u32 c2, u64& osp, u64& oep, u32* sectors = nullptr)
{
  u64 e1 = ep + 1;
  u64 bs = sp / BWT_W, be = e1 / BWT_W;
  u32 os = (u32)(sp - bs * BWT_W), oe = (u32)(e1 - be * BWT_W);
  u32 label = (c1 - 1) * 4 + (c2 - 1);
  ulonglong4 a = ld256(v.bwt2 + bs * 16 + label);
  ulonglong4 b = a;
  if(be != bs) { b = ld256(v.bwt2 + be * 16 + label); }
  if(sectors) { *sectors += (be != bs ? 2 : 1); }
  u32 js = popc_low88(a.y, (u32)(a.x >> 40), os);
  u32 je = popc_low88(b.y, (u32)(b.x >> 40), oe);
  u64 f = (a.x & M40) + js;
  u64 s = (b.x & M40) + je - 1;
  if(range_empty(f, s)) { return false; }
  osp = (a.z & M40) + popc_low88(a.w, (u32)(a.z >> 40), js + 1);
  oep = (b.z & M40) + popc_low88(b.w, (u32)(b.z >> 40), je);
  return true;
}

/*
  GCSA::LF(path_node), include/gcsa/gcsa.h:165-183: first predecessor, fast characters first.
  One 128-byte line holds the four fast sectors of the node's block.
*/
__device__ __forceinline__ u64 lf_node(const DevView& v, u64 i)
{
  u64 b = i / BWT_W; u32 off = (u32)(i - b * BWT_W);
  const ulonglong4* line = v.bwt + b * 4;
  // the four sectors of the line are requested together (named registers: an array of them ends up in local memory)
  const ulonglong4 q0 = ld256(line), q1 = ld256(line + 1), q2 = ld256(line + 2), q3 = ld256(line + 3);
  auto has = [off](const ulonglong4& q) -> bool { return (off < 64 ? (q.y >> off) & 1 : ((q.x >> 40) >> (off - 64)) & 1); };
  auto pred = [off](const ulonglong4& q) -> u64
  {
    u32 j = popc_low88(q.y, (u32)(q.x >> 40), off);
    return (q.z & M40) + popc_low88(q.w, (u32)(q.z >> 40), j + 1);
  };
  if(has(q0)) { return pred(q0); }
  if(has(q1)) { return pred(q1); }
  if(has(q2)) { return pred(q2); }
  if(has(q3)) { return pred(q3); }
  for(u32 c = GCSA_B200_FAST_CHARS + 1; c < GCSA_B200_SIGMA; c++)
  {
    int slot = sparse_slot(c);
    u64 r = sparse_rank(v.sparse_pos[slot], v.sparse_n[slot], i);
    if(r < v.sparse_n[slot] && v.sparse_pos[slot][r] == i) { return rv_rank(v.edges, v.C[c] + r); }
  }
  return rv_rank(v.edges, v.C[0] + sparse_rank(v.sparse_pos[0], v.sparse_n[0], i));
}

// bit B_c[i] for any comp (used by LF_fast / LF_all single-node shortcut, src/gcsa.cpp:748-756)
__device__ __forceinline__ bool bwt_bit(const DevView& v, u64 i, u32 c)
{
  if(c >= 1 && c <= GCSA_B200_FAST_CHARS)
  {
    u64 b = i / BWT_W; u32 off = (u32)(i - b * BWT_W);
    ulonglong4 q = ld256(v.bwt + b * 4 + (c - 1));
    return (off < 64 ? (q.y >> off) & 1 : ((q.x >> 40) >> (off - 64)) & 1);
  }
  int slot = sparse_slot(c);
  u64 r = sparse_rank(v.sparse_pos[slot], v.sparse_n[slot], i);
  return (r < v.sparse_n[slot] && v.sparse_pos[slot][r] == i);
}

//------------------------------------------------------------------------------
// Shared by several translation units: the single-node step and the packing of pattern characters
//------------------------------------------------------------------------------

// predecessor of node i by fast character c (0-based) from its fused sector, or false
__device__ __forceinline__ bool pred_fast(const DevView& v, u64 i, u32 c, u64& pred)
{
  u64 b = i / BWT_W; u32 off = (u32)(i - b * BWT_W);
  ulonglong4 q = ld256(v.bwt + b * 4 + c);
  bool bit = (off < 64 ? (q.y >> off) & 1 : ((q.x >> 40) >> (off - 64)) & 1);
  if(!bit) { return false; }
  u32 j = popc_low88(q.y, (u32)(q.x >> 40), off);
  pred = (q.z & M40) + popc_low88(q.w, (u32)(q.z >> 40), j + 1);
  return true;
}

// Pattern bytes are read through an 8-byte window (one aligned streaming load per 8 characters,
// evict-first: the pattern stream must not push index lines out of the L2).
struct CharWindow
{
  u64 word; u64 index;
  __device__ __forceinline__ CharWindow() : word(0), index(~0ull) {}
  __device__ __forceinline__ u32 get(const u8* chars, u64 pos)
  {
    u64 addr = (u64)(chars + pos);
    u64 wi = addr >> 3;
    if(wi != index) { word = __ldcs((const unsigned long long*)(wi << 3)); index = wi; }
    return (u32)((word >> ((addr & 7) * 8)) & 0xFF);
  }
};




// Eight pattern bytes of the default alphabet (w: lowest address in the low byte) -> their comp - 1 codes,
// 2 bits each, the LAST byte in the lowest bits.  *good = how many bytes, counted from the last one, are bases
// in either case (8 if all); the codes of the others are garbage.
__device__ __forceinline__ u32 pack8_reversed(u64 w, u32* good)
{
  const u64 L7 = 0x7F7F7F7F7F7F7F7Full, H8 = 0x8080808080808080ull;
  u64 x = w & 0xDFDFDFDFDFDFDFDFull;
  u64 zA = x ^ 0x4141414141414141ull, zC = x ^ 0x4343434343434343ull, zG = x ^ 0x4747474747474747ull, zT = x ^ 0x5454545454545454ull;
  // 0x80 in every byte that equals one of the four letters (exact zero-byte test, no carries between bytes)
  u64 valid = ~(((zA & L7) + L7) | zA | L7) | ~(((zC & L7) + L7) | zC | L7) | ~(((zG & L7) + L7) | zG | L7) | ~(((zT & L7) + L7) | zT | L7);
  u64 inv = ~valid & H8;
  *good = (inv == 0 ? 8u : 7u - (u32)((63 - __clzll((long long)inv)) >> 3));
  u64 t = (w >> 1) & 0x0303030303030303ull;                      // A 0, C 1, T 2, G 3
  u64 code = t ^ ((t >> 1) & 0x0101010101010101ull);               // A 0, C 1, G 2, T 3
  u64 y = (code | (code >> 6)) & 0x000F000F000F000Full;
  y = (y | (y >> 12)) & 0x000000FF000000FFull;
  y = (y | (y >> 24)) & 0xFFFFull;
  u32 r = __brev((u32)y) >> 16;                                    // reverse the order of the characters ...
  return ((r >> 1) & 0x5555u) | ((r & 0x5555u) << 1);              // ... not of the two bits of each
}

#endif
