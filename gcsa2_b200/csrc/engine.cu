/*
  engine.cu -- B200 (sm_100a) batched backward-search engine behind the C ABI of
  include/gcsa2_b200.h.  Hand-written CUDA; no tensor cores (integer rank/select work).

  Device layout (DESIGN.md "Data layout in HBM"):

  * FUSED BWT BLOCKS.  The four fast characters (A,C,G,T; fast_bwt[1..4] of
    include/gcsa/gcsa.h:217-219) are interleaved: block b covers path nodes [87b, 87b+87) and is
    one 128-byte line of four 32-byte sectors, one per character c:
        w0 = (C[c] + rank(B_c, 87b))            [40 bits] | B_c bits 64..86   [23 bits] << 40
        w1 = B_c bits 0..63 of the block
        w2 = rank(edges, P - 1)                 [40 bits] | window bits 64..87 [24 bits] << 40
        w3 = window bits 0..63,   window bit t = edges[P - 1 + t],  P = C[c] + rank(B_c, 87b)
    so that one endpoint of GCSA::LF(range, c) (include/gcsa/gcsa.h:155-162, 253-274) -- the B_c
    rank AND the dependent rank on `edges` -- is ONE 32-byte sector read (LDG.E.256) instead of
    two cache-line probes in two vectors, and GCSA::LF(node) (gcsa.h:165-183) is one 128-byte line.
  * RANK VECTORS (edges, sampled_paths, SadaSparse::filter): 32-byte sectors {cumulative count,
    192 data bits}; a rank probe or bit access is one sector.
  * SELECT VECTORS (SadaSparse::values, SadaCount::data): rank vector + one hint per 512 ones.
  * samples/select (gcsa.h:235-236) is replaced by an explicit start-offset array per sampled node.
  * sparse characters ($, N, #; sparse_bwt of gcsa.h:221-223) are sorted position lists.
  * optional k-mer table: find() results of all 4^k ACGT strings of length k, 8 bytes each
    (sp in 40 bits, range length in 24 bits; an empty result always has ep = sp - 1).
*/
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <omp.h>
#include <random>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/gcsa2_b200.h"
#include "internal.h"

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
// Errors
//------------------------------------------------------------------------------

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
  return fail(GCSA_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } } while(0)

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
  const u64* jump; u32 jump_k, jump_tbits;
  const u64* jump_short;               // the same table cut at 4 steps: for the tail of a pattern that is shorter than the long path   // jump table: len << 59 | 2-bit chars << jump_tbits | target (see jump_extend_kernel)
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
  ulonglong4 q[4];
  #pragma unroll
  for(int c = 0; c < 4; c++) { q[c] = ld256(line + c); }
  #pragma unroll
  for(int c = 0; c < 4; c++)
  {
    bool bit = (off < 64 ? (q[c].y >> off) & 1 : ((q[c].x >> 40) >> (off - 64)) & 1);
    if(bit)
    {
      u32 j = popc_low88(q[c].y, (u32)(q[c].x >> 40), off);
      return (q[c].z & M40) + popc_low88(q[c].w, (u32)(q[c].z >> 40), j + 1);
    }
  }
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
// Kernels: find
//------------------------------------------------------------------------------

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



struct FindStatsDev { u64 found, total_length, lf_steps, sector_probes, table_hits; };

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

/*
  GCSA::find(begin, end), include/gcsa/gcsa.h:96-110.  One query per lane.  Queries are pulled from a
  contiguous per-warp slice; lanes whose search ended are refilled together once half the warp is idle
  (one ballot + popc, no atomics).  A refilled lane packs the last 32 characters of its pattern into one
  register (2 bits each, the last character lowest): the k-mer table index is a bit field of it and a jump
  along a unary path is one XOR against the table entry.  Anything that does not fit the fast forms (other
  characters, another alphabet, short remainders) goes through the per-character path, which is the
  reference's loop verbatim.
*/
template<bool STATS, int MIN_BLOCKS, bool PACKED = false>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
find_kernel(const DevView v, const u8* __restrict__ chars, const u64* __restrict__ offsets, u64 char_base,
            u64 fixed_length, u64 n, u64* __restrict__ sp_out, u64* __restrict__ ep_out, FindStatsDev* stats, int refill_at)
{
  // PACKED: `chars` holds ceil(fixed_length / 32) 64-bit words per pattern, character p of a pattern at bits
  // [2 (p % 32), 2 (p % 32) + 2) of word p / 32, value comp - 1 (ACGT only; packed by the host entry point).
  __shared__ u8 c2c[256];
  if(!PACKED)
  {
    for(int i = threadIdx.x; i < 256; i += blockDim.x) { c2c[i] = v.char2comp[i]; }
    __syncthreads();
  }
  const u64 words_per_pattern = (fixed_length + 31) >> 5;
  const bool fast_pack = (PACKED || v.default_alphabet != 0);

  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  // contiguous slice of queries for this warp
  const u64 per = (n + n_warps - 1) / n_warps;
  u64 next = warp * per;
  const u64 slice_end = (next + per < n ? next + per : n);
  if(next >= n) { return; }

  u64 q = ~0ull, sp = 0, ep = 0, pos = 0, begin = 0;
  u64 tail = 0, tail_end = 0; u32 tail_n = 0;       // characters [tail_end - tail_n, tail_end), the one at tail_end - 1 - t in bits [2t, 2t + 2)
  bool live = false;
  u32 jump_mode = 1;                   // 0 once a jump failed on a character: this query dies within a few single steps
  CharWindow win;
  u64 st_found = 0, st_len = 0, st_steps = 0, st_sectors = 0, st_hits = 0;

  // comp value of the character at (batch-wide) position p of the current query: the general path
  auto comp_slow = [&](u64 p) -> u32
  {
    if(PACKED)
    {
      u64 rel = p - begin, wi = q * words_per_pattern + (rel >> 5);
      if(wi != win.index) { win.word = __ldcs((const unsigned long long*)chars + wi); win.index = wi; }
      return (u32)((win.word >> ((rel & 31) * 2)) & 3) + 1;
    }
    return c2c[win.get(chars, p)];
  };
  auto comp_at = [&](u64 p) -> u32
  {
    u64 off = tail_end - 1 - p;
    if(off < (u64)tail_n) { return (u32)((tail >> (2 * off)) & 3) + 1; }
    return comp_slow(p);
  };
  // pack the (up to) 32 characters that end at position `end_pos` (exclusive)
  auto pack_tail = [&](u64 end_pos)
  {
    tail = 0; tail_n = 0; tail_end = end_pos;
    if(!fast_pack) { return; }
    if constexpr(PACKED)
    {
      u64 have = end_pos - begin, m = (have < 32 ? have : 32), r0 = have - m;             // pattern-relative [r0, r0 + m)
      const unsigned long long* words = (const unsigned long long*)chars + q * words_per_pattern;
      u32 sh = (u32)(r0 & 31) * 2;
      u64 x = __ldcs(words + (r0 >> 5)) >> sh;
      if(sh != 0 && (r0 & 31) + m > 32) { x |= __ldcs(words + (r0 >> 5) + 1) << (64 - sh); }
      u64 r = __brevll(x);
      r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
      tail = (m < 32 ? r >> (2 * (32 - m)) : r);
      tail_n = (u32)m;
    }
    else
    {
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
    }
  };

  while(true)
  {
    // refill: all idle lanes at once, as soon as half the warp is idle (or nobody is working)
    u32 dead = __ballot_sync(0xFFFFFFFFu, !live);
    if(__popc(dead) >= refill_at)
    {
      u32 my = __popc(dead & ((1u << lane) - 1));
      if(!live)
      {
        u64 cand = next + my;
        if(cand < slice_end)
        {
          q = cand;
          u64 b, e;
          if(offsets != nullptr) { b = offsets[q] - char_base; e = offsets[q + 1] - char_base; }
          else { b = q * fixed_length; e = b + fixed_length; }
          begin = b; live = true; jump_mode = 1;
          tail = 0; tail_n = 0; tail_end = e;
          if(e == b || v.path_nodes == 0) { sp = 0; ep = v.path_nodes - 1; pos = b; }
          else
          {
            pack_tail(e);
            pos = e - 1;
            bool used_table = false;
            if(v.table_k > 0 && e - b >= (u64)v.table_k)
            {
              u64 idx = 0; bool ok = true;
              if(tail_n >= (u32)v.table_k) { idx = tail & ((1ull << (2 * v.table_k)) - 1); }
              else
              {
                for(int t = 0; t < v.table_k; t++)
                {
                  u32 c = comp_at(e - 1 - t);
                  ok = ok && (c >= 1 && c <= 4);
                  idx |= (u64)((c - 1) & 3) << (2 * t);
                }
              }
              if(ok)
              {
                u64 r, je = 0;
                if(v.table2 != nullptr) { ulonglong2 both = __ldg(v.table2 + idx); r = both.x; je = both.y; }
                else { r = __ldg(v.table + idx); }
                u64 len = r >> 40;
                if(len != TABLE_ESCAPE)
                {
                  sp = r & M40; ep = sp + len - 1; pos = e - v.table_k; used_table = true;
                  if(STATS) { st_hits++; }
                  // Fused table: the jump entry of a singleton result came with the same 16-byte load, so the first
                  // jump costs no probe.  Taken only when the whole path lies inside the packed tail and inside
                  // the pattern; everything else is left to the main loop.
                  u32 jl = (u32)(je >> 59);
                  if(jl >= 2 && (u64)jl <= pos - b && (u32)v.table_k + jl <= tail_n)
                  {
                    u64 stored = ((je << 5) >> 5) >> v.jump_tbits;
                    if((((tail >> (2 * v.table_k)) ^ stored) & ((1ull << (2 * jl)) - 1)) == 0)
                    {
                      sp = ep = (je & ((1ull << v.jump_tbits) - 1));
                      pos -= jl;
                      if(STATS) { st_steps += jl; }
                    }
                    else { jump_mode = 0; }                          // leaves the unary path: it dies within these steps
                  }
                }
              }
            }
            if(!used_table)
            {
              u32 c = comp_at(pos);
              sp = v.char_sp[c]; ep = v.char_ep[c];
            }
          }
        }
      }
      next += __popc(dead);
      if(next > slice_end) { next = slice_end; }
    }
    if(__ballot_sync(0xFFFFFFFFu, live) == 0)
    {
      if(next >= slice_end) { break; }
      continue;
    }

    if(live)
    {
      if(!(range_empty(sp, ep) || pos == begin))
      {
        u32 sectors = 0;
        bool done = false;
        // Singleton range: try the jump table (one load for up to jump_k backward steps along a unary path).
        // The table is chosen by what is left of the pattern, so that a path never overshoots its end: the long
        // table (paths of up to jump_k steps) while at least jump_k characters remain, the short one (4) below that.
        const u64* jump_from = nullptr;
        if(v.jump != nullptr && jump_mode != 0 && sp == ep)
        {
          u64 left = pos - begin;
          jump_from = (left >= (u64)v.jump_k ? v.jump : (left >= 4 ? v.jump_short : nullptr));
        }
        if(jump_from != nullptr)
        {
          u64 e = __ldg(jump_from + sp);
          u32 len = (u32)(e >> 59);
          if(STATS) { sectors++; }
          if(len >= 2)
          {
            u64 stored = ((e << 5) >> 5) >> v.jump_tbits;
            u64 off = tail_end - pos;
            if(off + len > (u64)tail_n && fast_pack && tail_n == 32) { pack_tail(pos); off = 0; }
            bool same = true;
            if(off + len <= (u64)tail_n) { same = ((((tail >> (2 * off)) ^ stored) & ((1ull << (2 * len)) - 1)) == 0); }
            else
            {
              for(u32 t = 0; t < len; t++)
              {
                u32 pc = comp_at(pos - 1 - t);
                same = same && (pc == ((u32)(stored >> (2 * t)) & 3) + 1);
              }
            }
            if(same)
            {
              sp = ep = (e & ((1ull << v.jump_tbits) - 1));
              pos -= len; done = true;
              if(STATS) { st_steps += len; }
            }
            else { jump_mode = 0; }                                  // it dies within these steps: the exact pair comes from single steps
          }
        }
        if(!done && tail_end - pos >= (u64)tail_n && fast_pack && tail_n == 32) { pack_tail(pos); }   // next window of a long pattern
        u32 c = (done ? 0 : comp_at(pos - 1));
        if(!done && v.bwt2 != nullptr && pos - begin >= 2 && c >= 1 && c <= 4)
        {
          u32 c1 = comp_at(pos - 2);
          if(c1 >= 1 && c1 <= 4 && lf2_range(v, sp, ep, c1, c, sp, ep, STATS ? &sectors : nullptr))
          {
            pos -= 2; done = true;
            if(STATS) { st_steps += 2; }
          }
        }
        if(!done)
        {
          pos--;
          lf_range(v, sp, ep, c, sp, ep, STATS ? &sectors : nullptr);
          if(STATS) { st_steps++; }
        }
        if(STATS) { st_sectors += sectors; }
      }
      if(range_empty(sp, ep) || pos == begin)
      {
        __stcs((unsigned long long*)sp_out + q, (unsigned long long)sp); __stcs((unsigned long long*)ep_out + q, (unsigned long long)ep);
        if(STATS && !range_empty(sp, ep)) { st_found++; st_len += ep + 1 - sp; }
        live = false;
      }
    }
  }

  if(STATS)
  {
    atomicAdd((ull*)&stats->found, (ull)st_found); atomicAdd((ull*)&stats->total_length, (ull)st_len);
    atomicAdd((ull*)&stats->lf_steps, (ull)st_steps); atomicAdd((ull*)&stats->sector_probes, (ull)st_sectors);
    atomicAdd((ull*)&stats->table_hits, (ull)st_hits);
  }
}

/*
  k-mer table.  Entry idx describes the string whose t-th character from the END is comp
  ((idx >> 2t) & 3) + 1 and holds exactly what find() returns for it, early exit included: an
  empty result keeps the uncanonicalised pair of the step where the search died, and such a pair
  always has ep = sp - 1 (rank is monotone), so (sp, length) loses nothing.
  The table is grown one character at a time: level j+1 is one LF step away from level j.
*/
__global__ void __launch_bounds__(256)
table_init_kernel(const DevView v, ulonglong2* tmp)
{
  u32 idx = threadIdx.x;
  if(idx < 4) { tmp[idx] = make_ulonglong2(v.char_sp[idx + 1], v.char_ep[idx + 1]); }
}

// level j (4^j entries in tmp[0, 4^j)) -> level j + 1 in place: slot idx | c << 2j
__global__ void __launch_bounds__(256)
table_extend_kernel(const DevView v, int j, ulonglong2* tmp)
{
  u64 total = 1ull << (2 * j);
  for(u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x)
  {
    ulonglong2 r = tmp[idx];
    #pragma unroll
    for(u32 c = 4; c-- > 0; )
    {
      u64 sp = r.x, ep = r.y;
      if(!range_empty(sp, ep)) { lf_range(v, sp, ep, c + 1, sp, ep); }
      tmp[idx | ((u64)c << (2 * j))] = make_ulonglong2(sp, ep);
    }
  }
}

// last level: level k - 1 in tmp -> packed level k in table (k >= 2); for k == 1 pack tmp itself
// With table2 != nullptr the fused form is written instead: next to each entry the jump-table entry of its sp when the
// result is a single path node (find_kernel then takes the first jump without another probe).
__global__ void __launch_bounds__(256)
table_final_kernel(const DevView v, int k, const ulonglong2* tmp, u64* table, ulonglong2* table2)
{
  u64 total = (k == 1 ? 4 : 1ull << (2 * (k - 1)));
  for(u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x)
  {
    ulonglong2 r = tmp[idx];
    for(u32 c = 0; c < (k == 1 ? 1u : 4u); c++)
    {
      u64 sp = r.x, ep = r.y;
      if(k > 1 && !range_empty(sp, ep)) { lf_range(v, sp, ep, c + 1, sp, ep); }
      u64 len = ep + 1 - sp;
      u64 entry = (len >= TABLE_ESCAPE || sp > M40) ? (TABLE_ESCAPE << 40) : (sp | (len << 40));
      u64 slot = (k == 1 ? idx : (idx | ((u64)c << (2 * (k - 1)))));
      if(table2 == nullptr) { table[slot] = entry; }
      else { table2[slot] = make_ulonglong2(entry, (len == 1 && v.jump != nullptr) ? __ldg(v.jump + sp) : 0ull); }
    }
  }
}

//------------------------------------------------------------------------------
// Kernels: construction of the two-step blocks from the one-step blocks
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

// 16-bit mask per node: bit c1 * 4 + c2 set iff the 2-path (c1, c2) into the node exists;
// per block and label the number of set bits.
__global__ void __launch_bounds__(128)
two_step_mask_kernel(const DevView v, u64 n_blocks, unsigned short* m2, u32* blockpop)
{
  for(u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x)
  {
    u32 count[16];
    #pragma unroll
    for(int p = 0; p < 16; p++) { count[p] = 0; }
    for(u32 t = 0; t < BWT_W; t++)
    {
      u64 i = b * BWT_W + t;
      if(i >= v.path_nodes) { break; }
      u32 m = 0;
      for(u32 c2 = 0; c2 < 4; c2++)
      {
        u64 j;
        if(!pred_fast(v, i, c2, j)) { continue; }
        for(u32 c1 = 0; c1 < 4; c1++)
        {
          u64 h;
          if(pred_fast(v, j, c1, h)) { m |= 1u << (c1 * 4 + c2); }
        }
      }
      m2[i] = (unsigned short)m;
      #pragma unroll
      for(int p = 0; p < 16; p++) { count[p] += (m >> p) & 1; }
    }
    #pragma unroll
    for(int p = 0; p < 16; p++) { blockpop[(u64)p * n_blocks + b] = count[p]; }
  }
}

// source node of every 2-path, label by label, in target order
__global__ void __launch_bounds__(128)
two_step_source_kernel(const DevView v, u64 n_blocks, const unsigned short* m2, const u64* blockcnt,
                       const u64* label_base, u64* src)
{
  for(u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x)
  {
    u64 x[16];
    #pragma unroll
    for(int p = 0; p < 16; p++) { x[p] = label_base[p] + blockcnt[(u64)p * n_blocks + b]; }
    for(u32 t = 0; t < BWT_W; t++)
    {
      u64 i = b * BWT_W + t;
      if(i >= v.path_nodes) { break; }
      u32 m = m2[i];
      if(m == 0) { continue; }
      for(u32 c2 = 0; c2 < 4; c2++)
      {
        if(((m >> c2) & 0x1111u) == 0) { continue; }
        u64 j;
        if(!pred_fast(v, i, c2, j)) { continue; }
        for(u32 c1 = 0; c1 < 4; c1++)
        {
          u32 p = c1 * 4 + c2;
          u64 h;
          if(((m >> p) & 1) && pred_fast(v, j, c1, h))
          {
            #pragma unroll
            for(int q = 0; q < 16; q++) { if(q == (int)p) { src[x[q]] = h; x[q]++; } }
          }
        }
      }
    }
  }
}

// consecutive sources of one label must be equal or differ by one node
__global__ void __launch_bounds__(256)
two_step_validate_kernel(const u64* src, const u64* label_base, u32* violations)
{
  for(int p = 0; p < 16; p++)
  {
    u64 lo = label_base[p], hi = label_base[p + 1];
    for(u64 x = lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; x + 1 < hi; x += (u64)gridDim.x * blockDim.x)
    {
      u64 d = src[x + 1] - src[x];
      if(d > 1) { atomicAdd(violations, 1u); }
    }
  }
}

// one thread per (block, label): assemble the sector
__global__ void __launch_bounds__(256)
two_step_build_kernel(u64 path_nodes, u64 n_blocks, const unsigned short* m2, const u64* blockcnt,
                      const u64* label_base, const u64* src, ulonglong4* out)
{
  u64 total = n_blocks * 16;
  for(u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x)
  {
    u64 b = t >> 4; u32 p = (u32)(t & 15);
    u64 x0 = blockcnt[(u64)p * n_blocks + b];
    u64 blo = 0, bhi = 0;
    for(u32 k = 0; k < BWT_W; k++)
    {
      u64 i = b * BWT_W + k;
      if(i >= path_nodes) { break; }
      u64 bit = (m2[i] >> p) & 1;
      if(k < 64) { blo |= bit << k; } else { bhi |= bit << (k - 64); }
    }
    const u64* list = src + label_base[p];
    u64 len = label_base[p + 1] - label_base[p];
    u64 h0 = 0, wlo = 0, whi = 0;
    if(len > 0)
    {
      // window bit k describes 2-path x0 - 1 + k: 1 iff it is the last 2-path of its source
      h0 = (x0 == 0 ? list[0] : list[x0 - 1]);
      for(u32 k = (x0 == 0 ? 1 : 0); k < 88; k++)
      {
        u64 x = x0 - 1 + k;
        if(x + 1 >= len) { break; }
        u64 bit = (list[x] != list[x + 1]) ? 1 : 0;
        if(k < 64) { wlo |= bit << k; } else { whi |= bit << (k - 64); }
      }
    }
    ulonglong4 q;
    q.x = (x0 & M40) | (bhi << 40); q.y = blo;
    q.z = (h0 & M40) | (whi << 40); q.w = wlo;
    out[t] = q;
  }
}

//------------------------------------------------------------------------------
// Kernels: LF, count
//------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
lf_kernel(const DevView v, const u64* __restrict__ sp, const u64* __restrict__ ep, const u8* __restrict__ comp,
          u64 n, u64* __restrict__ osp, u64* __restrict__ oep)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 a, b;
    lf_range(v, sp[i], ep[i], comp[i], a, b);
    osp[i] = a; oep[i] = b;
  }
}

__global__ void __launch_bounds__(256)
lf_node_kernel(const DevView v, const u64* __restrict__ nodes, u64 n, u64* __restrict__ out)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    out[i] = lf_node(v, nodes[i]);
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
    if(c >= 1 && c <= last && !range_empty(s, e))
    {
      if(s == e)                                             // single path node: follow set bits only
      {
        if(bwt_bit(v, s, c)) { lf_range(v, s, e, c, a, b); }
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
    if(s == e) { if(bwt_bit(v, s, c)) { lf_range(v, s, e, c, a, b); } }      // gcsa.cpp:748-756, 774-789
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

//------------------------------------------------------------------------------
// Kernels: compareKMers (src/algorithms.cpp:505-616) -- the tries of two indexes in lockstep
//------------------------------------------------------------------------------

// One child range of LF_fast / LF_all (src/gcsa.cpp:742-798): empty input and a single path node
// without the predecessor give Range::empty_range(); the general case gives LF() uncanonicalised.
__device__ __forceinline__ void trie_child(const DevView& v, u64 s, u64 e, u32 c, u64& a, u64& b)
{
  a = 1; b = 0;
  if(range_empty(s, e)) { return; }
  if(s == e) { if(bwt_bit(v, s, c)) { lf_range(v, s, e, c, a, b); } }
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
  ull mine[3] = { 0, 0, 0 };
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 llen = st[n + i] + 1 - st[i], rlen = st[3 * n + i] + 1 - st[2 * n + i];
    u32 side = (llen > 0 && rlen > 0 ? 0 : (llen > 0 ? 1 : 2));
    mine[side]++;
    if(left_flag != nullptr) { left_flag[i] = (side == 1); right_flag[i] = (side == 2); }
  }
  for(int k = 0; k < 3; k++)
  {
    ull x = mine[k];
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

// Walk table for locate: one entry per path node, so that a step of locateInternal()
// (sampled(i) + LF(i), gcsa.cpp:882-887) is a single load.
template<class T>
__global__ void __launch_bounds__(256)
walk_table_kernel(const DevView v, T* table)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < v.path_nodes; i += (u64)gridDim.x * blockDim.x)
  {
    u64 r;
    if(rv_get_rank(v.sampled, i, r)) { table[i] = (T)((r << 1) | 1); }
    else { table[i] = (T)(lf_node(v, i) << 1); }
  }
}

/*
  Jump table for find(): for a path node i whose backward path is unary for len steps (every node on it has
  exactly one predecessor character, a base), the entry holds those len characters and the node reached:
  LF applied len times to the singleton range [i, i] gives exactly [target, target] when the pattern continues
  with these characters (each step maps a singleton to a singleton), so one load replaces len backward steps.
  Entry: len (5 bits) << 59 | characters (comp - 1, 2 bits each, first step lowest) << tbits | target (tbits).
  Level 1 is computed from the fused blocks and the sparse lists, longer paths by appending level-1 entries.
*/
__global__ void __launch_bounds__(256)
jump_init_kernel(const DevView v, u32 tbits, u64* __restrict__ one, u64* __restrict__ table)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < v.path_nodes; i += (u64)gridDim.x * blockDim.x)
  {
    u64 b = i / BWT_W; u32 off = (u32)(i - b * BWT_W);
    const ulonglong4* line = v.bwt + b * 4;
    u32 found = 0, which = 0; u64 target = 0;
    #pragma unroll
    for(int c = 0; c < 4; c++)
    {
      ulonglong4 q = ld256(line + c);
      bool bit = (off < 64 ? (q.y >> off) & 1 : ((q.x >> 40) >> (off - 64)) & 1);
      if(bit)
      {
        u32 j = popc_low88(q.y, (u32)(q.x >> 40), off);
        target = (q.z & M40) + popc_low88(q.w, (u32)(q.z >> 40), j + 1);
        which = (u32)c; found++;
      }
    }
    bool sparse = false;
    for(int slot = 0; slot < 3; slot++)
    {
      u64 r = sparse_rank(v.sparse_pos[slot], v.sparse_n[slot], i);
      if(r < v.sparse_n[slot] && v.sparse_pos[slot][r] == i) { sparse = true; }
    }
    u64 e = 0;
    if(found == 1 && !sparse) { e = (1ull << 59) | ((u64)which << tbits) | target; }
    one[i] = e; table[i] = e;
  }
}

// entries of length exactly j grow to j + 1 if the node they reach has a level-1 entry
__global__ void __launch_bounds__(256)
jump_extend_kernel(u64 n, u32 tbits, u32 j, const u64* __restrict__ one, u64* __restrict__ table)
{
  const u64 tmask = (1ull << tbits) - 1;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 e = table[i];
    if((e >> 59) != j) { continue; }
    u64 next = __ldg(one + (e & tmask));
    if((next >> 59) == 0) { continue; }
    u64 chars = ((e << 5) >> 5) >> tbits;
    chars |= ((next >> tbits) & 3) << (2 * j);
    table[i] = ((u64)(j + 1) << 59) | (chars << tbits) | (next & tmask);
  }
}

// Locate table: the whole of locateInternal() (gcsa.cpp:880-896) per path node, precomputed from the walk
// table.  A node whose sampled ancestor stores one start position holds that position + steps directly
// (bit 63 set); otherwise the rank of the sampled node and the number of steps.  *overflow is set if a
// field does not fit (the table is then dropped).
__global__ void __launch_bounds__(256)
locate_table_kernel(const DevView v, u64* table, int* overflow)
{
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < v.path_nodes; i += (u64)gridDim.x * blockDim.x)
  {
    u64 r, steps = 0;
    if(v.walk32 != nullptr)
    {
      u32 e = __ldg(v.walk32 + i);
      while(!(e & 1)) { e = __ldg(v.walk32 + (e >> 1)); steps++; }
      r = e >> 1;
    }
    else
    {
      u64 e = __ldg(v.walk64 + i);
      while(!(e & 1)) { e = __ldg(v.walk64 + (e >> 1)); steps++; }
      r = e >> 1;
    }
    u64 s0 = v.sample_start[r], s1 = v.sample_start[r + 1];
    u64 value = v.stored_samples[s0] + steps;
    if(steps >= (1ull << 24) || r >= (1ull << 39)) { *overflow = 1; table[i] = 0; }
    else if(s1 - s0 == 1 && value < (1ull << 63)) { table[i] = (1ull << 63) | value; }
    else { table[i] = (r << 24) | steps; }
  }
}

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
// MODE 2: count AND write the first `stride` matches of pattern q at scratch slot q * stride (one pass; the
// few patterns with more matches are redone in MODE 1 over the id list `ids`).
// JUMP: singleton ranges advance along the unary backward path of their node with one load (the jump tables of
// find_kernel): the pattern is kept 2-bit packed, 32 characters at a time, and a path of up to 16 steps is one XOR
// against it.  A path that the pattern leaves after t characters is followed by t + 1 single steps (the last of
// which fails, as it must), so matches, depths and ranges are those of the single-step loop.
template<int MODE, bool JUMP = false>
__global__ void __launch_bounds__(256)
mem_kernel(const DevView v, const LcpView l, const u8* __restrict__ chars, const u64* __restrict__ offsets, u64 char_base,
           u64 n, u64* __restrict__ counts, const u64* __restrict__ out_offsets, u64* __restrict__ matches,
           const u64* __restrict__ ids, u64 stride, u32 parent_batch)
{
  constexpr bool WRITE = (MODE == 1);
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
  u64 tail = 0, tail_end = 0; u32 tail_n = 0, skip = 0;     // JUMP: characters [tail_end - tail_n, tail_end) packed as in find_kernel

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
    if(JUMP)
    {
      u64 off = tail_end - 1 - p;
      if(off < (u64)tail_n) { return (u32)((tail >> (2 * off)) & 3) + 1; }
    }
    return c2c[chars[p]];
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
          if(JUMP) { tail = 0; tail_n = 0; tail_end = pos; skip = 0; }
          if(WRITE) { out_at = out_offsets[q]; }
          if(MODE == 2) { out_at = q * stride; }
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
        if(WRITE || (MODE == 2 && emitted < stride)) { u64* m = matches + 4 * (out_at + emitted); m[0] = 0; m[1] = depth; m[2] = sp; m[3] = ep; }
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
    u64 nsp, nep;
    lf_range(v, sp, ep, comp_at(pos - 1), nsp, nep);
    if(!range_empty(nsp, nep)) { sp = nsp; ep = nep; depth++; pos--; extended = true; continue; }
    if(depth == 0) { pos--; continue; }
    if(extended)
    {
      if(WRITE || (MODE == 2 && emitted < stride)) { u64* m = matches + 4 * (out_at + emitted); m[0] = pos - begin; m[1] = depth; m[2] = sp; m[3] = ep; }
      emitted++; extended = false;
    }
    need_parent = true;
  }
}

// scratch (stride matches per pattern) -> CSR; patterns with more than `stride` matches are listed in `overflow`
__global__ void __launch_bounds__(256)
mem_gather_kernel(const ulonglong4* __restrict__ scratch, const u64* __restrict__ counts, const u64* __restrict__ out_offsets,
                  u64 n, u64 stride, ulonglong4* __restrict__ matches, u64* __restrict__ overflow, ull* __restrict__ n_overflow)
{
  for(u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x)
  {
    u64 c = counts[q];
    if(c > stride) { overflow[atomicAdd(n_overflow, 1ull)] = q; continue; }
    const ulonglong4* src = scratch + q * stride;
    ulonglong4* dst = matches + out_offsets[q];
    for(u64 e = 0; e < c; e++) { dst[e] = src[e]; }
  }
}

__global__ void __launch_bounds__(256)
mem_count_overflow_kernel(const u64* __restrict__ counts, u64 n, u64 stride, ull* __restrict__ n_overflow)
{
  ull mine = 0;
  for(u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x) { mine += (counts[q] > stride ? 1 : 0); }
  for(int d = 16; d > 0; d >>= 1) { mine += __shfl_down_sync(0xFFFFFFFFu, mine, d); }
  if((threadIdx.x & 31) == 0 && mine > 0) { atomicAdd(n_overflow, mine); }
}

//------------------------------------------------------------------------------
// Host side: handles
//------------------------------------------------------------------------------

struct gcsa_b200_index
{
  int device = 0;
  int sm_count = 148;
  DevView view;
  std::vector<void*> allocations;
  u64 device_bytes = 0;
  gcsa_flat_index header;            // scalars only (pointers nulled)

  // Host-side 2-bit packing of fixed-length patterns (pack.cpp): byte -> comp - 1 or 0xFF, and
  // whether that table is exactly ACGT / acgt.  Pinned staging buffers are pooled per handle.
  u8 pack_code[256];
  bool pack_default = false;
  mutable std::mutex pool_mutex;
  mutable std::vector<std::pair<void*, size_t>> pinned_pool;

  void* takePinned(size_t bytes) const
  {
    {
      std::lock_guard<std::mutex> lock(pool_mutex);
      for(size_t i = 0; i < pinned_pool.size(); i++)
      {
        if(pinned_pool[i].second >= bytes)
        {
          void* p = pinned_pool[i].first;
          pinned_pool.erase(pinned_pool.begin() + i);
          return p;
        }
      }
    }
    void* p = nullptr;
    if(cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    sizes_add(p, bytes);
    return p;
  }
  void givePinned(void* p) const
  {
    std::lock_guard<std::mutex> lock(pool_mutex);
    size_t bytes = 0;
    for(auto& e : pinned_sizes) { if(e.first == p) { bytes = e.second; } }
    pinned_pool.push_back(std::make_pair(p, bytes));
  }
  void sizes_add(void* p, size_t bytes) const { std::lock_guard<std::mutex> lock(pool_mutex); pinned_sizes.push_back(std::make_pair(p, bytes)); }
  mutable std::vector<std::pair<void*, size_t>> pinned_sizes;     // every pinned buffer ever allocated for this handle
};

struct gcsa_b200_lcp
{
  int device = 0;
  int sm_count = 148;
  LcpView view;
  void* data = nullptr;
};

namespace {

struct HostBits
{
  const uint64_t* words; u64 n_bits;
  inline u64 word(u64 w) const
  {
    u64 n_words = (n_bits + 63) / 64;
    if(words == nullptr || w >= n_words) { return 0; }
    u64 x = words[w];
    u64 rem = n_bits - w * 64;
    if(rem < 64) { x &= ((1ull << rem) - 1); }
    return x;
  }
  // up to 64 bits starting at bit `start` (bits past the end read as 0)
  inline u64 get(u64 start, u32 len) const
  {
    if(len == 0) { return 0; }
    u64 w = start >> 6; u32 off = start & 63;
    u64 x = word(w) >> off;
    if(off && off + len > 64) { x |= word(w + 1) << (64 - off); }
    if(len < 64) { x &= ((1ull << len) - 1); }
    return x;
  }
};

// ones before each word; cum[n_words] = total
std::vector<u64> wordCum(const HostBits& b)
{
  u64 n_words = (b.n_bits + 63) / 64;
  std::vector<u64> cum(n_words + 2, 0);
  for(u64 w = 0; w < n_words; w++) { cum[w + 1] = cum[w] + __builtin_popcountll(b.word(w)); }
  cum[n_words + 1] = cum[n_words];
  return cum;
}

inline u64 hostRank(const HostBits& b, const std::vector<u64>& cum, u64 i)
{
  if(i >= b.n_bits) { return cum[(b.n_bits + 63) / 64]; }
  u64 w = i >> 6; u32 r = i & 63;
  return cum[w] + (r ? __builtin_popcountll(b.word(w) & ((1ull << r) - 1)) : 0);
}

int upload(gcsa_b200_index* idx, const void* host, size_t bytes, const void** dev)
{
  void* p = nullptr;
  size_t alloc = std::max<size_t>(bytes, 256);
  CUDA_TRY(cudaMalloc(&p, alloc));
  idx->allocations.push_back(p);
  idx->device_bytes += alloc;
  if(bytes) { CUDA_TRY(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice)); }
  *dev = p;
  return 0;
}

int buildRankVec(gcsa_b200_index* idx, const HostBits& b, RankVecDev* out, std::vector<ulonglong4>* keep = nullptr)
{
  u64 n_sec = b.n_bits / RV_W + 1;
  std::vector<ulonglong4> sec(n_sec);
  u64 cum = 0;
  for(u64 s = 0; s < n_sec; s++)
  {
    ulonglong4 q;
    q.x = cum; q.y = b.word(3 * s); q.z = b.word(3 * s + 1); q.w = b.word(3 * s + 2);
    cum += __builtin_popcountll(q.y) + __builtin_popcountll(q.z) + __builtin_popcountll(q.w);
    sec[s] = q;
  }
  const void* d = nullptr;
  int rc = upload(idx, sec.data(), sec.size() * sizeof(ulonglong4), &d);
  if(rc) { return rc; }
  out->sec = (const ulonglong4*)d; out->n_bits = b.n_bits; out->n_sec = n_sec;
  if(keep) { keep->swap(sec); }
  return 0;
}

int buildSelVec(gcsa_b200_index* idx, const HostBits& b, SelVecDev* out)
{
  std::vector<ulonglong4> sec;
  int rc = buildRankVec(idx, b, &out->rv, &sec);
  if(rc) { return rc; }
  u64 ones = 0;
  if(!sec.empty())
  {
    const ulonglong4& q = sec.back();
    ones = q.x + __builtin_popcountll(q.y) + __builtin_popcountll(q.z) + __builtin_popcountll(q.w);
  }
  u64 n_hints = ones / SEL_HINT + 2;
  std::vector<u32> hints(n_hints, (u32)(sec.size() - 1));
  u64 h = 0;
  for(u64 s = 0; s < sec.size() && h < n_hints; s++)
  {
    const ulonglong4& q = sec[s];
    u64 end = q.x + __builtin_popcountll(q.y) + __builtin_popcountll(q.z) + __builtin_popcountll(q.w);
    while(h < n_hints && h * SEL_HINT + 1 <= end) { if(h * SEL_HINT + 1 > q.x) { hints[h] = (u32)s; } h++; }
  }
  const void* d = nullptr;
  rc = upload(idx, hints.data(), hints.size() * sizeof(u32), &d);
  if(rc) { return rc; }
  out->hints = (const u32*)d; out->ones = ones;
  return 0;
}

inline int gridFor(u64 n, int sm_count, int per_sm = 8)
{
  u64 blocks = (n + 255) / 256;
  u64 cap = (u64)sm_count * per_sm;
  return (int)std::max<u64>(1, std::min(blocks, cap));
}

struct DeviceGuard
{
  int prev = 0; bool ok = false;
  explicit DeviceGuard(int device) { ok = (cudaGetDevice(&prev) == cudaSuccess) && (cudaSetDevice(device) == cudaSuccess); }
  ~DeviceGuard() { if(ok) { cudaSetDevice(prev); } }
};

} // namespace

//------------------------------------------------------------------------------
// C ABI
//------------------------------------------------------------------------------

const char* gcsa_b200_last_error(void) { return g_last_error.c_str(); }
void gcsa_b200_internal_set_error(const char* message) { g_last_error = (message != nullptr ? message : ""); }
const char* gcsa_b200_version(void) { return "gcsa2_b200 0.1 (sm_100a; GCSA v3 / LCP v1 semantics of gcsa2 1.3.0)"; }

int gcsa_b200_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void gcsa_b200_free(void* p) { std::free(p); }

int gcsa_b200_index_create(const gcsa_flat_index* host, int device, const gcsa_b200_options* options, gcsa_b200_index** out)
{
  if(host == nullptr || out == nullptr) { return fail(GCSA_B200_ERR_INVALID, "index_create: null argument"); }
  *out = nullptr;
  if(host->sigma != GCSA_B200_SIGMA || host->fast_chars != GCSA_B200_FAST_CHARS)
  {
    return fail(GCSA_B200_ERR_INVALID, "index_create: only the default alphabet (sigma 7, 4 fast characters) is supported");
  }
  if(host->path_nodes >= (1ull << 40) || host->edge_count >= (1ull << 40))
  {
    return fail(GCSA_B200_ERR_INVALID, "index_create: more than 2^40 path nodes or edges");
  }
  int n_dev = gcsa_b200_device_count();
  if(n_dev <= 0) { return fail(GCSA_B200_ERR_CUDA, "index_create: no CUDA device available (this engine has no CPU fallback)"); }
  if(device < 0 || device >= n_dev) { return fail(GCSA_B200_ERR_INVALID, "index_create: bad device ordinal"); }
  DeviceGuard guard(device);
  if(!guard.ok) { return fail(GCSA_B200_ERR_CUDA, "index_create: cudaSetDevice failed"); }

  {
    cudaMemPool_t pool;
    if(cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
    {
      uint64_t keep = ~0ull;      // do not hand stream-ordered allocations back to the OS between calls
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if(const char* g = std::getenv("GCSA_B200_L2_FETCH"))
  {
    // Random 32-byte sector probes: ask the L2 not to over-fetch neighbouring sectors from HBM.
    size_t bytes = (size_t)std::atoi(g);
    if(bytes == 32 || bytes == 64 || bytes == 128) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, bytes); }
  }
  gcsa_b200_index* idx = new gcsa_b200_index();
  idx->device = device;
  cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, device);
  idx->header = *host;
  for(int c = 0; c < GCSA_B200_SIGMA; c++) { idx->header.bwt[c] = nullptr; }
  idx->header.edges = idx->header.sampled_paths = idx->header.stored_samples = idx->header.samples = nullptr;
  idx->header.extra_filter = idx->header.extra_values = idx->header.redundant = nullptr;

  DevView& v = idx->view;
  std::memset(&v, 0, sizeof(v));
  const u64 N = host->path_nodes;
  v.path_nodes = N; v.edge_count = host->edge_count;
  for(int c = 0; c <= GCSA_B200_SIGMA; c++) { v.C[c] = host->C[c]; }
  std::memcpy(v.char2comp, host->char2comp, 256);
  {
    u8 def[256]; gcsa_b200_default_char2comp(def);
    idx->pack_default = true;
    for(int i = 0; i < 256; i++)
    {
      u8 c = host->char2comp[i];
      bool fast = (c >= 1 && c <= GCSA_B200_FAST_CHARS);
      idx->pack_code[i] = (fast ? (u8)(c - 1) : (u8)0xFF);
      bool def_fast = (def[i] >= 1 && def[i] <= GCSA_B200_FAST_CHARS);
      if(fast != def_fast || (fast && c != def[i])) { idx->pack_default = false; }
    }
    v.default_alphabet = (idx->pack_default ? 1u : 0u);
  }
  for(int i = 0; i < 256; i++) { if(v.char2comp[i] >= GCSA_B200_SIGMA) { delete idx; return fail(GCSA_B200_ERR_INVALID, "index_create: char2comp value out of range"); } }

  int rc = 0;
  #define TRY_RC(expr) do { rc = (expr); if(rc) { gcsa_b200_index_destroy(idx); return rc; } } while(0)

  HostBits edges = { host->edges, host->edge_count };
  std::vector<u64> edge_cum = wordCum(edges);
  TRY_RC(buildRankVec(idx, edges, &v.edges));

  // charRange(comp) = pathNodeRange(C[comp], C[comp+1] - 1), gcsa.h:150-153; C[comp+1] == 0 -> (0, ~0)
  for(int c = 0; c < GCSA_B200_SIGMA; c++)
  {
    if(host->C[c + 1] == 0) { v.char_sp[c] = 0; v.char_ep[c] = ~0ull; }
    else { v.char_sp[c] = hostRank(edges, edge_cum, host->C[c]); v.char_ep[c] = hostRank(edges, edge_cum, host->C[c + 1] - 1); }
  }

  // fused BWT blocks
  {
    u64 n_blocks = N / BWT_W + 1;
    std::vector<ulonglong4> blocks(n_blocks * 4);
    for(int c = 1; c <= GCSA_B200_FAST_CHARS; c++)
    {
      HostBits B = { host->bwt[c], N };
      std::vector<u64> cum = wordCum(B);
      const u64 Cc = host->C[c];
      #pragma omp parallel for schedule(static)
      for(long long bb = 0; bb < (long long)n_blocks; bb++)
      {
        u64 b = (u64)bb, start = b * BWT_W;
        u64 cnt = hostRank(B, cum, start);
        u64 P = Cc + cnt;
        u64 blo = B.get(start, 64), bhi = B.get(start + 64, BWT_W - 64);
        u64 e0, wlo, whi;
        if(P == 0) { e0 = 0; wlo = edges.get(0, 63) << 1; whi = edges.get(63, 24); }
        else { e0 = hostRank(edges, edge_cum, P - 1); wlo = edges.get(P - 1, 64); whi = edges.get(P - 1 + 64, 24); }
        ulonglong4 q;
        q.x = (P & M40) | (bhi << 40); q.y = blo;
        q.z = (e0 & M40) | (whi << 40); q.w = wlo;
        blocks[b * 4 + (c - 1)] = q;
      }
    }
    const void* d = nullptr;
    TRY_RC(upload(idx, blocks.data(), blocks.size() * sizeof(ulonglong4), &d));
    v.bwt = (const ulonglong4*)d;
  }

  // sparse characters
  {
    const int comps[3] = { 0, 5, 6 };
    for(int s = 0; s < 3; s++)
    {
      HostBits B = { host->bwt[comps[s]], N };
      std::vector<u64> pos;
      u64 n_words = (N + 63) / 64;
      for(u64 w = 0; w < n_words; w++)
      {
        u64 x = B.word(w);
        while(x) { pos.push_back(w * 64 + __builtin_ctzll(x)); x &= x - 1; }
      }
      const void* d = nullptr;
      TRY_RC(upload(idx, pos.data(), pos.size() * sizeof(u64), &d));
      v.sparse_pos[s] = (const u64*)d; v.sparse_n[s] = pos.size();
    }
  }

  // samples
  {
    HostBits sampled = { host->sampled_paths, N };
    TRY_RC(buildRankVec(idx, sampled, &v.sampled));
    HostBits last = { host->samples, host->sample_count };
    std::vector<u64> start; start.push_back(0);
    u64 n_words = (host->sample_count + 63) / 64;
    for(u64 w = 0; w < n_words; w++)
    {
      u64 x = last.word(w);
      while(x) { start.push_back(w * 64 + __builtin_ctzll(x) + 1); x &= x - 1; }
    }
    start.push_back(host->sample_count);   // guard entry
    const void* d = nullptr;
    TRY_RC(upload(idx, start.data(), start.size() * sizeof(u64), &d));
    v.sample_start = (const u64*)d;
    TRY_RC(upload(idx, host->stored_samples, host->sample_count * sizeof(u64), &d));
    v.stored_samples = (const u64*)d; v.sample_count = host->sample_count;
  }

  // counting structures
  {
    HostBits filter = { host->extra_filter, N };
    TRY_RC(buildRankVec(idx, filter, &v.extra_filter));
    HostBits values = { host->extra_values, host->extra_values_len };
    TRY_RC(buildSelVec(idx, values, &v.extra_values));
    HostBits red = { host->redundant, host->redundant_len };
    TRY_RC(buildSelVec(idx, red, &v.redundant));
  }

  // locate walk table (optional; 4 or 8 bytes per path node)
  if(N > 0 && (options == nullptr || options->walk_table != 0))
  {
    bool narrow = (N < (1ull << 31));
    size_t bytes = (size_t)N * (narrow ? sizeof(u32) : sizeof(u64));
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    bool forced = (options != nullptr && options->walk_table > 0);
    if(forced || bytes < free_b / 4)
    {
      void* p = nullptr;
      cudaError_t e = cudaMalloc(&p, bytes);
      if(e == cudaSuccess)
      {
        if(narrow) { walk_table_kernel<u32><<<gridFor(N, idx->sm_count, 8), 256>>>(v, (u32*)p); }
        else { walk_table_kernel<u64><<<gridFor(N, idx->sm_count, 8), 256>>>(v, (u64*)p); }
        e = cudaDeviceSynchronize();
      }
      if(e != cudaSuccess)
      {
        if(p) { cudaFree(p); }
        gcsa_b200_index_destroy(idx);
        return fail(GCSA_B200_ERR_CUDA, std::string("walk table: ") + cudaGetErrorString(e));
      }
      idx->allocations.push_back(p); idx->device_bytes += bytes;
      if(narrow) { v.walk32 = (const u32*)p; } else { v.walk64 = (const u64*)p; }

      // Locate table (8 bytes per path node) from the walk table, which it then replaces: one load per
      // located node instead of one per LF step.  walk_table = 2 keeps the walk table instead.
      size_t loc_bytes = (size_t)N * sizeof(u64);
      cudaMemGetInfo(&free_b, &total_b);
      if((options == nullptr || options->walk_table != 2) && (forced || loc_bytes < free_b / 2))
      {
        u64* loc = nullptr; int* overflow = nullptr; int host_overflow = 0;
        e = cudaMalloc((void**)&loc, loc_bytes);
        if(e == cudaSuccess) { e = cudaMalloc((void**)&overflow, sizeof(int)); }
        if(e == cudaSuccess) { e = cudaMemset(overflow, 0, sizeof(int)); }
        if(e == cudaSuccess)
        {
          locate_table_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(v, loc, overflow);
          e = cudaMemcpy(&host_overflow, overflow, sizeof(int), cudaMemcpyDeviceToHost);
        }
        if(overflow) { cudaFree(overflow); }
        if(e == cudaSuccess && host_overflow == 0)
        {
          cudaFree(p); idx->allocations.pop_back(); idx->device_bytes -= bytes;
          v.walk32 = nullptr; v.walk64 = nullptr;
          idx->allocations.push_back(loc); idx->device_bytes += loc_bytes;
          v.loc64 = loc;
        }
        else
        {
          if(loc) { cudaFree(loc); }
          cudaGetLastError();               // keep the walk table
        }
      }
    }
  }

  // jump table (optional; 8 bytes per path node + as much again while it is built)
  {
    int want = (options != nullptr ? options->jump_table : 0);        // 0 = automatic, 1 = build, -1 = do not
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t bytes = (size_t)N * sizeof(u64);
    u32 tbits = 1; while((1ull << tbits) < N) { tbits++; }
    int max_len = std::min<int>(16, (59 - (int)tbits) / 2);
    if(N > 0 && want >= 0 && max_len >= 2 && (want > 0 || 2 * bytes < free_b / 2))
    {
      u64 *one = nullptr, *table = nullptr, *short_table = nullptr;
      const int short_len = 4;
      cudaError_t e = cudaMalloc((void**)&one, bytes);
      if(e == cudaSuccess) { e = cudaMalloc((void**)&table, bytes); }
      if(e == cudaSuccess && max_len > short_len && 3 * bytes < free_b / 2) { if(cudaMalloc((void**)&short_table, bytes) != cudaSuccess) { short_table = nullptr; cudaGetLastError(); } }
      if(e == cudaSuccess)
      {
        jump_init_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(v, tbits, one, table);
        for(int j = 1; j < max_len; j++)
        {
          if(j == short_len && short_table != nullptr) { e = cudaMemcpyAsync(short_table, table, bytes, cudaMemcpyDeviceToDevice, 0); }   // paths of up to 4 steps
          jump_extend_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, (u32)j, one, table);
        }
        if(e == cudaSuccess) { e = cudaDeviceSynchronize(); }
      }
      if(one) { cudaFree(one); }
      if(e == cudaSuccess)
      {
        idx->allocations.push_back(table); idx->device_bytes += bytes;
        v.jump = table; v.jump_k = (u32)max_len; v.jump_tbits = tbits;
        if(short_table != nullptr) { idx->allocations.push_back(short_table); idx->device_bytes += bytes; v.jump_short = short_table; }
      }
      else
      {
        if(table) { cudaFree(table); }
        if(short_table) { cudaFree(short_table); }
        cudaGetLastError();
        if(want > 0) { gcsa_b200_index_destroy(idx); return fail(GCSA_B200_ERR_CUDA, std::string("jump table: ") + cudaGetErrorString(e)); }
      }
    }
  }

  // two-step blocks (optional): built on the device from the one-step blocks
  // -1 = automatic: worth it once the one-step blocks are far beyond the L2 (the probe rate no longer
  // depends on the footprint there, so halving the probes halves the time; measured in DESIGN.md)
  bool want_two_step = (options != nullptr && (options->two_step > 0 || (options->two_step < 0 && N >= 400000000ull)));
  if(want_two_step && N > 0)
  {
    u64 n_blocks = N / BWT_W + 1;
    unsigned short* m2 = nullptr; u32* blockpop = nullptr; u64* blockcnt = nullptr; u64* d_base = nullptr; u64* src = nullptr;
    u32* d_viol = nullptr; void* scan_tmp = nullptr; ulonglong4* blocks2 = nullptr;
    cudaError_t e = cudaSuccess;
    auto cleanup2 = [&]() { cudaFree(m2); cudaFree(blockpop); cudaFree(blockcnt); cudaFree(d_base); cudaFree(src); cudaFree(d_viol); cudaFree(scan_tmp); };
    #define TWO_TRY(expr) do { e = (expr); if(e != cudaSuccess) { cleanup2(); if(blocks2) { cudaFree(blocks2); } gcsa_b200_index_destroy(idx); \
      return fail(GCSA_B200_ERR_CUDA, std::string("two-step build: " #expr ": ") + cudaGetErrorString(e)); } } while(0)
    TWO_TRY(cudaMalloc(&m2, (N + 1) * sizeof(unsigned short)));
    TWO_TRY(cudaMalloc(&blockpop, 16 * n_blocks * sizeof(u32)));
    TWO_TRY(cudaMalloc(&blockcnt, 16 * n_blocks * sizeof(u64)));
    TWO_TRY(cudaMalloc(&d_base, 17 * sizeof(u64)));
    TWO_TRY(cudaMalloc(&d_viol, sizeof(u32)));
    TWO_TRY(cudaMemset(d_viol, 0, sizeof(u32)));
    two_step_mask_kernel<<<gridFor(n_blocks, idx->sm_count, 16), 128>>>(v, n_blocks, m2, blockpop);
    TWO_TRY(cudaGetLastError());
    size_t scan_bytes = 0;
    TWO_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, blockpop, blockcnt, n_blocks));
    TWO_TRY(cudaMalloc(&scan_tmp, std::max<size_t>(scan_bytes, 16)));
    u64 base[17]; base[0] = 0;
    for(int p = 0; p < 16; p++)
    {
      TWO_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, blockpop + (u64)p * n_blocks, blockcnt + (u64)p * n_blocks, n_blocks));
      u64 last_cnt = 0; u32 last_pop = 0;
      TWO_TRY(cudaMemcpy(&last_cnt, blockcnt + (u64)p * n_blocks + (n_blocks - 1), sizeof(u64), cudaMemcpyDeviceToHost));
      TWO_TRY(cudaMemcpy(&last_pop, blockpop + (u64)p * n_blocks + (n_blocks - 1), sizeof(u32), cudaMemcpyDeviceToHost));
      base[p + 1] = base[p] + last_cnt + last_pop;
    }
    TWO_TRY(cudaMemcpy(d_base, base, sizeof(base), cudaMemcpyHostToDevice));
    TWO_TRY(cudaMalloc(&src, std::max<u64>(base[16], 1) * sizeof(u64)));
    two_step_source_kernel<<<gridFor(n_blocks, idx->sm_count, 16), 128>>>(v, n_blocks, m2, blockcnt, d_base, src);
    two_step_validate_kernel<<<gridFor(base[16] / 16 + 1, idx->sm_count, 8), 256>>>(src, d_base, d_viol);
    u32 violations = 0;
    TWO_TRY(cudaMemcpy(&violations, d_viol, sizeof(u32), cudaMemcpyDeviceToHost));
    if(violations == 0)
    {
      TWO_TRY(cudaMalloc(&blocks2, n_blocks * 16 * sizeof(ulonglong4)));
      two_step_build_kernel<<<gridFor(n_blocks * 16, idx->sm_count, 8), 256>>>(N, n_blocks, m2, blockcnt, d_base, src, blocks2);
      TWO_TRY(cudaDeviceSynchronize());
      idx->allocations.push_back(blocks2); idx->device_bytes += n_blocks * 16 * sizeof(ulonglong4);
      v.bwt2 = blocks2;
    }
    cleanup2();
    #undef TWO_TRY
  }

  // k-mer table
  int k = (options ? options->kmer_table_k : 0);
  if(k < 0) { k = 0; }
  if(k > 16) { k = 16; }
  if(k > 0 && N > 0)
  {
    u64 entries = 1ull << (2 * k);
    u64 tmp_entries = (k == 1 ? 4 : 1ull << (2 * (k - 1)));
    // Fused form (16 bytes per entry: the jump entry of a singleton result rides along) when there is a jump table
    // and the doubled table still leaves most of the device free; fused_table = 1 forces it, -1 forbids it.
    int want_fused = (options != nullptr ? options->fused_table : 0);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    bool fused = (v.jump != nullptr && want_fused >= 0 &&
                  (want_fused > 0 || (entries + tmp_entries) * sizeof(ulonglong2) < free_b / 10 * 6));
    size_t entry_bytes = (fused ? sizeof(ulonglong2) : sizeof(u64));
    void* p = nullptr; void* tmp = nullptr;
    cudaError_t e = cudaMalloc(&p, entries * entry_bytes);
    if(e == cudaSuccess) { e = cudaMalloc(&tmp, tmp_entries * sizeof(ulonglong2)); }
    if(e != cudaSuccess)
    {
      if(p) { cudaFree(p); }
      gcsa_b200_index_destroy(idx);
      return fail(GCSA_B200_ERR_NOMEM, "index_create: k-mer table allocation failed");
    }
    idx->allocations.push_back(p); idx->device_bytes += entries * entry_bytes;
    table_init_kernel<<<1, 256>>>(v, (ulonglong2*)tmp);
    for(int j = 1; j + 1 < k; j++) { table_extend_kernel<<<gridFor(1ull << (2 * j), idx->sm_count, 8), 256>>>(v, j, (ulonglong2*)tmp); }
    table_final_kernel<<<gridFor(tmp_entries, idx->sm_count, 8), 256>>>(v, k, (const ulonglong2*)tmp, fused ? nullptr : (u64*)p, fused ? (ulonglong2*)p : nullptr);
    e = cudaDeviceSynchronize();
    cudaFree(tmp);
    if(e != cudaSuccess) { gcsa_b200_index_destroy(idx); return fail(GCSA_B200_ERR_CUDA, std::string("k-mer table kernels: ") + cudaGetErrorString(e)); }
    if(fused) { v.table2 = (const ulonglong2*)p; } else { v.table = (const u64*)p; }
    v.table_k = k;
  }
  #undef TRY_RC

  *out = idx;
  return 0;
}

void gcsa_b200_index_destroy(gcsa_b200_index* index)
{
  if(index == nullptr) { return; }
  DeviceGuard guard(index->device);
  for(void* p : index->allocations) { cudaFree(p); }
  for(auto& e : index->pinned_sizes) { cudaFreeHost(e.first); }
  delete index;
}

int gcsa_b200_index_info(const gcsa_b200_index* index, gcsa_b200_info* info)
{
  if(index == nullptr || info == nullptr) { return fail(GCSA_B200_ERR_INVALID, "index_info: null argument"); }
  std::memset(info, 0, sizeof(*info));
  info->path_nodes = index->header.path_nodes; info->edge_count = index->header.edge_count;
  info->order = index->header.order; info->sample_count = index->header.sample_count;
  info->device_bytes = index->device_bytes; info->kmer_table_k = index->view.table_k;
  info->device = index->device; info->sm_count = index->sm_count;
  info->two_step = (index->view.bwt2 != nullptr ? 1 : 0);
  info->jump_k = (index->view.jump != nullptr ? (int)index->view.jump_k : 0);
  info->fused_table = (index->view.table2 != nullptr ? 1 : 0);
  return 0;
}

int gcsa_b200_char_range(const gcsa_b200_index* index, uint64_t comp, uint64_t* sp, uint64_t* ep)
{
  if(index == nullptr || sp == nullptr || ep == nullptr) { return fail(GCSA_B200_ERR_INVALID, "char_range: null argument"); }
  if(comp >= GCSA_B200_SIGMA) { *sp = 0; *ep = ~0ull; return 0; }
  *sp = index->view.char_sp[comp]; *ep = index->view.char_ep[comp];
  return 0;
}

//------------------------------------------------------------------------------
// find
//------------------------------------------------------------------------------

static int launchFind(const gcsa_b200_index* index, const u8* d_chars, const u64* d_offsets, u64 char_base, u64 fixed_length,
                      u64 n, u64* d_sp, u64* d_ep, FindStatsDev* d_stats, cudaStream_t stream, bool packed = false)
{
  if(n == 0) { return 0; }
  // persistent grid: 4 CTAs of 256 threads per SM by default (59 registers, no spills; with the packed pattern
  // tail 5 CTAs/SM spill: 13.3 vs 13.1 G queries/s with the 16-mer table, but 7.0 vs 8.4 with the 14-mer table,
  // where more single steps run), one contiguous slice of queries per warp
  static const int min_blocks = []() { const char* e = std::getenv("GCSA_B200_FIND_MINBLOCKS"); return (e ? std::atoi(e) : 4); }();
  int per_sm = (min_blocks >= 8 ? 8 : (min_blocks <= 4 ? 4 : min_blocks));
  int grid = gridFor(n, index->sm_count, per_sm);
  // idle lanes of a warp are refilled together once this many are idle (16 measured best: DESIGN.md)
  static const int refill_at = []() { const char* e = std::getenv("GCSA_B200_FIND_REFILL"); int r = (e ? std::atoi(e) : 16); return std::min(32, std::max(1, r)); }();
  #define LAUNCH_FIND(S, B) find_kernel<S, B><<<grid, 256, 0, stream>>>(index->view, d_chars, d_offsets, char_base, fixed_length, n, d_sp, d_ep, d_stats, refill_at)
  if(packed)
  {
    find_kernel<false, 4, true><<<gridFor(n, index->sm_count, 4), 256, 0, stream>>>(index->view, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, nullptr, refill_at);
  }
  else if(d_stats) { LAUNCH_FIND(true, 1); }
  else if(per_sm == 8) { LAUNCH_FIND(false, 8); }
  else if(per_sm == 5) { LAUNCH_FIND(false, 5); }
  else if(per_sm == 4) { LAUNCH_FIND(false, 4); }
  else { LAUNCH_FIND(false, 6); }
  #undef LAUNCH_FIND
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_find_batch(const gcsa_b200_index* index, const uint8_t* d_chars, const uint64_t* d_offsets,
                         uint64_t n, uint64_t* d_sp, uint64_t* d_ep, void* stream)
{
  if(index == nullptr || (n > 0 && (d_chars == nullptr || d_offsets == nullptr || d_sp == nullptr || d_ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_batch: null argument");
  }
  DeviceGuard guard(index->device);
  return launchFind(index, d_chars, (const u64*)d_offsets, 0, 0, n, (u64*)d_sp, (u64*)d_ep, nullptr, (cudaStream_t)stream);
}

int gcsa_b200_find_fixed_batch(const gcsa_b200_index* index, const uint8_t* d_chars, uint64_t pattern_length,
                               uint64_t n, uint64_t* d_sp, uint64_t* d_ep, void* stream)
{
  if(index == nullptr || (n > 0 && (d_chars == nullptr || d_sp == nullptr || d_ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_fixed_batch: null argument");
  }
  DeviceGuard guard(index->device);
  return launchFind(index, d_chars, nullptr, 0, pattern_length, n, (u64*)d_sp, (u64*)d_ep, nullptr, (cudaStream_t)stream);
}

/*
  Host-buffer find: the batch is cut into chunks that are pipelined over three streams
  (H2D of chunk i+1 overlaps the kernel of chunk i and the D2H of chunk i-1).
*/
// Host-side 2-bit packing in the host entry point of find() (pack.cpp): 4x fewer bytes over PCIe, which is what
// bounds that entry point -- but only a gain when the host packs faster than the link moves the raw bytes
// (measured on a 16-vCPU B200 box with the first packer: break-even at 16 threads).  Policy:
//   GCSA_B200_HOST_PACK=0     never;   =N (> 0)  always, with N OpenMP threads;
//   unset or "auto"           decided by measurement: the first large batch is packed with all OpenMP threads
//                             (GCSA_B200_HOST_PACK_THREADS overrides the count) while the packing rate of its first
//                             two chunks is timed; packing stays on iff the better of the two reaches
//                             GCSA_B200_HOST_PACK_MIN_GBS (default 55 GB/s of pattern bytes: the ~50 GB/s the link
//                             sustains plus a margin).  The decision is kept for the process.
enum { PACK_UNKNOWN = -1, PACK_OFF = 0, PACK_ON = 1 };
static std::atomic<int> g_pack_auto(PACK_UNKNOWN);
struct PackPolicy { int threads; bool calibrate; };

static PackPolicy hostPackPolicy()
{
  const char* e = std::getenv("GCSA_B200_HOST_PACK");
  if(e != nullptr && *e != 0 && std::strcmp(e, "auto") != 0)
  {
    int v = std::atoi(e);
    return PackPolicy{ (v < 0 ? 0 : v), false };
  }
  int state = g_pack_auto.load();
  if(state == PACK_OFF) { return PackPolicy{ 0, false }; }
  const char* t = std::getenv("GCSA_B200_HOST_PACK_THREADS");
  int threads = (t != nullptr && *t != 0 ? std::atoi(t) : omp_get_max_threads());
  return PackPolicy{ std::max(1, threads), state == PACK_UNKNOWN };
}

static double hostPackMinRate()
{
  const char* e = std::getenv("GCSA_B200_HOST_PACK_MIN_GBS");
  return (e != nullptr && *e != 0 ? std::atof(e) : 55.0) * 1e9;
}

// Test / measurement hooks (not part of the C ABI): forget the automatic decision; report it.
extern "C" void gcsa_b200_internal_pack_reset(void) { g_pack_auto.store(PACK_UNKNOWN); }
extern "C" int gcsa_b200_internal_pack_state(void) { return g_pack_auto.load(); }

static int findHost(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets, uint64_t fixed_length,
                    uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats)
{
  if(index == nullptr || (n > 0 && (chars == nullptr || sp == nullptr || ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_host: null argument");
  }
  if(stats) { std::memset(stats, 0, sizeof(*stats)); stats->queries = n; }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);

  // Chunks of >= 256 k queries (8 MB of 32-mers: PCIe is at its streaming rate), at most ~24 per batch: the H2D
  // engine is the busy resource from the first byte on, so what the pipeline adds to the transfer time is the kernel
  // and the D2H of the LAST chunk -- the smaller the chunks, the smaller that tail (8 chunks: +0.45 ms on 6.4 ms).
  // With host cores to spare, fixed-length batches are 2-bit packed on the host first (pack.cpp):
  // 4x fewer bytes over PCIe, which is what bounds this entry point; a chunk containing any character
  // other than ACGT/acgt goes through the byte path.  Finer chunks then, so that packing chunk i+1
  // overlaps the transfers of chunk i.
  const int STREAMS = 3;
  const PackPolicy policy = hostPackPolicy();
  const int pack_threads = policy.threads;
  const bool pack = (pack_threads > 0 && fixed_length > 0 && offsets == nullptr && stats == nullptr &&
                     n >= (policy.calibrate ? (1u << 20) : (1u << 16)));
  bool calibrating = (pack && policy.calibrate);
  double best_rate = 0.0; int timed_chunks = 0;
  const u64 CHUNK = (pack ? std::max<u64>(1ull << 18, (n + 15) / 16) : std::max<u64>(1ull << 18, (n + 23) / 24));
  const u64 words_per_pattern = (fixed_length + 31) / 32;
  cudaStream_t streams[STREAMS];
  for(int s = 0; s < STREAMS; s++) { CUDA_TRY(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking)); }
  FindStatsDev* d_stats = nullptr;
  if(stats) { CUDA_TRY(cudaMalloc(&d_stats, sizeof(FindStatsDev))); CUDA_TRY(cudaMemset(d_stats, 0, sizeof(FindStatsDev))); }
  u64* staging[STREAMS] = { nullptr, nullptr, nullptr };
  cudaEvent_t staged[STREAMS] = { nullptr, nullptr, nullptr };
  bool use_pack = pack;
  if(use_pack)
  {
    for(int s = 0; s < STREAMS; s++)
    {
      staging[s] = (u64*)index->takePinned(CHUNK * words_per_pattern * sizeof(u64));
      if(staging[s] == nullptr || cudaEventCreateWithFlags(&staged[s], cudaEventDisableTiming) != cudaSuccess) { use_pack = false; }
    }
  }

  int rc = 0;
  u64 n_chunks = (n + CHUNK - 1) / CHUNK;
  for(u64 c = 0; c < n_chunks && rc == 0; c++)
  {
    const int slot = (int)(c % STREAMS);
    cudaStream_t st = streams[slot];
    u64 q0 = c * CHUNK, q1 = std::min(n, q0 + CHUNK), m = q1 - q0;
    u64 c0 = (offsets ? offsets[q0] : q0 * fixed_length), c1 = (offsets ? offsets[q1] : q1 * fixed_length), bytes = c1 - c0;
    bool packed = false;
    if(use_pack)
    {
      if(c >= (u64)STREAMS) { cudaEventSynchronize(staged[slot]); }          // the slot's previous copy has left the buffer
      double t0 = (calibrating ? omp_get_wtime() : 0.0);
      packed = (gcsa_b200_internal_pack_patterns(chars + c0, m, fixed_length, index->pack_code, index->pack_default ? 1 : 0,
                                                 staging[slot], pack_threads) != 0);
      if(calibrating && packed && m == CHUNK)
      {
        double secs = omp_get_wtime() - t0;
        if(secs > 0.0) { best_rate = std::max(best_rate, (double)bytes / secs); }
        if(++timed_chunks == 2)
        {
          calibrating = false;
          bool keep = (best_rate >= hostPackMinRate());
          g_pack_auto.store(keep ? PACK_ON : PACK_OFF);
          if(!keep) { use_pack = false; }                                      // the rest of this batch goes as raw bytes
        }
      }
    }
    if(packed) { bytes = m * words_per_pattern * sizeof(u64); }
    u8* d_chars = nullptr; u64* d_off = nullptr; u64* d_res = nullptr;
    cudaError_t e;
    if((e = cudaMallocAsync(&d_chars, bytes + 16, st)) != cudaSuccess ||
       (offsets && (e = cudaMallocAsync(&d_off, (m + 1) * sizeof(u64), st)) != cudaSuccess) ||
       (e = cudaMallocAsync(&d_res, 2 * m * sizeof(u64), st)) != cudaSuccess)
    { rc = fail(GCSA_B200_ERR_NOMEM, std::string("find_host: ") + cudaGetErrorString(e)); break; }
    if(packed)
    {
      cudaMemcpyAsync(d_chars, staging[slot], bytes, cudaMemcpyHostToDevice, st);
      cudaEventRecord(staged[slot], st);
    }
    else if(bytes) { cudaMemcpyAsync(d_chars, chars + c0, bytes, cudaMemcpyHostToDevice, st); }
    if(offsets) { cudaMemcpyAsync(d_off, offsets + q0, (m + 1) * sizeof(u64), cudaMemcpyHostToDevice, st); }
    rc = launchFind(index, d_chars, d_off, c0, fixed_length, m, d_res, d_res + m, d_stats, st, packed);
    cudaMemcpyAsync(sp + q0, d_res, m * sizeof(u64), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(ep + q0, d_res + m, m * sizeof(u64), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d_chars, st); if(d_off) { cudaFreeAsync(d_off, st); } cudaFreeAsync(d_res, st);
  }
  cudaError_t err = cudaSuccess;
  for(int s = 0; s < STREAMS; s++)
  {
    cudaError_t e = cudaStreamSynchronize(streams[s]);
    if(e != cudaSuccess) { err = e; }
  }
  if(stats && err == cudaSuccess && rc == 0)
  {
    FindStatsDev h;
    err = cudaMemcpy(&h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    stats->found = h.found; stats->total_length = h.total_length; stats->lf_steps = h.lf_steps;
    stats->sector_probes = h.sector_probes; stats->table_hits = h.table_hits;
  }
  if(d_stats) { cudaFree(d_stats); }
  for(int s = 0; s < STREAMS; s++)
  {
    cudaStreamDestroy(streams[s]);
    if(staged[s]) { cudaEventDestroy(staged[s]); }
    if(staging[s]) { index->givePinned(staging[s]); }
  }
  if(rc) { return rc; }
  if(err != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string("find_host: ") + cudaGetErrorString(err)); }
  return 0;
}

int gcsa_b200_find_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                        uint64_t n, uint64_t* sp, uint64_t* ep)
{
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_host: null offsets"); }
  return findHost(index, chars, offsets, 0, n, sp, ep, nullptr);
}

int gcsa_b200_find_fixed_host(const gcsa_b200_index* index, const uint8_t* chars, uint64_t pattern_length,
                              uint64_t n, uint64_t* sp, uint64_t* ep)
{
  return findHost(index, chars, nullptr, pattern_length, n, sp, ep, nullptr);
}

int gcsa_b200_find_stats_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                              uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats)
{
  if(stats == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_stats_host: null stats"); }
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_stats_host: null offsets"); }
  return findHost(index, chars, offsets, 0, n, sp, ep, stats);
}

//------------------------------------------------------------------------------
// Generic host wrapper: copy inputs, run, copy outputs
//------------------------------------------------------------------------------

namespace {

struct Scratch
{
  cudaStream_t stream = nullptr;
  std::vector<void*> ptrs;
  int init() { return (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess ? 0 : -1); }
  template<class T> T* alloc(u64 count)
  {
    void* p = nullptr;
    if(cudaMallocAsync(&p, std::max<u64>(count, 1) * sizeof(T), stream) != cudaSuccess) { return nullptr; }
    ptrs.push_back(p);
    return (T*)p;
  }
  template<class T> T* in(const T* host, u64 count)
  {
    T* p = alloc<T>(count);
    if(p && count) { cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream); }
    return p;
  }
  template<class T> void out(T* host, const T* dev, u64 count)
  {
    if(count) { cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, stream); }
  }
  cudaError_t finish()
  {
    for(void* p : ptrs) { cudaFreeAsync(p, stream); }
    ptrs.clear();
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream); stream = nullptr;
    return e;
  }
};

} // namespace

#define HOST_PROLOGUE(name, handle) \
  if((handle) == nullptr) { return fail(GCSA_B200_ERR_INVALID, name ": null handle"); } \
  DeviceGuard guard((handle)->device); \
  Scratch sc; if(sc.init()) { return fail(GCSA_B200_ERR_CUDA, name ": cannot create stream"); }

#define HOST_EPILOGUE(name, rc) \
  { cudaError_t e_ = sc.finish(); if((rc) != 0) { return (rc); } \
    if(e_ != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string(name ": ") + cudaGetErrorString(e_)); } return 0; }

int gcsa_b200_lf_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                       const uint8_t* d_comp, uint64_t n, uint64_t* d_sp_out, uint64_t* d_ep_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lf_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  lf_kernel<<<gridFor(n, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_sp, (const u64*)d_ep, d_comp, n, (u64*)d_sp_out, (u64*)d_ep_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_lf_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                      const uint8_t* comp, uint64_t n, uint64_t* sp_out, uint64_t* ep_out)
{
  HOST_PROLOGUE("lf_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n); u8* c = sc.in(comp, n);
  u64* oa = sc.alloc<u64>(n); u64* ob = sc.alloc<u64>(n);
  int rc = gcsa_b200_lf_batch(index, a, b, c, n, oa, ob, sc.stream);
  sc.out((u64*)sp_out, oa, n); sc.out((u64*)ep_out, ob, n);
  HOST_EPILOGUE("lf_host", rc);
}

int gcsa_b200_lf_node_batch(const gcsa_b200_index* index, const uint64_t* d_nodes, uint64_t n, uint64_t* d_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lf_node_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  lf_node_kernel<<<gridFor(n, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_nodes, n, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_lf_node_host(const gcsa_b200_index* index, const uint64_t* nodes, uint64_t n, uint64_t* out)
{
  HOST_PROLOGUE("lf_node_host", index);
  u64* a = sc.in((const u64*)nodes, n); u64* o = sc.alloc<u64>(n);
  int rc = gcsa_b200_lf_node_batch(index, a, n, o, sc.stream);
  sc.out((u64*)out, o, n);
  HOST_EPILOGUE("lf_node_host", rc);
}

int gcsa_b200_lf_multi_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                             uint64_t n, int all_chars, uint64_t* d_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lf_multi_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  lf_multi_kernel<<<gridFor(n * GCSA_B200_SIGMA, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_sp, (const u64*)d_ep, n, all_chars, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_lf_multi_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                            uint64_t n, int all_chars, uint64_t* out)
{
  HOST_PROLOGUE("lf_multi_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  u64* o = sc.alloc<u64>(n * GCSA_B200_SIGMA * 2);
  int rc = gcsa_b200_lf_multi_batch(index, a, b, n, all_chars, o, sc.stream);
  sc.out((u64*)out, o, n * GCSA_B200_SIGMA * 2);
  HOST_EPILOGUE("lf_multi_host", rc);
}

int gcsa_b200_count_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                          uint64_t n, uint64_t* d_out, void* stream)
{
  if(index == nullptr) { return fail(GCSA_B200_ERR_INVALID, "count_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);
  count_kernel<<<gridFor(n, index->sm_count), 256, 0, (cudaStream_t)stream>>>(index->view, (const u64*)d_sp, (const u64*)d_ep, n, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_count_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n, uint64_t* out)
{
  HOST_PROLOGUE("count_host", index);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n); u64* o = sc.alloc<u64>(n);
  int rc = gcsa_b200_count_batch(index, a, b, n, o, sc.stream);
  sc.out((u64*)out, o, n);
  HOST_EPILOGUE("count_host", rc);
}

//------------------------------------------------------------------------------
// locate
//------------------------------------------------------------------------------

namespace {

template<class T> int scanExclusive(const T* in, T* out, u64 count, cudaStream_t st)
{
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, st));
  void* tmp = nullptr;
  CUDA_TRY(cudaMallocAsync(&tmp, std::max<size_t>(bytes, 16), st));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, st);
  cudaFreeAsync(tmp, st);
  CUDA_TRY(e);
  return 0;
}

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
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(cudaMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { return nullptr; } tmp.push_back(p); return p; };
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
      LOC_TRY(cudaMallocAsync(&p, std::max<u64>(total, 1) * sizeof(u64), st));
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
    LOC_TRY(cudaMallocAsync(&p, std::max<u64>(distinct, 1) * sizeof(u64), st));
    *d_values_alloc = (u64*)p; d_values = (u64*)p; capacity = distinct;
  }
  if(d_values == nullptr || capacity < distinct) { rc = GCSA_B200_ERR_CAPACITY; g_last_error = "locate: output capacity too small"; }
  else { locate_compact_kernel<<<gridFor(total, sm), 256, 0, st>>>(sorted, flag, flag_scan, total, d_values, capacity); }
  LOC_TRY(cudaGetLastError());
  cleanup();
  return rc;
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
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(cudaMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { return nullptr; } tmp.push_back(p); return p; };
  auto cleanup = [&]() { for(void* p : tmp) { cudaFreeAsync(p, st); } tmp.clear(); if(gvals) { cudaFreeAsync(gvals, st); gvals = nullptr; } };

  u64* cnt = (u64*)alloc((n + 1) * sizeof(u64));
  u64* stash = (u64*)alloc(n * sizeof(u64));
  u64* glist = (u64*)alloc(n * sizeof(u64));
  ull* d_general = (ull*)alloc(sizeof(ull));
  if(!cnt || !stash || !glist || !d_general) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "locate: out of device memory"); }
  LOC_TRY(cudaMemsetAsync(cnt + n, 0, sizeof(u64), st));
  LOC_TRY(cudaMemsetAsync(d_general, 0, sizeof(ull), st));
  locate_small_count_kernel<<<gridFor(n, sm), 256, 0, st>>>(v, d_sp, d_ep, n, cnt, stash, glist, d_general);
  ull n_general = 0;
  LOC_TRY(cudaMemcpyAsync(&n_general, d_general, sizeof(ull), cudaMemcpyDeviceToHost, st));
  LOC_TRY(cudaStreamSynchronize(st));
  if(std::getenv("GCSA_B200_LOCATE_DEBUG") != nullptr) { std::fprintf(stderr, "locate: %llu of %llu ranges through the general pipeline\n", n_general, (ull)n); }

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
    LOC_TRY(cudaMallocAsync(&p, std::max<u64>(distinct, 1) * sizeof(u64), st));
    *d_values_alloc = (u64*)p; d_values = (u64*)p; capacity = distinct;
  }
  if(d_values == nullptr || capacity < distinct) { rc = GCSA_B200_ERR_CAPACITY; g_last_error = "locate: output capacity too small"; }
  else if(distinct > 0) { locate_small_fill_kernel<<<gridFor(n, sm), 256, 0, st>>>(v, d_sp, d_ep, n, d_out_offsets, stash, goffs, gvals, d_values); }
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
  const int STREAMS = 2;
  const u64 CHUNK = std::max<u64>(1ull << 18, (n + 7) / 8);
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
    if(cudaMallocAsync((void**)&ch.d_sp, m * sizeof(u64), st) != cudaSuccess || cudaMallocAsync((void**)&ch.d_ep, m * sizeof(u64), st) != cudaSuccess ||
       cudaMallocAsync((void**)&ch.d_offs, (m + 1) * sizeof(u64), st) != cudaSuccess)
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
  out_offsets[0] = 0;
  {
    u64 t = 0;
    for(u64 i = 0; i < n; i++)
    {
      while(t < full_id.size() && full_id[t] < i) { t++; }
      auto it = special.find(i);
      u64 size = 0;
      if(it != special.end()) { size = it->second.result.size(); }
      else if(t < full_id.size() && full_id[t] == i) { size = full_offs[t + 1] - full_offs[t]; }
      out_offsets[i + 1] = out_offsets[i] + size;
    }
  }
  u64* vals = (u64*)std::malloc(std::max<u64>(out_offsets[n], 1) * sizeof(u64));
  {
    u64 t = 0;
    for(u64 i = 0; i < n; i++)
    {
      while(t < full_id.size() && full_id[t] < i) { t++; }
      auto it = special.find(i);
      if(it != special.end()) { std::copy(it->second.result.begin(), it->second.result.end(), vals + out_offsets[i]); }
      else if(t < full_id.size() && full_id[t] == i) { std::copy(full_vals + full_offs[t], full_vals + full_offs[t + 1], vals + out_offsets[i]); }
    }
  }
  std::free(full_vals);
  *values = (uint64_t*)vals;
  return 0;
}

//------------------------------------------------------------------------------
// LCP
//------------------------------------------------------------------------------

int gcsa_b200_lcp_create(const gcsa_flat_lcp* host, int device, gcsa_b200_lcp** out)
{
  if(host == nullptr || out == nullptr) { return fail(GCSA_B200_ERR_INVALID, "lcp_create: null argument"); }
  *out = nullptr;
  if(host->levels + 1 > 16 || host->levels == 0 || host->branching < 2) { return fail(GCSA_B200_ERR_INVALID, "lcp_create: bad tree shape"); }
  int n_dev = gcsa_b200_device_count();
  if(n_dev <= 0) { return fail(GCSA_B200_ERR_CUDA, "lcp_create: no CUDA device available (this engine has no CPU fallback)"); }
  if(device < 0 || device >= n_dev) { return fail(GCSA_B200_ERR_INVALID, "lcp_create: bad device ordinal"); }
  DeviceGuard guard(device);
  gcsa_b200_lcp* l = new gcsa_b200_lcp();
  l->device = device;
  cudaDeviceGetAttribute(&l->sm_count, cudaDevAttrMultiProcessorCount, device);
  LcpView& v = l->view;
  std::memset(&v, 0, sizeof(v));
  v.size = host->size; v.branching = host->branching; v.levels = host->levels;
  for(u64 i = 0; i <= host->levels; i++) { v.offsets[i] = host->offsets[i]; }
  for(u64 i = host->levels + 1; i < 16; i++) { v.offsets[i] = ~0ull; }
  v.values = host->offsets[host->levels];
  v.shift = -1;
  if((host->branching & (host->branching - 1)) == 0) { v.shift = 0; while((1ull << v.shift) < host->branching) { v.shift++; } }
  cudaError_t e = cudaMalloc(&l->data, ((std::max<u64>(v.values, 16) + 15) / 8) * 8);      // whole 8-byte words (the scans read words)
  if(e == cudaSuccess) { e = cudaMemset(l->data, 0xFF, ((std::max<u64>(v.values, 16) + 15) / 8) * 8); }
  if(e == cudaSuccess && v.values) { e = cudaMemcpy(l->data, host->data, v.values, cudaMemcpyHostToDevice); }
  if(e != cudaSuccess) { if(l->data) { cudaFree(l->data); } delete l; return fail(GCSA_B200_ERR_CUDA, std::string("lcp_create: ") + cudaGetErrorString(e)); }
  v.data = (const u8*)l->data;
  *out = l;
  return 0;
}

void gcsa_b200_lcp_destroy(gcsa_b200_lcp* lcp)
{
  if(lcp == nullptr) { return; }
  DeviceGuard guard(lcp->device);
  cudaFree(lcp->data);
  delete lcp;
}

int gcsa_b200_parent_batch(const gcsa_b200_lcp* lcp, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                           gcsa_b200_stnode* d_out, void* stream)
{
  if(lcp == nullptr) { return fail(GCSA_B200_ERR_INVALID, "parent_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(lcp->device);
  parent_kernel<<<gridFor(n, lcp->sm_count), 256, 0, (cudaStream_t)stream>>>(lcp->view, (const u64*)d_sp, (const u64*)d_ep, n, d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_parent_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n, gcsa_b200_stnode* out)
{
  HOST_PROLOGUE("parent_host", lcp);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  gcsa_b200_stnode* o = sc.alloc<gcsa_b200_stnode>(n);
  int rc = gcsa_b200_parent_batch(lcp, a, b, n, o, sc.stream);
  sc.out(out, o, n);
  HOST_EPILOGUE("parent_host", rc);
}

int gcsa_b200_depth_batch(const gcsa_b200_lcp* lcp, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                          uint64_t* d_out, void* stream)
{
  if(lcp == nullptr) { return fail(GCSA_B200_ERR_INVALID, "depth_batch: null handle"); }
  if(n == 0) { return 0; }
  DeviceGuard guard(lcp->device);
  depth_kernel<<<gridFor(n, lcp->sm_count), 256, 0, (cudaStream_t)stream>>>(lcp->view, (const u64*)d_sp, (const u64*)d_ep, n, (u64*)d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int gcsa_b200_depth_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n, uint64_t* out)
{
  HOST_PROLOGUE("depth_host", lcp);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n); u64* o = sc.alloc<u64>(n);
  int rc = gcsa_b200_depth_batch(lcp, a, b, n, o, sc.stream);
  sc.out((u64*)out, o, n);
  HOST_EPILOGUE("depth_host", rc);
}

int gcsa_b200_lcp_sv_host(const gcsa_b200_lcp* lcp, int which, const uint64_t* pos, uint64_t n,
                          uint64_t* out_pos, uint64_t* out_val)
{
  if(which < 0 || which > 3) { return fail(GCSA_B200_ERR_INVALID, "lcp_sv_host: which must be 0..3"); }
  HOST_PROLOGUE("lcp_sv_host", lcp);
  u64* a = sc.in((const u64*)pos, n); u64* op = sc.alloc<u64>(n); u64* ov = sc.alloc<u64>(n);
  int rc = 0;
  if(n) { lcp_sv_kernel<<<gridFor(n, lcp->sm_count), 256, 0, sc.stream>>>(lcp->view, which, a, n, op, ov); }
  sc.out((u64*)out_pos, op, n); sc.out((u64*)out_val, ov, n);
  HOST_EPILOGUE("lcp_sv_host", rc);
}

int gcsa_b200_lcp_rmq_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                           uint64_t* out_pos, uint64_t* out_val)
{
  HOST_PROLOGUE("lcp_rmq_host", lcp);
  u64* a = sc.in((const u64*)sp, n); u64* b = sc.in((const u64*)ep, n);
  u64* op = sc.alloc<u64>(n); u64* ov = sc.alloc<u64>(n);
  int rc = 0;
  if(n) { lcp_rmq_kernel<<<gridFor(n, lcp->sm_count), 256, 0, sc.stream>>>(lcp->view, a, b, n, op, ov); }
  sc.out((u64*)out_pos, op, n); sc.out((u64*)out_val, ov, n);
  HOST_EPILOGUE("lcp_rmq_host", rc);
}



//------------------------------------------------------------------------------
// countKMers
//------------------------------------------------------------------------------

/*
  countKMers(index, k, parameters), src/algorithms.cpp:387-421: the number of distinct k-mers over
  the bases (include_Ns: bases and N).  The reference walks the trie depth-first, one OpenMP task
  per 5-mer seed; here every level of the trie is one frontier expanded by one kernel launch.
  If ranges != NULL, *ranges receives the final frontier (malloc'ed sp[0..count) then ep[0..count)).
*/
int gcsa_b200_count_kmers(const gcsa_b200_index* index, uint64_t k, int include_Ns, uint64_t* result, uint64_t** ranges)
{
  if(index == nullptr || result == nullptr) { return fail(GCSA_B200_ERR_INVALID, "count_kmers: null argument"); }
  *result = 0;
  if(ranges) { *ranges = nullptr; }
  if(k == 0) { *result = 1; return 0; }
  if(index->header.path_nodes == 0) { return 0; }
  DeviceGuard guard(index->device);
  cudaStream_t st;
  CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const u32 chars = (include_Ns ? GCSA_B200_SIGMA - 2 : GCSA_B200_FAST_CHARS);
  u64 n = 1;
  u64 *sp = nullptr, *ep = nullptr;
  cudaError_t e = cudaSuccess;
  int rc = 0;
  #define KM_TRY(expr) do { e = (expr); if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("count_kmers: " #expr ": ") + cudaGetErrorString(e)); goto done; } } while(0)
  {
    u64 root[2] = { 0, index->header.path_nodes - 1 };
    KM_TRY(cudaMallocAsync(&sp, sizeof(u64), st)); KM_TRY(cudaMallocAsync(&ep, sizeof(u64), st));
    KM_TRY(cudaMemcpyAsync(sp, &root[0], sizeof(u64), cudaMemcpyHostToDevice, st));
    KM_TRY(cudaMemcpyAsync(ep, &root[1], sizeof(u64), cudaMemcpyHostToDevice, st));
    for(u64 level = 0; level < k && n > 0; level++)
    {
      u64 total = n * chars;
      u64 *csp = nullptr, *cep = nullptr, *flag = nullptr, *pos = nullptr;
      KM_TRY(cudaMallocAsync(&csp, total * sizeof(u64), st)); KM_TRY(cudaMallocAsync(&cep, total * sizeof(u64), st));
      KM_TRY(cudaMallocAsync(&flag, (total + 1) * sizeof(u64), st)); KM_TRY(cudaMallocAsync(&pos, (total + 1) * sizeof(u64), st));
      KM_TRY(cudaMemsetAsync(flag + total, 0, sizeof(u64), st));
      kmer_expand_kernel<<<gridFor(total, index->sm_count), 256, 0, st>>>(index->view, sp, ep, n, chars, csp, cep, flag);
      rc = scanExclusive(flag, pos, total + 1, st);
      if(rc) { goto done; }
      u64 next = 0;
      KM_TRY(cudaMemcpyAsync(&next, pos + total, sizeof(u64), cudaMemcpyDeviceToHost, st));
      KM_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(sp, st); cudaFreeAsync(ep, st); sp = ep = nullptr;
      KM_TRY(cudaMallocAsync(&sp, std::max<u64>(next, 1) * sizeof(u64), st)); KM_TRY(cudaMallocAsync(&ep, std::max<u64>(next, 1) * sizeof(u64), st));
      kmer_compact_kernel<<<gridFor(total, index->sm_count), 256, 0, st>>>(csp, cep, flag, pos, total, sp, ep);
      cudaFreeAsync(csp, st); cudaFreeAsync(cep, st); cudaFreeAsync(flag, st); cudaFreeAsync(pos, st);
      n = next;
    }
    *result = n;
    if(ranges && n > 0)
    {
      u64* out = (u64*)std::malloc(2 * n * sizeof(u64));
      KM_TRY(cudaMemcpyAsync(out, sp, n * sizeof(u64), cudaMemcpyDeviceToHost, st));
      KM_TRY(cudaMemcpyAsync(out + n, ep, n * sizeof(u64), cudaMemcpyDeviceToHost, st));
      *ranges = (uint64_t*)out;
    }
  }
done:
  if(sp) { cudaFreeAsync(sp, st); }
  if(ep) { cudaFreeAsync(ep, st); }
  e = cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  #undef KM_TRY
  if(rc == 0 && e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("count_kmers: ") + cudaGetErrorString(e)); }
  return rc;
}

/*
  compareKMers(left, right, k, parameters), src/algorithms.cpp:535-616: result = (kmers in both, only
  in left, only in right).  The reference walks both tries depth-first in lockstep, one OpenMP task per
  5-mer seed; here every level is one frontier of (left range, right range) states expanded by one launch.
  Both indexes must live on the same device.
*/
int gcsa_b200_compare_kmers(const gcsa_b200_index* left, const gcsa_b200_index* right, uint64_t k, int include_Ns,
                            uint64_t* result, gcsa_b200_kmer_state** left_kmers, gcsa_b200_kmer_state** right_kmers)
{
  if(left == nullptr || right == nullptr || result == nullptr) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: null argument"); }
  result[0] = result[1] = result[2] = 0;
  if(left_kmers) { *left_kmers = nullptr; }
  if(right_kmers) { *right_kmers = nullptr; }
  if(k == 0) { result[0] = 1; return 0; }                                         // algorithms.cpp:540
  if(k > 64) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: comparison is only supported for k <= 64"); }   // KMerComparisonState::MAX_K
  if(left->device != right->device) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: the indexes live on different devices"); }
  if(left->header.path_nodes == 0 && right->header.path_nodes == 0) { return 0; }
  const bool want = (left_kmers != nullptr || right_kmers != nullptr);
  HOST_PROLOGUE("compare_kmers", left);
  cudaStream_t st = sc.stream;
  const u32 chars = (include_Ns ? GCSA_B200_SIGMA - 2 : GCSA_B200_FAST_CHARS);
  int rc = 0;
  u64 n = 1;
  u64 root[4] = { 0, left->header.path_nodes - 1, 0, right->header.path_nodes - 1 };
  u64 zero_kmer[3] = { 0, 0, 0 };
  u64* state = nullptr; u64* kmer = nullptr;
  cudaError_t e = cudaSuccess;
  #define CK_TRY(expr) do { e = (expr); if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("compare_kmers: " #expr ": ") + cudaGetErrorString(e)); goto done; } } while(0)
  #define CK_ALLOC(ptr, count) do { CK_TRY(cudaMallocAsync((void**)&(ptr), std::max<u64>((count), 1) * sizeof(u64), st)); } while(0)
  {
    CK_ALLOC(state, 4); CK_TRY(cudaMemcpyAsync(state, root, sizeof(root), cudaMemcpyHostToDevice, st));
    if(want) { CK_ALLOC(kmer, 3); CK_TRY(cudaMemcpyAsync(kmer, zero_kmer, sizeof(zero_kmer), cudaMemcpyHostToDevice, st)); }
    for(u64 level = 0; level < k && n > 0; level++)
    {
      u64 total = n * chars, next = 0;
      u64 *child = nullptr, *child_kmer = nullptr, *flag = nullptr, *pos = nullptr, *new_state = nullptr, *new_kmer = nullptr;
      CK_ALLOC(child, 4 * total); CK_ALLOC(flag, total + 1); CK_ALLOC(pos, total + 1);
      if(want) { CK_ALLOC(child_kmer, 3 * total); }
      CK_TRY(cudaMemsetAsync(flag + total, 0, sizeof(u64), st));
      compare_expand_kernel<<<gridFor(total, left->sm_count), 256, 0, st>>>(left->view, right->view, state, n, kmer, chars, level, child, child_kmer, flag);
      rc = scanExclusive(flag, pos, total + 1, st);
      if(rc) { goto done; }
      CK_TRY(cudaMemcpyAsync(&next, pos + total, sizeof(u64), cudaMemcpyDeviceToHost, st));
      CK_TRY(cudaStreamSynchronize(st));
      CK_ALLOC(new_state, 4 * next);
      if(want) { CK_ALLOC(new_kmer, 3 * next); }
      compare_compact_kernel<<<gridFor(total, left->sm_count), 256, 0, st>>>(child, child_kmer, flag, pos, total, next, new_state, new_kmer);
      cudaFreeAsync(child, st); cudaFreeAsync(flag, st); cudaFreeAsync(pos, st); cudaFreeAsync(state, st);
      if(child_kmer) { cudaFreeAsync(child_kmer, st); }
      if(kmer) { cudaFreeAsync(kmer, st); }
      state = new_state; kmer = new_kmer; n = next;
    }
    if(n > 0)
    {
      ull* counts = nullptr; u64 *lflag = nullptr, *rflag = nullptr, *lpos = nullptr, *rpos = nullptr;
      CK_TRY(cudaMallocAsync((void**)&counts, 3 * sizeof(ull), st));
      CK_TRY(cudaMemsetAsync(counts, 0, 3 * sizeof(ull), st));
      if(want)
      {
        CK_ALLOC(lflag, n + 1); CK_ALLOC(rflag, n + 1); CK_ALLOC(lpos, n + 1); CK_ALLOC(rpos, n + 1);
        CK_TRY(cudaMemsetAsync(lflag + n, 0, sizeof(u64), st)); CK_TRY(cudaMemsetAsync(rflag + n, 0, sizeof(u64), st));
      }
      compare_classify_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(state, n, counts, lflag, rflag);
      ull host_counts[3] = { 0, 0, 0 };
      CK_TRY(cudaMemcpyAsync(host_counts, counts, sizeof(host_counts), cudaMemcpyDeviceToHost, st));
      CK_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(counts, st);
      for(int i = 0; i < 3; i++) { result[i] = host_counts[i]; }
      if(want)
      {
        rc = scanExclusive(lflag, lpos, n + 1, st); if(rc) { goto done; }
        rc = scanExclusive(rflag, rpos, n + 1, st); if(rc) { goto done; }
        for(int side = 0; side < 2; side++)
        {
          gcsa_b200_kmer_state** target = (side == 0 ? left_kmers : right_kmers);
          u64 count = result[1 + side];
          if(target == nullptr || count == 0) { continue; }
          u64* records = nullptr;
          CK_ALLOC(records, 8 * count);
          compare_emit_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(state, kmer, n, k, side == 0 ? lflag : rflag, side == 0 ? lpos : rpos, records);
          gcsa_b200_kmer_state* host = (gcsa_b200_kmer_state*)std::malloc(count * sizeof(gcsa_b200_kmer_state));
          if(host == nullptr) { rc = fail(GCSA_B200_ERR_NOMEM, "compare_kmers: out of host memory"); cudaFreeAsync(records, st); goto done; }
          *target = host;
          CK_TRY(cudaMemcpyAsync(host, records, count * sizeof(gcsa_b200_kmer_state), cudaMemcpyDeviceToHost, st));
          CK_TRY(cudaStreamSynchronize(st));
          cudaFreeAsync(records, st);
        }
        cudaFreeAsync(lflag, st); cudaFreeAsync(rflag, st); cudaFreeAsync(lpos, st); cudaFreeAsync(rpos, st);
      }
    }
  }
done:
  if(state) { cudaFreeAsync(state, st); }
  if(kmer) { cudaFreeAsync(kmer, st); }
  #undef CK_TRY
  #undef CK_ALLOC
  if(rc != 0)
  {
    if(left_kmers && *left_kmers) { std::free(*left_kmers); *left_kmers = nullptr; }
    if(right_kmers && *right_kmers) { std::free(*right_kmers); *right_kmers = nullptr; }
  }
  HOST_EPILOGUE("compare_kmers", rc);
}

//------------------------------------------------------------------------------
// MEM-style scan
//------------------------------------------------------------------------------

/*
  One pass over the patterns: every lane counts its matches and writes the first `stride` of them into a
  scratch slot of its pattern; after the scan of the counts a gather kernel moves them into the CSR, and the
  few patterns with more matches are redone writing at their final positions.  (The first version ran the
  whole scan twice, once to count and once to write.)  d_matches_alloc != NULL: the values are allocated
  here (stream-ordered) instead of being written to d_matches.
*/
static int memDevice(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* d_chars, const uint64_t* d_offsets,
                     uint64_t n, uint64_t* d_out_offsets, uint64_t* d_matches, uint64_t capacity, uint64_t* needed, cudaStream_t st,
                     u64** d_matches_alloc)
{
  if(index == nullptr || lcp == nullptr || d_out_offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "mem_batch: null argument"); }
  if(index->device != lcp->device) { return fail(GCSA_B200_ERR_INVALID, "mem_batch: index and LCP array live on different devices"); }
  if(index->header.path_nodes != lcp->view.size) { return fail(GCSA_B200_ERR_INVALID, "mem_batch: index and LCP array have different sizes"); }
  DeviceGuard guard(index->device);
  if(needed) { *needed = 0; }
  if(d_matches_alloc) { *d_matches_alloc = nullptr; }
  CUDA_TRY(cudaMemsetAsync(d_out_offsets, 0, (n + 1) * sizeof(u64), st));
  if(n == 0 || index->header.path_nodes == 0) { return 0; }

  // scratch: up to 16 matches per pattern, fewer for huge batches, none (two full passes) if even 4 do not fit
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  u64 stride = std::min<u64>(16, (free_b / 8) / (n * 32));
  if(const char* e = std::getenv("GCSA_B200_MEM_STRIDE")) { stride = std::min<u64>(stride, (u64)std::atoi(e)); }   // tests: 0 = two passes
  if(stride < 4 && std::getenv("GCSA_B200_MEM_STRIDE") == nullptr) { stride = 0; }

  std::vector<void*> tmp;
  auto alloc = [&](u64 bytes) -> void* { void* p = nullptr; if(cudaMallocAsync(&p, std::max<u64>(bytes, 16), st) != cudaSuccess) { cudaGetLastError(); return nullptr; } tmp.push_back(p); return p; };
  auto cleanup = [&]() { for(void* p : tmp) { cudaFreeAsync(p, st); } tmp.clear(); };
  #define MEM_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { cleanup(); \
    return fail(GCSA_B200_ERR_CUDA, std::string("mem_batch: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)

  u64* counts = (u64*)alloc((n + 1) * sizeof(u64));
  ull* n_overflow = (ull*)alloc(sizeof(ull));
  u64* scratch = (stride > 0 ? (u64*)alloc(n * stride * 32) : nullptr);
  if(counts == nullptr || n_overflow == nullptr) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "mem_batch: out of device memory"); }
  if(scratch == nullptr) { stride = 0; }
  MEM_TRY(cudaMemsetAsync(counts, 0, (n + 1) * sizeof(u64), st));
  MEM_TRY(cudaMemsetAsync(n_overflow, 0, sizeof(ull), st));
  int grid = gridFor(n, index->sm_count, 4);
  u32 parent_batch = 8;
  if(const char* e = std::getenv("GCSA_B200_MEM_PARENT_BATCH")) { parent_batch = (u32)std::max(1, std::atoi(e)); }
  // GCSA_B200_MEM_JUMP=1: singleton ranges follow the jump tables (mem_kernel<.., JUMP>); off by default until measured
  bool jump = false;
  if(const char* e = std::getenv("GCSA_B200_MEM_JUMP")) { jump = (std::atoi(e) != 0 && index->view.jump != nullptr && index->view.default_alphabet != 0); }
  if(stride > 0)
  {
    if(jump) { mem_kernel<2, true><<<grid, 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, scratch, nullptr, stride, parent_batch); }
    else { mem_kernel<2><<<grid, 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, scratch, nullptr, stride, parent_batch); }
  }
  else
  {
    if(jump) { mem_kernel<0, true><<<grid, 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, nullptr, nullptr, 0, parent_batch); }
    else { mem_kernel<0><<<grid, 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, counts, nullptr, nullptr, nullptr, 0, parent_batch); }
  }
  int rc = scanExclusive(counts, (u64*)d_out_offsets, n + 1, st);
  if(rc) { cleanup(); return rc; }
  if(stride > 0) { mem_count_overflow_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>(counts, n, stride, n_overflow); }
  u64 total = 0; ull overflowing = 0;
  MEM_TRY(cudaMemcpyAsync(&total, (u64*)d_out_offsets + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
  MEM_TRY(cudaMemcpyAsync(&overflowing, n_overflow, sizeof(ull), cudaMemcpyDeviceToHost, st));
  MEM_TRY(cudaStreamSynchronize(st));
  if(needed) { *needed = total; }
  if(d_matches_alloc != nullptr)
  {
    void* p = nullptr;
    MEM_TRY(cudaMallocAsync(&p, std::max<u64>(total, 1) * 32, st));
    *d_matches_alloc = (u64*)p; d_matches = (u64*)p; capacity = total;
  }
  if(d_matches == nullptr || capacity < total) { cleanup(); return fail(GCSA_B200_ERR_CAPACITY, "mem_batch: output capacity too small"); }
  if(stride == 0)
  {
    if(jump) { mem_kernel<1, true><<<grid, 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, nullptr, (const u64*)d_out_offsets, (u64*)d_matches, nullptr, 0, parent_batch); }
    else { mem_kernel<1><<<grid, 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, n, nullptr, (const u64*)d_out_offsets, (u64*)d_matches, nullptr, 0, parent_batch); }
  }
  else
  {
    u64* overflow = (u64*)alloc(std::max<u64>(overflowing, 1) * sizeof(u64));
    if(overflow == nullptr) { cleanup(); return fail(GCSA_B200_ERR_NOMEM, "mem_batch: out of device memory"); }
    MEM_TRY(cudaMemsetAsync(n_overflow, 0, sizeof(ull), st));
    mem_gather_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>((const ulonglong4*)scratch, counts, (const u64*)d_out_offsets, n, stride,
                                                                   (ulonglong4*)d_matches, overflow, n_overflow);
    if(overflowing > 0)
    {
      if(jump)
      {
        mem_kernel<1, true><<<gridFor(overflowing, index->sm_count, 4), 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, overflowing,
                                                                                    nullptr, (const u64*)d_out_offsets, (u64*)d_matches, overflow, 0, parent_batch);
      }
      else
      {
        mem_kernel<1><<<gridFor(overflowing, index->sm_count, 4), 256, 0, st>>>(index->view, lcp->view, d_chars, (const u64*)d_offsets, 0, overflowing,
                                                                              nullptr, (const u64*)d_out_offsets, (u64*)d_matches, overflow, 0, parent_batch);
      }
    }
  }
  MEM_TRY(cudaGetLastError());
  cleanup();
  #undef MEM_TRY
  return 0;
}

int gcsa_b200_mem_batch(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* d_chars, const uint64_t* d_offsets,
                        uint64_t n, uint64_t* d_out_offsets, uint64_t* d_matches, uint64_t capacity, uint64_t* needed, void* stream)
{
  return memDevice(index, lcp, d_chars, d_offsets, n, d_out_offsets, d_matches, capacity, needed, (cudaStream_t)stream, nullptr);
}

int gcsa_b200_mem_host(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* chars, const uint64_t* offsets,
                       uint64_t n, uint64_t* out_offsets, uint64_t** matches)
{
  if(out_offsets == nullptr || matches == nullptr || (n > 0 && (chars == nullptr || offsets == nullptr))) { return fail(GCSA_B200_ERR_INVALID, "mem_host: null argument"); }
  *matches = nullptr;
  HOST_PROLOGUE("mem_host", index);
  u64 total_chars = (n ? offsets[n] : 0);
  u8* d_chars = sc.in(chars, total_chars + 1 > 1 ? total_chars : 1);
  u64* d_off = sc.in((const u64*)offsets, n + 1);
  u64* d_out = sc.alloc<u64>(n + 1);
  u64 needed = 0;
  u64* d_vals = nullptr;
  int rc = memDevice(index, lcp, d_chars, d_off, n, d_out, nullptr, 0, &needed, sc.stream, &d_vals);
  if(rc == 0)
  {
    u64* vals = (u64*)std::malloc(std::max<u64>(4 * needed, 1) * sizeof(u64));
    if(d_vals != nullptr) { sc.out(vals, d_vals, 4 * needed); sc.ptrs.push_back(d_vals); }
    sc.out((u64*)out_offsets, d_out, n + 1);
    *matches = (uint64_t*)vals;
  }
  HOST_EPILOGUE("mem_host", rc);
}
