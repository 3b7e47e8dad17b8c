/* device/find.cuh -- find(): the batched backward-search kernels.
   Included by find.cu only (its kernels must be defined in one translation unit); sm_100a only. */
#ifndef GCSA2_B200_DEVICE_FIND_CUH
#define GCSA2_B200_DEVICE_FIND_CUH

//------------------------------------------------------------------------------
// Kernels: find
//------------------------------------------------------------------------------

struct FindStatsDev { u64 found, total_length, lf_steps, sector_probes, table_hits; };

/*
  GCSA::find(begin, end), include/gcsa/gcsa.h:96-110.  One query per lane.  Queries are pulled from a
  contiguous per-warp slice; lanes whose search ended are refilled together once half the warp is idle
  (one ballot + popc, no atomics).  A refilled lane packs the last 32 characters of its pattern into one
  register (2 bits each, the last character lowest): the k-mer table index is a bit field of it and a jump
  along a unary path is one XOR against the table entry.  Anything that does not fit the fast forms (other
  characters, another alphabet, short remainders) goes through the per-character path, which is the
  reference's loop verbatim.
*/
// Work-list entries of the two-kernel form (find_fast_kernel below leaves what it cannot finish with one or two
// probes to this kernel): query number | remaining characters << 48 | flags.
#define WORK_QUERY_MASK ((1ull << 48) - 1)
#define WORK_NO_JUMP (1ull << 62)        // a jump failed on a character: the query dies within a few single steps
#define WORK_FRESH   (1ull << 63)        // nothing is known yet: search from the last character
#define WORK_QUAD    (1ull << 61)        // (inside find_fast_kernel only) goes to the list of find_quad_kernel

template<bool STATS, int MIN_BLOCKS, bool PACKED = false, bool WORK = false>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
find_kernel(const DevView v, const u8* __restrict__ chars, const u64* __restrict__ offsets, u64 char_base,
            u64 fixed_length, u64 n, u64* __restrict__ sp_out, u64* __restrict__ ep_out, FindStatsDev* stats, int refill_at,
            const u64* __restrict__ work = nullptr, const unsigned long long* __restrict__ work_count = nullptr)
{
  // WORK: the queries are those of the work list (fixed-length patterns); an entry that is not FRESH resumes from the
  // range stored in sp_out / ep_out with `remaining` characters to go.
  if(WORK) { n = *work_count; }
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
          u64 entry = 0;
          if(WORK) { entry = work[cand]; q = entry & WORK_QUERY_MASK; }
          u64 b, e;
          if(offsets != nullptr) { b = offsets[q] - char_base; e = offsets[q + 1] - char_base; }
          else { b = q * fixed_length; e = b + fixed_length; }
          begin = b; live = true; jump_mode = 1;
          tail = 0; tail_n = 0; tail_end = e;
          if(e == b || v.path_nodes == 0) { sp = 0; ep = v.path_nodes - 1; pos = b; }
          else if(WORK && !(entry & WORK_FRESH))
          {
            pack_tail(e);
            sp = sp_out[q]; ep = ep_out[q];
            pos = b + ((entry >> 48) & 0xFF);
            if(entry & WORK_NO_JUMP) { jump_mode = 0; }
          }
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
        const u64* jump_from = nullptr; bool jump_wide = false;
        if((v.jump != nullptr || v.jump_wide != nullptr) && jump_mode != 0 && sp == ep)
        {
          u64 left = pos - begin;
          if(left >= (u64)v.jump_k) { jump_from = v.jump; jump_wide = (v.jump_wide != nullptr); }
          else if(left >= 4) { jump_from = v.jump_short; }
        }
        if(jump_from != nullptr || jump_wide)
        {
          JumpPath path = (jump_wide ? jump_decode_wide(__ldg(v.jump_wide + sp)) : jump_decode(__ldg(jump_from + sp), v.jump_tbits));
          u32 len = path.len;
          if(STATS) { sectors++; }
          if(len >= 2)
          {
            u64 stored = path.chars;
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
              sp = ep = path.target;
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
  The first kernel of the two-kernel form of find() for batches of k-mers (one fixed length L >= table_k, default
  alphabet, a k-mer table): ONE QUERY PER THREAD, no loop.  It works on the last 32 characters of a pattern; a longer
  pattern is continued by the general kernel from where the table and the first jump left it.  A thread reads its pattern (consecutive threads,
  consecutive patterns: the loads coalesce), packs it to 2 bits per character, looks the last table_k characters up in
  the k-mer table and, if the result is one path node and characters remain, takes one jump (fused with the table
  entry, or one more load).  That finishes most k-mers of a large reference (a 32-mer over a 16-mer table: table
  entry + 16-step jump); the queries it cannot finish -- a range of several nodes, a jump that is too short or fails,
  another character in the pattern -- are appended to a work list with their state (the range goes into sp_out /
  ep_out) and the general kernel resumes them (find_kernel<.., WORK = true>).  Compared with running everything through
  the general kernel: no refill logic, no divergence between lanes in different phases, a third of the instructions.
*/
// Queries per thread and round.  What bounds this kernel is how many random table probes the chip has in flight
// (ncu: 94 % occupancy, 31 registers, and still 62 cycles of long-scoreboard stall per issue with ONE probe per thread:
// 17.5 G probes/s where dependent-chain microbenchmarks reach 44 G/s), so a thread issues the probes of several
// queries back to back before it looks at any of them.

template<bool STATS, bool PACKED, int U>
__global__ void __launch_bounds__(256, STATS ? 1 : (U >= 4 ? 4 : (U == 2 ? 6 : 8)))
find_fast_kernel(const DevView v, const u8* __restrict__ chars, u32 L, u64 n, u64* __restrict__ sp_out, u64* __restrict__ ep_out,
                 u64* __restrict__ work, unsigned long long* __restrict__ work_count,
                 ulonglong2* __restrict__ quad_work, unsigned long long* __restrict__ quad_count, FindStatsDev* stats)
{
  const u32 lane = threadIdx.x & 31;
  const u32 k = (u32)v.table_k;
  const u64 kmask = (k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1));
  u64 st_found = 0, st_len = 0, st_steps = 0, st_sectors = 0, st_hits = 0;
  // a warp takes 32 * U consecutive queries per round: lane l the queries base + 32 j + l (coalesced for every j)
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u64 per_round = 32ull * U;
  const u64 rounds = (n + n_warps * per_round - 1) / (n_warps * per_round);
  for(u64 r = 0; r < rounds; r++)
  {
    const u64 base = (r * n_warps + warp) * per_round + lane;
    u64 tail[U], res[U], je[U], entry[U];
    bool valid[U];
    // ---- the patterns, 2 bits per character, the LAST character in the lowest bits ----
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      const u64 q = base + 32ull * j;
      tail[j] = 0; valid[j] = (q < n); entry[j] = 0; res[j] = 0; je[j] = 0;
      if(q < n)
      {
        if(PACKED)
        {
          // the last min(L, 32) characters of the pattern's words (they straddle two words when L > 32 is not a multiple of 32)
          const u32 m = (L < 32 ? L : 32), r0 = L - m, sh = (r0 & 31) * 2;
          const unsigned long long* words = (const unsigned long long*)chars + q * (u64)((L + 31) >> 5) + (r0 >> 5);
          u64 x = __ldcs(words) >> sh;
          if(sh != 0 && (r0 & 31) + m > 32) { x |= __ldcs(words + 1) << (64 - sh); }
          u64 t = __brevll(x);
          t = ((t >> 1) & 0x5555555555555555ull) | ((t & 0x5555555555555555ull) << 1);
          tail[j] = (m < 32 ? t >> (2 * (32 - m)) : t);
        }
        else
        {
          // the 32 bytes that end with the pattern (the first of them belong to the previous pattern if L < 32)
          const u64 end = (u64)chars + (q + 1) * (u64)L;
          if(end - 32 < (u64)chars) { entry[j] = q | WORK_FRESH; }          // would read before the buffer: left to the general kernel
          else
          {
            const u64 a = end - 32; const u32 sh = (u32)(a & 7) * 8;
            const unsigned long long* words = (const unsigned long long*)(a - (a & 7));
            u64 w0 = __ldcs(words), w1 = __ldcs(words + 1), w2 = __ldcs(words + 2), w3 = __ldcs(words + 3);
            if(sh != 0)
            {
              u64 w4 = __ldcs(words + 4);
              w0 = (w0 >> sh) | (w1 << (64 - sh)); w1 = (w1 >> sh) | (w2 << (64 - sh));
              w2 = (w2 >> sh) | (w3 << (64 - sh)); w3 = (w3 >> sh) | (w4 << (64 - sh));
            }
            u32 g0, g1, g2, g3;
            u64 p3 = pack8_reversed(w3, &g3), p2 = pack8_reversed(w2, &g2), p1 = pack8_reversed(w1, &g1), p0 = pack8_reversed(w0, &g0);
            tail[j] = p3 | (p2 << 16) | (p1 << 32) | (p0 << 48);
            // how many characters, counted from the last one, are bases
            u32 good = (g3 < 8 ? g3 : 8 + (g2 < 8 ? g2 : 8 + (g1 < 8 ? g1 : 8 + g0)));
            if(good < (L < 32 ? L : 32)) { entry[j] = q | WORK_FRESH; }       // (characters further to the left are looked at by whoever gets there)
          }
        }
      }
    }
    // ---- the table probes of all U queries, issued before any is used ----
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      if(valid[j] && entry[j] == 0)
      {
        const u64 idx = tail[j] & kmask;
        if(v.table2 != nullptr) { ulonglong2 both = __ldg(v.table2 + idx); res[j] = both.x; je[j] = both.y; }
        else { res[j] = __ldg(v.table + idx); }
        if(STATS) { st_hits++; }
      }
    }
    // ---- separate long jump table: the second probe of the queries whose k-mer is one path node ----
    u64 jx[U], jy[U];
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      jx[j] = 0; jy[j] = 0;
      const u32 rem = L - k;
      if(valid[j] && entry[j] == 0 && v.table2 == nullptr && (res[j] >> 40) == 1 && rem >= v.jump_k && rem > 0)
      {
        const u64 node = res[j] & M40;
        if(v.jump_wide != nullptr) { ulonglong2 e = __ldg(v.jump_wide + node); jx[j] = e.x; jy[j] = e.y; if(STATS) { st_sectors++; } }
        else if(v.jump != nullptr) { jx[j] = __ldg(v.jump + node); if(STATS) { st_sectors++; } }
      }
    }
    // ---- resolve ----
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      const u64 q = base + 32ull * j;
      if(valid[j] && entry[j] == 0)
      {
        const u64 len = res[j] >> 40;
        if(len == TABLE_ESCAPE) { entry[j] = q | WORK_FRESH; }
        else
        {
          u64 sp = res[j] & M40, ep = sp + len - 1;
          u32 rem = L - k;
          if(len == 1 && rem > 0)
          {
            // one path node: its unary backward path, from the fused entry or from the long jump table
            JumpPath path; path.len = 0; path.chars = 0; path.target = 0;
            if(v.table2 != nullptr) { path = jump_decode(je[j], v.jump_tbits); }
            else if(v.jump_wide != nullptr) { path = jump_decode_wide(make_ulonglong2(jx[j], jy[j])); }
            else if(jx[j] != 0) { path = jump_decode(jx[j], v.jump_tbits); }
            if(path.len >= 2 && path.len <= rem)
            {
              if((((tail[j] >> (2 * k)) ^ path.chars) & ((1ull << (2 * path.len)) - 1)) == 0)
              {
                sp = ep = path.target; rem -= path.len;
                if(STATS) { st_steps += path.len; }
              }
              else { entry[j] = q | ((u64)rem << 48) | WORK_NO_JUMP; }    // it dies within these steps: the exact pair comes from single steps
            }
          }
          if(entry[j] == 0 && rem > 0 && len > 0)
          {
            entry[j] = q | ((u64)rem << 48);
            // a range of a few path nodes with a whole long path to go: their paths are probed side by side (find_quad_kernel)
            if(quad_work != nullptr && len >= 2 && len <= 4 && rem >= v.jump_k) { entry[j] |= WORK_QUAD; }
          }
          __stcs((unsigned long long*)sp_out + q, (unsigned long long)sp); __stcs((unsigned long long*)ep_out + q, (unsigned long long)ep);
          if(STATS && entry[j] == 0 && !range_empty(sp, ep)) { st_found++; st_len += ep + 1 - sp; }
        }
      }
    }
    // ---- append the unfinished queries of the warp to the work lists (one atomic per warp, list and round) ----
    u32 todo[U], quads[U]; u32 total = 0, total_quads = 0;
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      todo[j] = __ballot_sync(0xFFFFFFFFu, entry[j] != 0 && !(entry[j] & WORK_QUAD));
      quads[j] = __ballot_sync(0xFFFFFFFFu, (entry[j] & WORK_QUAD) != 0);
      total += __popc(todo[j]); total_quads += __popc(quads[j]);
    }
    if(total != 0)
    {
      unsigned long long at = 0;
      if(lane == 0) { at = atomicAdd(work_count, (unsigned long long)total); }
      at = __shfl_sync(0xFFFFFFFFu, at, 0);
      #pragma unroll
      for(int j = 0; j < U; j++)
      {
        if((todo[j] >> lane) & 1) { work[at + __popc(todo[j] & ((1u << lane) - 1))] = entry[j]; }
        at += __popc(todo[j]);
      }
    }
    if(total_quads != 0)
    {
      unsigned long long at = 0;
      if(lane == 0) { at = atomicAdd(quad_count, (unsigned long long)total_quads); }
      at = __shfl_sync(0xFFFFFFFFu, at, 0);
      #pragma unroll
      for(int j = 0; j < U; j++)
      {
        if((quads[j] >> lane) & 1) { quad_work[at + __popc(quads[j] & ((1u << lane) - 1))] = make_ulonglong2(entry[j] & ~WORK_QUAD, tail[j]); }
        at += __popc(quads[j]);
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
  The second kernel of the k-mer form: the queries whose k-mer table result is a range of two to four path nodes (a
  k-mer that occurs a few times: half of the 16-mers of a 3 Gbp reference) and that still have a whole long jump to go.
  LF of a range is the union of LF of its nodes, so instead of single steps until the range is one node, the long paths
  of all its nodes are probed SIDE BY SIDE: four lanes per query, lane t takes node sp + t, and the quad combines what
  the lanes found -- the nodes whose path spells the pattern's next characters map to the new range.  Exact when every
  node of the range has a path of one common length and the targets of the matching ones are contiguous (then the new
  range is [min, max] of them); anything else -- paths of different lengths, no match at all (the early-exit pair must
  come from the step that fails) -- goes on to the general kernel's work list unchanged.
*/
template<bool STATS>
__global__ void __launch_bounds__(256)
find_quad_kernel(const DevView v, u32 L, const ulonglong2* __restrict__ quad_work, const unsigned long long* __restrict__ quad_count,
                 u64* __restrict__ sp_out, u64* __restrict__ ep_out, u64* __restrict__ work, unsigned long long* __restrict__ work_count,
                 FindStatsDev* stats)
{
  const u64 n = *quad_count;
  const u32 lane = threadIdx.x & 31, t = lane & 3, quad_shift = lane & ~3u;
  u64 st_found = 0, st_len = 0, st_steps = 0, st_sectors = 0;
  const u64 stride = ((u64)gridDim.x * blockDim.x) >> 2;
  const u64 rounds = (n + stride - 1) / stride;
  for(u64 r = 0; r < rounds; r++)
  {
    // (every lane of the warp goes through the same collectives; a quad without an item carries dummies)
    const u64 i = r * stride + (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 2);
    const bool active = (i < n);
    u64 push = 0;                                                      // lane 0 of a quad: entry for the general kernel's list
    u64 q = 0, tail = 0, sp = 0, ep = 0; u32 rem = 0, nodes = 0;
    if(active)
    {
      const ulonglong2 item = quad_work[i];
      q = item.x & WORK_QUERY_MASK; tail = item.y; rem = (u32)((item.x >> 48) & 0xFF);
      sp = sp_out[q]; ep = ep_out[q];
      nodes = (u32)(ep - sp + 1);
    }
    // lane t: the long path of node sp + t
    JumpPath path; path.len = 0; path.chars = 0; path.target = 0;
    if(active && t < nodes)
    {
      if(v.jump_wide != nullptr) { path = jump_decode_wide(__ldg(v.jump_wide + sp + t)); }
      else { path = jump_decode(__ldg(v.jump + sp + t), v.jump_tbits); }
      if(STATS) { st_sectors++; }
    }
    const u32 consumed = L - rem;
    const bool exists = (active && t < nodes);
    const bool usable = !exists || (path.len >= 2 && path.len <= rem);
    const bool match = exists && usable && ((((tail >> (2 * consumed)) ^ path.chars) & ((1ull << (2 * path.len)) - 1)) == 0);
    // combine over the quad: one common length, the matching targets
    const u32 len0 = __shfl_sync(0xFFFFFFFFu, path.len, 0, 4);                                   // node sp always exists
    const u32 all_usable = (__ballot_sync(0xFFFFFFFFu, usable && (!exists || path.len == len0)) >> quad_shift) & 0xFu;
    const u32 matches = (__ballot_sync(0xFFFFFFFFu, match) >> quad_shift) & 0xFu;
    u64 lo = (match ? path.target : ~0ull), hi = (match ? path.target : 0ull);
    #pragma unroll
    for(int d = 1; d < 4; d <<= 1)
    {
      u64 olo = __shfl_xor_sync(0xFFFFFFFFu, lo, d, 4), ohi = __shfl_xor_sync(0xFFFFFFFFu, hi, d, 4);
      lo = (olo < lo ? olo : lo); hi = (ohi > hi ? ohi : hi);
    }
    if(active && t == 0)
    {
      const u32 count = (u32)__popc(matches);
      if(all_usable != 0xFu) { push = q | ((u64)rem << 48); }                                   // paths of different lengths: single steps
      else if(count == 0) { push = q | ((u64)rem << 48) | WORK_NO_JUMP; }                       // dies within these steps: the exact pair from single steps
      else if(hi - lo + 1 != (u64)count) { push = q | ((u64)rem << 48); }                       // (cannot happen: LF of a range is a range)
      else
      {
        rem -= len0;
        __stcs((unsigned long long*)sp_out + q, (unsigned long long)lo); __stcs((unsigned long long*)ep_out + q, (unsigned long long)hi);
        if(STATS) { st_steps += len0; }
        if(rem > 0) { push = q | ((u64)rem << 48); }
        else if(STATS) { st_found++; st_len += hi + 1 - lo; }
      }
    }
    const u32 todo = __ballot_sync(0xFFFFFFFFu, push != 0);
    if(todo != 0)
    {
      unsigned long long at = 0;
      if(lane == 0) { at = atomicAdd(work_count, (unsigned long long)__popc(todo)); }
      at = __shfl_sync(0xFFFFFFFFu, at, 0);
      if(push != 0) { work[at + __popc(todo & ((1u << lane) - 1))] = push; }
    }
  }
  if(STATS)
  {
    atomicAdd((ull*)&stats->found, (ull)st_found); atomicAdd((ull*)&stats->total_length, (ull)st_len);
    atomicAdd((ull*)&stats->lf_steps, (ull)st_steps); atomicAdd((ull*)&stats->sector_probes, (ull)st_sectors);
  }
}


/*
  The third kernel of the k-mer form, for patterns longer than the k-mer table plus one long jump (64-mers on a
  variation graph: find_fast_kernel leaves every one of them with 30+ characters to go): the entries of the work list
  whose range is one path node are followed to the end of their pattern here -- one jump-table entry (up to jump_k, or 4,
  backward steps along the node's unary path) or one fused sector (a single step, where the path branches or fewer than
  four characters are left) per round, U entries per thread with their probes issued together, straight-line and
  predicated like find_fast_kernel: patterns of one length need about the same number of rounds, so the lanes of a warp
  stay together, which the refill loop of the general kernel cannot offer (13 of 32 lanes active per instruction on this
  workload).  What cannot be finished here -- a range of several nodes, a character outside ACGT, and every query that
  is about to fail (the early-exit pair must come from the single step that fails, include/gcsa/gcsa.h:160) -- goes on to
  the general kernel through a second work list, with the state reached so far.
*/
// the up to 64 characters that end at position `end` (exclusive, > 0) of pattern q, 2 bits each: lo = the last 32 (the
// LAST one in the lowest bits), hi = the 32 before them; *good = how many, counted from the last, are usable (bases,
// inside the pattern).  One call serves 48 or more characters of the chain.
template<bool PACKED>
__device__ __forceinline__ void chain_window(const u8* __restrict__ chars, u64 origin, u32 end, u64* lo, u64* hi, u32* good)
{
  // origin: index of the pattern's first byte (PACKED: of its first 8-byte word)
  const u32 m = (end < 64 ? end : 64);
  *lo = 0; *hi = 0;
  if(PACKED)
  {
    // characters [end - m, end) of the pattern's words (32 per word, the first one in the lowest bits), reversed
    const unsigned long long* words = (const unsigned long long*)chars + origin;
    #pragma unroll
    for(int half = 0; half < 2; half++)
    {
      const u32 e2 = (half == 0 ? end : (end > 32 ? end - 32 : 0));          // this half covers [e2 - mm, e2)
      const u32 mm = (e2 < 32 ? e2 : 32);
      if(mm == 0) { continue; }
      const u32 r0 = e2 - mm, sh = (r0 & 31) * 2;
      u64 x = __ldcs(words + (r0 >> 5)) >> sh;
      if(sh != 0 && (r0 & 31) + mm > 32) { x |= __ldcs(words + (r0 >> 5) + 1) << (64 - sh); }
      u64 t = __brevll(x);
      t = ((t >> 1) & 0x5555555555555555ull) | ((t & 0x5555555555555555ull) << 1);
      t = (mm < 32 ? t >> (2 * (32 - mm)) : t);
      if(half == 0) { *lo = t; } else { *hi = t; }
    }
    *good = m;
    return;
  }
  const u64 last = (u64)chars + origin + end;                    // one past the last byte of the window
  if(last - 64 < (u64)chars) { *good = 0; return; }                // would read before the buffer: left to the general kernel
  const u64 a = last - 64; const u32 sh = (u32)(a & 7) * 8;
  const unsigned long long* words = (const unsigned long long*)(a - (a & 7));
  u64 w[9];
  #pragma unroll
  for(int i = 0; i < 8; i++) { w[i] = __ldcs(words + i); }
  w[8] = 0;
  if(sh != 0)
  {
    w[8] = __ldcs(words + 8);
    #pragma unroll
    for(int i = 0; i < 8; i++) { w[i] = (w[i] >> sh) | (w[i + 1] << (64 - sh)); }
  }
  // w[7] holds the last eight characters; usable characters are counted from there
  u32 g = 0; bool all = true;
  u64 packed[2] = { 0, 0 };
  #pragma unroll
  for(int i = 7; i >= 0; i--)
  {
    u32 gi;
    u64 p = pack8_reversed(w[i], &gi);
    packed[(7 - i) >> 2] |= p << (16 * ((7 - i) & 3));
    if(all) { g += gi; all = (gi == 8); }
  }
  *lo = packed[0]; *hi = packed[1];
  *good = (g < m ? g : m);
}

// FRESH: no first kernel and no input work list -- the patterns are those of the offsets form (any lengths): each entry
// starts with the k-mer table probe of its last characters and goes on as above.  Patterns shorter than the table or
// longer than 255 characters, table escapes and characters outside ACGT go to the general kernel as FRESH entries.
template<bool STATS, bool PACKED, int U, bool FRESH = false>
__global__ void __launch_bounds__(256, U == 4 ? 2 : (U == 2 ? 3 : 4))
find_chain_kernel(const DevView v, const u8* __restrict__ chars, u32 L, u64* __restrict__ sp_out, u64* __restrict__ ep_out,
                  const u64* __restrict__ work, const unsigned long long* __restrict__ work_count,
                  u64* __restrict__ work2, unsigned long long* __restrict__ work2_count, FindStatsDev* stats,
                  u32 straggle, const u64* __restrict__ offsets = nullptr, u64 char_base = 0, u64 n_fresh = 0)
{
  const u32 lane = threadIdx.x & 31;
  const u64 n = (FRESH ? n_fresh : *work_count);
  u64 st_found = 0, st_len = 0, st_steps = 0, st_sectors = 0, st_hits = 0;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u64 per_round = 32ull * U;
  const u64 rounds = (n + n_warps * per_round - 1) / (n_warps * per_round);
  const bool have_long = (v.jump != nullptr || v.jump_wide != nullptr);
  for(u64 r = 0; r < rounds; r++)
  {
    const u64 base = (r * n_warps + warp) * per_round + lane;
    // state of the U entries of this thread: node, characters left, the packed window [wend - 64, wend) and how many of its
    // characters (from the end) are usable; hand != 0: the entry for the second work list
    u64 q[U], node[U], win[U], win_hi[U], hand[U], origin[U];
    u32 rem[U], wend[U], wgood[U];
    bool active[U], step_next[U], moved[U];
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      const u64 at = base + 32ull * j;
      active[j] = false; step_next[j] = false; moved[j] = false; hand[j] = 0; q[j] = 0; node[j] = 0; win[j] = 0; win_hi[j] = 0; rem[j] = 0; wend[j] = 0; wgood[j] = 0;
      origin[j] = 0;
      if(FRESH)
      {
        if(at < n)
        {
          q[j] = at;
          // (patterns of one length L without an offsets array: pattern q starts at q * L)
          const u64 b = (offsets != nullptr ? offsets[at] - char_base : at * (u64)L);
          const u64 len = (offsets != nullptr ? offsets[at + 1] - offsets[at] : (u64)L);
          origin[j] = b;
          const u32 k = (u32)v.table_k;
          if(len == 0 || v.path_nodes == 0)                                // gcsa.h:99: the empty pattern matches everything
          {
            __stcs((unsigned long long*)sp_out + at, 0ull); __stcs((unsigned long long*)ep_out + at, (unsigned long long)(v.path_nodes - 1));
            if(STATS && v.path_nodes != 0) { st_found++; st_len += v.path_nodes; }
          }
          else if(len > 255 || len < k) { hand[j] = at | WORK_FRESH; }
          else
          {
            chain_window<PACKED>(chars, b, (u32)len, &win[j], &win_hi[j], &wgood[j]); wend[j] = (u32)len;
            if(wgood[j] < k) { hand[j] = at | WORK_FRESH; }
            else
            {
              const u64 idx = win[j] & (k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1));
              u64 res, fused_jump = 0;
              if(v.table2 != nullptr) { ulonglong2 both = __ldg(v.table2 + idx); res = both.x; fused_jump = both.y; }
              else { res = __ldg(v.table + idx); }
              const u64 cnt = res >> 40;
              if(STATS) { st_hits++; }
              if(cnt == TABLE_ESCAPE) { hand[j] = at | WORK_FRESH; }
              else
              {
                const u64 s = res & M40, e = s + cnt - 1;
                rem[j] = (u32)len - k;
                if(cnt == 1 && rem[j] > 0)
                {
                  node[j] = s; active[j] = true; moved[j] = true;
                  // fused table: the long jump entry of the node came with the same load -- the first jump costs no probe
                  if(fused_jump != 0)
                  {
                    JumpPath path = jump_decode(fused_jump, v.jump_tbits);
                    if(path.len >= 2 && path.len <= rem[j] && k + path.len <= wgood[j] && k + path.len <= 32)
                    {
                      if((((win[j] >> (2 * k)) ^ path.chars) & ((1ull << (2 * path.len)) - 1)) == 0)
                      {
                        node[j] = path.target; rem[j] -= path.len;
                        if(STATS) { st_steps += path.len; }
                        if(rem[j] == 0)
                        {
                          active[j] = false;
                          __stcs((unsigned long long*)sp_out + at, (unsigned long long)node[j]); __stcs((unsigned long long*)ep_out + at, (unsigned long long)node[j]);
                          if(STATS) { st_found++; st_len++; }
                        }
                      }
                      else { active[j] = false; hand[j] = at | ((u64)rem[j] << 48) | WORK_NO_JUMP; }
                    }
                  }
                }
                else
                {
                  __stcs((unsigned long long*)sp_out + at, (unsigned long long)s); __stcs((unsigned long long*)ep_out + at, (unsigned long long)e);
                  if(cnt >= 2 && rem[j] > 0) { hand[j] = at | ((u64)rem[j] << 48); }
                  else if(STATS && cnt >= 1) { st_found++; st_len += cnt; }
                }
              }
            }
          }
        }
      }
      else if(at < n)
      {
        const u64 entry = work[at];
        q[j] = entry & WORK_QUERY_MASK; rem[j] = (u32)((entry >> 48) & 0xFF);
        origin[j] = (PACKED ? q[j] * (u64)((L + 31) >> 5) : q[j] * (u64)L);
        if((entry & (WORK_FRESH | WORK_NO_JUMP)) != 0 || rem[j] == 0) { hand[j] = entry; }
        else
        {
          const u64 s = sp_out[q[j]], e = ep_out[q[j]];
          if(s != e) { hand[j] = entry; }
          else { node[j] = s; active[j] = true; }
        }
      }
    }
    while(true)
    {
      bool any = false;
      #pragma unroll
      for(int j = 0; j < U; j++) { any = any || active[j]; }
      // (warp-uniform loop: the lanes leave together.)  Once fewer than `straggle` lanes of the warp still have an entry to
      // follow -- the long patterns of a batch of mixed lengths, the queries that met several branches -- those go on to the
      // general kernel with what they have reached: its refill loop keeps every lane busy, these rounds would not.
      const u32 busy = __ballot_sync(0xFFFFFFFFu, any);
      if(busy == 0) { break; }
      if((u32)__popc(busy) < straggle)
      {
        #pragma unroll
        for(int j = 0; j < U; j++)
        {
          if(active[j]) { active[j] = false; hand[j] = q[j] | ((u64)rem[j] << 48); }
        }
        break;
      }
      // ---- windows: at least min(rem, 16) characters in front of the current position, or the entry is handed on ----
      #pragma unroll
      for(int j = 0; j < U; j++)
      {
        if(active[j])
        {
          const u32 need = (rem[j] < 16 ? rem[j] : 16);
          if(wend[j] < rem[j] || wgood[j] < (wend[j] - rem[j]) + need)
          {
            chain_window<PACKED>(chars, origin[j], rem[j], &win[j], &win_hi[j], &wgood[j]); wend[j] = rem[j];
            if(wgood[j] < need) { active[j] = false; hand[j] = q[j] | ((u64)rem[j] << 48); }
          }
        }
      }
      // ---- one probe per entry, all issued before any is used ----
      u64 px[U], py[U], pz[U], pw[U]; u32 kind[U];                  // kind: 0 none, 1 long jump (8 bytes), 2 wide, 3 short jump, 4 sector
      #pragma unroll
      for(int j = 0; j < U; j++)
      {
        px[j] = py[j] = pz[j] = pw[j] = 0; kind[j] = 0;
        if(active[j])
        {
          if(!step_next[j] && have_long && rem[j] >= (u32)v.jump_k)
          {
            if(v.jump_wide != nullptr) { ulonglong2 e = __ldg(v.jump_wide + node[j]); px[j] = e.x; py[j] = e.y; kind[j] = 2; }
            else { px[j] = __ldg(v.jump + node[j]); kind[j] = 1; }
          }
          else if(!step_next[j] && v.jump_short != nullptr && rem[j] >= 4) { px[j] = __ldg(v.jump_short + node[j]); kind[j] = 3; }
          else
          {
            const u32 off = wend[j] - rem[j];
            const u32 c = (u32)((off < 32 ? win[j] >> (2 * off) : win_hi[j] >> (2 * (off - 32)))) & 3;
            const u64 b = node[j] / BWT_W;
            ulonglong4 s4 = ld256(v.bwt + b * 4 + c);
            px[j] = s4.x; py[j] = s4.y; pz[j] = s4.z; pw[j] = s4.w; kind[j] = 4;
          }
          if(STATS) { st_sectors++; }
        }
      }
      // ---- resolve ----
      #pragma unroll
      for(int j = 0; j < U; j++)
      {
        if(active[j])
        {
          const u32 off = wend[j] - rem[j];
          if(kind[j] == 4)
          {
            const u64 b = node[j] / BWT_W; const u32 o = (u32)(node[j] - b * BWT_W);
            const bool bit = (o < 64 ? (py[j] >> o) & 1 : ((px[j] >> 40) >> (o - 64)) & 1);
            if(!bit) { active[j] = false; hand[j] = q[j] | ((u64)rem[j] << 48) | WORK_NO_JUMP; }      // it fails here: the exact pair comes from the general kernel
            else
            {
              const u32 t = popc_low88(py[j], (u32)(px[j] >> 40), o);
              node[j] = (pz[j] & M40) + popc_low88(pw[j], (u32)(pz[j] >> 40), t + 1);
              rem[j]--; step_next[j] = false; moved[j] = true;
              if(STATS) { st_steps++; }
            }
          }
          else
          {
            JumpPath path = (kind[j] == 2 ? jump_decode_wide(make_ulonglong2(px[j], py[j])) : jump_decode(px[j], v.jump_tbits));
            if(path.len >= 2 && path.len <= rem[j])
            {
              // the next 32 characters of the pattern from the two window words
              const u64 ahead = (off == 0 ? win[j] : (off < 32 ? (win[j] >> (2 * off)) | (win_hi[j] << (64 - 2 * off)) : win_hi[j] >> (2 * (off - 32))));
              if(((ahead ^ path.chars) & ((1ull << (2 * path.len)) - 1)) == 0)
              {
                node[j] = path.target; rem[j] -= path.len; moved[j] = true;
                if(STATS) { st_steps += path.len; }
              }
              else { active[j] = false; hand[j] = q[j] | ((u64)rem[j] << 48) | WORK_NO_JUMP; }        // it dies within these steps
            }
            else { step_next[j] = true; }                              // the path branches here (or is longer than what is left): one single step
          }
          if(active[j] && rem[j] == 0)
          {
            active[j] = false;
            __stcs((unsigned long long*)sp_out + q[j], (unsigned long long)node[j]); __stcs((unsigned long long*)ep_out + q[j], (unsigned long long)node[j]);
            if(STATS) { st_found++; st_len++; }
          }
        }
      }
    }
    // ---- what is left goes to the second work list, with the node reached (one atomic per warp and round) ----
    u32 todo[U]; u32 total = 0;
    #pragma unroll
    for(int j = 0; j < U; j++)
    {
      if(hand[j] != 0 && moved[j])
      {
        __stcs((unsigned long long*)sp_out + q[j], (unsigned long long)node[j]); __stcs((unsigned long long*)ep_out + q[j], (unsigned long long)node[j]);
      }
      todo[j] = __ballot_sync(0xFFFFFFFFu, hand[j] != 0);
      total += __popc(todo[j]);
    }
    if(total != 0)
    {
      unsigned long long at = 0;
      if(lane == 0) { at = atomicAdd(work2_count, (unsigned long long)total); }
      at = __shfl_sync(0xFFFFFFFFu, at, 0);
      #pragma unroll
      for(int j = 0; j < U; j++)
      {
        if((todo[j] >> lane) & 1) { work2[at + __popc(todo[j] & ((1u << lane) - 1))] = hand[j]; }
        at += __popc(todo[j]);
      }
    }
  }
  if(STATS)
  {
    atomicAdd((ull*)&stats->found, (ull)st_found); atomicAdd((ull*)&stats->total_length, (ull)st_len);
    atomicAdd((ull*)&stats->lf_steps, (ull)st_steps); atomicAdd((ull*)&stats->sector_probes, (ull)st_sectors);
    if(FRESH) { atomicAdd((ull*)&stats->table_hits, (ull)st_hits); }
  }
}

#endif
