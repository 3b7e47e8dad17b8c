/* device/tables.cuh -- construction kernels of the optional per-node and per-k-mer tables: locate walk / locate tables,
   jump tables (8- and 16-byte entries), k-mer table.  Included by engine.cu only (index creation); sm_100a only. */
#ifndef GCSA2_B200_DEVICE_TABLES_CUH
#define GCSA2_B200_DEVICE_TABLES_CUH

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

// The long table with 16-byte entries (indexes with more than 2^27 path nodes: an 8-byte entry has no room for 16
// characters next to a node number of that size).  Same construction: level 1 from `one`, one more step per round.
__global__ void __launch_bounds__(256)
jump_wide_init_kernel(u64 n, u32 tbits, const u64* __restrict__ one, ulonglong2* __restrict__ wide)
{
  const u64 tmask = (1ull << tbits) - 1;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    u64 e = one[i];
    wide[i] = ((e >> 59) == 0 ? make_ulonglong2(0, 0) : make_ulonglong2((e & tmask) | (1ull << 40), (e >> tbits) & 3));
  }
}

__global__ void __launch_bounds__(256)
jump_wide_extend_kernel(u64 n, u32 tbits, u32 j, const u64* __restrict__ one, ulonglong2* __restrict__ wide)
{
  const u64 tmask = (1ull << tbits) - 1;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
  {
    ulonglong2 e = wide[i];
    if(((e.x >> 40) & 63u) != j) { continue; }
    u64 next = __ldg(one + (e.x & M40));
    if((next >> 59) == 0) { continue; }
    wide[i] = make_ulonglong2((next & tmask) | ((u64)(j + 1) << 40), e.y | (((next >> tbits) & 3) << (2 * j)));
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

#endif
