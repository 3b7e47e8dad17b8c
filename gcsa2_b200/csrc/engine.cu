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
    (sp in 40 bits, range length in 24 bits; an empty result always has ep = sp - 1); in its fused
    form 16 bytes: the same entry and the jump-table entry of its path node.
  * optional per-node tables: jump tables for find() (the unary backward path of a node, up to 16 and up
    to 4 steps), walk / locate tables for locate().

  This file is the host side (handles, construction of the layout, C ABI entry points and their
  pipelines); the kernels are in device/*.cuh, included below in dependency order.
*/#include "engine.h"
#include "device/two_step.cuh"
#include "device/tables.cuh"

thread_local std::string g_last_error;

namespace {
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};
}

cudaError_t enginePoolAlloc(void** p, size_t bytes, cudaStream_t stream)
{
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if(e != cudaSuccess) { return e; }
  if(device < 0 || device >= 64) { return cudaErrorInvalidValue; }
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if(g_pools[device] == nullptr)
    {
      cudaMemPoolProps props;
      std::memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      e = cudaMemPoolCreate(&g_pools[device], &props);
      if(e != cudaSuccess) { g_pools[device] = nullptr; return e; }
      uint64_t keep = ~0ull;      // do not hand the memory back to the driver between calls
      cudaMemPoolSetAttribute(g_pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = g_pools[device];
  }
  return cudaMallocFromPoolAsync(p, bytes, pool, stream);
}

// Hands the pool's idle memory of the current device back to the driver.  Called where the free memory decides what
// gets built (index creation sizes its optional tables by it): the scratch of an earlier query batch -- the MEM-style
// scan keeps up to a quarter of the device -- must not count as used.
void enginePoolTrim()
{
  int device = 0;
  if(cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) { return; }
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    pool = g_pools[device];
  }
  if(pool == nullptr) { return; }
  cudaDeviceSynchronize();
  cudaMemPoolTrimTo(pool, 0);
}

// free device memory as a query batch sees it: what the driver reports plus what the pool holds idle
size_t engineFreeMemory()
{
  size_t free_b = 0, total_b = 0;
  if(cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
  int device = 0;
  if(cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) { return free_b; }
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    pool = g_pools[device];
  }
  if(pool != nullptr)
  {
    uint64_t reserved = 0, used = 0;
    if(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
       cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used) { free_b += (size_t)(reserved - used); }
    else { cudaGetLastError(); }
  }
  return free_b;
}

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
  enginePoolTrim();           // the optional tables are sized by the free memory: scratch kept from earlier batches is free

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

  // jump tables (optional; 8 bytes per path node each + as much again while they are built)
  {
    int want = (options != nullptr ? options->jump_table : 0);        // 0 = automatic, 1 = build, 2 = build with 16-byte entries, -1 = do not
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t bytes = (size_t)N * sizeof(u64);
    u32 tbits = 1; while((1ull << tbits) < N) { tbits++; }
    int max_len = std::min<int>(16, (59 - (int)tbits) / 2);
    // An 8-byte entry holds (59 - tbits) / 2 characters: 16 up to 2^27 path nodes, 13 at 3 G.  Beyond that the long
    // table gets 16-byte entries (16 characters again: a 32-mer is the k-mer table and one jump), memory permitting.
    // (while it is built: level 1, the short table and the 16-byte table, 4 x 8 bytes per node; the k-mer table comes after)
    size_t kmer_table_bytes = 0;
    if(options != nullptr && options->kmer_table_k > 0)
    {
      int tk = std::min(16, options->kmer_table_k);
      kmer_table_bytes = ((size_t)1 << (2 * tk)) * sizeof(u64) + ((size_t)1 << (2 * std::max(1, tk - 1))) * sizeof(ulonglong2);
    }
    bool wide = (want == 2 || (want >= 0 && max_len < 16 && (double)(4 * bytes + kmer_table_bytes) < 0.9 * (double)free_b));
    const int short_len = 4;
    if(N > 0 && want >= 0 && (max_len >= 2 || wide) && (want > 0 || wide || 2 * bytes < free_b / 2))
    {
      u64 *one = nullptr, *table = nullptr, *short_table = nullptr; ulonglong2* wide_table = nullptr;
      cudaError_t e = cudaMalloc((void**)&one, bytes);
      if(e == cudaSuccess) { e = cudaMalloc((void**)&table, bytes); }
      if(e == cudaSuccess && wide) { e = cudaMalloc((void**)&wide_table, 2 * bytes); }
      if(e == cudaSuccess && !wide && max_len > short_len && 3 * bytes < free_b / 2) { if(cudaMalloc((void**)&short_table, bytes) != cudaSuccess) { short_table = nullptr; cudaGetLastError(); } }
      if(e == cudaSuccess)
      {
        jump_init_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(v, tbits, one, table);
        // with 16-byte entries for the long paths, `table` only grows to the short length and IS the short table
        const int grow_to = (wide ? std::min(short_len, max_len) : max_len);
        for(int j = 1; j < grow_to; j++)
        {
          if(j == short_len && short_table != nullptr) { e = cudaMemcpyAsync(short_table, table, bytes, cudaMemcpyDeviceToDevice, 0); }   // paths of up to 4 steps
          jump_extend_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, (u32)j, one, table);
        }
        if(wide)
        {
          jump_wide_init_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, one, wide_table);
          for(int j = 1; j < 16; j++) { jump_wide_extend_kernel<<<gridFor(N, idx->sm_count, 8), 256>>>(N, tbits, (u32)j, one, wide_table); }
        }
        if(e == cudaSuccess) { e = cudaDeviceSynchronize(); }
      }
      if(one) { cudaFree(one); }
      if(e == cudaSuccess)
      {
        idx->allocations.push_back(table); idx->device_bytes += bytes;
        v.jump_tbits = tbits;
        if(wide)
        {
          idx->allocations.push_back(wide_table); idx->device_bytes += 2 * bytes;
          v.jump_wide = wide_table; v.jump_short = table; v.jump_k = 16;
        }
        else
        {
          v.jump = table; v.jump_k = (u32)max_len;
          if(short_table != nullptr) { idx->allocations.push_back(short_table); idx->device_bytes += bytes; v.jump_short = short_table; }
        }
      }
      else
      {
        if(table) { cudaFree(table); }
        if(short_table) { cudaFree(short_table); }
        if(wide_table) { cudaFree(wide_table); }
        cudaGetLastError();
        if(want > 0) { gcsa_b200_index_destroy(idx); return fail(GCSA_B200_ERR_CUDA, std::string("jump table: ") + cudaGetErrorString(e)); }
      }
    }
  }

  // two-step blocks (optional): built on the device from the one-step blocks
  // -1 = automatic: worth it once the one-step blocks are far beyond the L2 (the probe rate no longer
  // depends on the footprint there, so halving the probes halves the time; measured in DESIGN.md)
  // (measured at 3 G nodes, profiles/r02_cfg4_3gbp_option_variants.json: with the jump tables the two-step blocks LOSE --
  // the single steps left over are few and the 17.7 GB of extra sectors only dilute the TLB reach -- so the automatic
  // choice takes them only for an index without jump tables)
  bool want_two_step = (options != nullptr && (options->two_step > 0 ||
                        (options->two_step < 0 && N >= 400000000ull && v.jump == nullptr && v.jump_wide == nullptr)));
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
  for(HostPipe* p : index->pipes) { p->destroy(); delete p; }
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
  info->jump_k = (index->view.jump != nullptr || index->view.jump_wide != nullptr ? (int)index->view.jump_k : 0);
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

