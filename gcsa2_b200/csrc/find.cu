/*
  find.cu -- find(): the launch logic of the kernels, the device- and host-buffer entry points with their pipeline, the multi-GPU forms.
  One of the CUDA translation units of libgcsa2_b200.so (see engine.h); host side of the C ABI of include/gcsa2_b200.h,
  kernels in the device/*.cuh it includes.
*/
#include "engine.h"
#include "device/find.cuh"
//------------------------------------------------------------------------------
// find
//------------------------------------------------------------------------------

// Measurement hook (not part of the C ABI): how many batches of this process took the two-kernel form.
static std::atomic<unsigned long long> g_fast_launches(0);
extern "C" unsigned long long gcsa_b200_internal_fast_launches(void) { return g_fast_launches.load(); }

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
  // chain kernel: a warp hands its last entries on to the general kernel once fewer than this many of its lanes are busy
  // (off: measured on 64-mers and on mixed lengths 16..256, 4 M queries on a 20 Mbp graph, thresholds 0 / 8 / 16 / 24:
  // 1.08 / 1.14 / 1.17 / 1.52 ms and 1.04 / 0.99 / 0.94 / 0.96 ms -- what the mixed batch gains the uniform one loses)
  static const u32 straggle = []() { const char* e = std::getenv("GCSA_B200_CHAIN_STRAGGLE"); int t = (e ? std::atoi(e) : 0); return (u32)std::min(32, std::max(0, t)); }();
  static const bool fast_off = []() { const char* e = std::getenv("GCSA_B200_FIND_FAST"); return (e != nullptr && std::atoi(e) == 0); }();
  const DevView& v = index->view;
  // Batches of patterns of any lengths (the offsets form), and of one length beyond the k-mer table plus one long jump: the
  // chain kernel starts every pattern from the k-mer table and follows single path nodes to the end; the general kernel
  // finishes what that leaves (GCSA_B200_FIND_CHAIN_OFFSETS=0: the general kernel alone, or the k-mer form below).  4 M
  // 64-mers on the configs[2] graph: 0.95 ms this way, 1.09 ms through the k-mer form, 1.33 ms through the general kernel.
  static const bool chain_offsets_off = []() { const char* e = std::getenv("GCSA_B200_FIND_CHAIN_OFFSETS"); return (e != nullptr && std::atoi(e) == 0); }();
  const u64 one_long_jump = (u64)v.table_k + ((v.jump_wide != nullptr || v.jump != nullptr) ? (u64)v.jump_k : 0);
  const bool long_fixed = (d_offsets == nullptr && fixed_length > one_long_jump && fixed_length <= 255);
  if(!fast_off && !chain_offsets_off && !packed && (d_offsets != nullptr || long_fixed) && v.table_k > 0 && v.default_alphabet != 0 && n >= 4096 && n < (1ull << 47))
  {
    u64* buffer = nullptr;
    CUDA_TRY(engineMallocAsync(&buffer, n * sizeof(u64) + 256, stream));
    u64* work = buffer;
    unsigned long long* count = (unsigned long long*)(work + n);
    cudaError_t e = cudaMemsetAsync(count, 0, 2 * sizeof(unsigned long long), stream);
    if(e == cudaSuccess)
    {
      int chain_grid = gridFor(n, index->sm_count, 4);
      int slow_grid = gridFor(n, index->sm_count, d_stats ? 1 : 4);
      if(d_stats)
      {
        find_chain_kernel<true, false, 1, true><<<chain_grid, 256, 0, stream>>>(v, d_chars, (u32)fixed_length, d_sp, d_ep, nullptr, nullptr, work, count, d_stats, straggle, d_offsets, char_base, n);
        find_kernel<true, 1, false, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, d_offsets, char_base, fixed_length, n, d_sp, d_ep, d_stats, refill_at, work, count);
      }
      else
      {
        find_chain_kernel<false, false, 1, true><<<chain_grid, 256, 0, stream>>>(v, d_chars, (u32)fixed_length, d_sp, d_ep, nullptr, nullptr, work, count, nullptr, straggle, d_offsets, char_base, n);
        find_kernel<false, 4, false, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, d_offsets, char_base, fixed_length, n, d_sp, d_ep, nullptr, refill_at, work, count);
      }
      e = cudaGetLastError();
    }
    cudaFreeAsync(buffer, stream);
    CUDA_TRY(e);
    if(long_fixed) { g_fast_launches.fetch_add(1); }
    return 0;
  }
  // Batches of k-mers (one length, at least the k-mer table's, the default alphabet): the two-kernel form -- one
  // probe (or two) per thread for everything, then the general kernel for the work list of what that left unfinished.
  if(!fast_off && d_offsets == nullptr && fixed_length <= 255 && v.table_k > 0 && fixed_length >= (u64)v.table_k &&
     (packed || v.default_alphabet != 0) && n >= 4096 && n < (1ull << 47))
  {
    // work list of the general kernel (8 bytes per entry), work list of the quad kernel (16), the two counters
    const bool use_quads = (v.jump_wide != nullptr || v.jump != nullptr);
    // (the 16-byte entries first: the allocation is aligned, the end of an odd number of 8-byte entries is not)
    // Patterns with more to go than one long jump after the table: the singleton entries of the work list are followed by
    // find_chain_kernel, which leaves a second work list to the general kernel (GCSA_B200_FIND_CHAIN=0: straight to it).
    static const bool chain_off = []() { const char* e = std::getenv("GCSA_B200_FIND_CHAIN"); return (e != nullptr && std::atoi(e) == 0); }();
    const u64 one_jump = (u64)v.table_k + (use_quads ? (u64)v.jump_k : 0);
    const bool use_chain = (!chain_off && fixed_length > one_jump);
    u64* buffer = nullptr;
    CUDA_TRY(engineMallocAsync(&buffer, n * sizeof(u64) * ((use_quads ? 3 : 1) + (use_chain ? 1 : 0)) + 256, stream));
    ulonglong2* quad_work = (use_quads ? (ulonglong2*)buffer : nullptr);
    u64* work = buffer + (use_quads ? 2 * n : 0);
    u64* work2 = (use_chain ? work + n : nullptr);
    unsigned long long* count = (unsigned long long*)(work + (use_chain ? 2 * n : n));
    unsigned long long* quad_count = count + 1;
    unsigned long long* count2 = count + 2;
    const u64* last_work = (use_chain ? work2 : work); const unsigned long long* last_count = (use_chain ? count2 : count);
    cudaError_t e = cudaMemsetAsync(count, 0, 4 * sizeof(unsigned long long), stream);
    if(e == cudaSuccess)
    {
      // queries per thread and round in the first kernel (GCSA_B200_FIND_UNROLL = 1, 2 or 4: measured in DESIGN.md)
      static const int unroll = []() { const char* e = std::getenv("GCSA_B200_FIND_UNROLL"); int u = (e ? std::atoi(e) : 4); return (u == 1 || u == 2 ? u : 4); }();
      int fast_grid = gridFor((n + unroll - 1) / unroll, index->sm_count, 8);
      int slow_grid = gridFor(n, index->sm_count, d_stats ? 1 : 4);
      // entries per thread and round in the chain kernel (GCSA_B200_CHAIN_UNROLL = 1, 2 or 4)
      static const int chain_unroll = []() { const char* e = std::getenv("GCSA_B200_CHAIN_UNROLL"); int u = (e ? std::atoi(e) : 1); return (u == 2 || u == 4 ? u : 1); }();
      int chain_grid = gridFor((n + chain_unroll - 1) / chain_unroll, index->sm_count, chain_unroll == 4 ? 2 : (chain_unroll == 2 ? 3 : 4));
      #define LAUNCH_CHAIN(S, P) do { \
        if(chain_unroll == 1) { find_chain_kernel<S, P, 1><<<chain_grid, 256, 0, stream>>>(v, d_chars, L, d_sp, d_ep, work, count, work2, count2, d_stats, straggle); } \
        else if(chain_unroll == 2) { find_chain_kernel<S, P, 2><<<chain_grid, 256, 0, stream>>>(v, d_chars, L, d_sp, d_ep, work, count, work2, count2, d_stats, straggle); } \
        else { find_chain_kernel<S, P, 4><<<chain_grid, 256, 0, stream>>>(v, d_chars, L, d_sp, d_ep, work, count, work2, count2, d_stats, straggle); } } while(0)
      const u32 L = (u32)fixed_length;
      #define LAUNCH_FAST(S, P, U) find_fast_kernel<S, P, U><<<fast_grid, 256, 0, stream>>>(v, d_chars, L, n, d_sp, d_ep, work, count, quad_work, quad_count, d_stats)
      #define LAUNCH_FAST_U(S, P) do { if(unroll == 1) { LAUNCH_FAST(S, P, 1); } else if(unroll == 2) { LAUNCH_FAST(S, P, 2); } else { LAUNCH_FAST(S, P, 4); } } while(0)
      if(d_stats)
      {
        if(packed) { LAUNCH_FAST(true, true, 4); } else { LAUNCH_FAST(true, false, 4); }
        if(use_quads) { find_quad_kernel<true><<<gridFor(n, index->sm_count, 8), 256, 0, stream>>>(v, L, quad_work, quad_count, d_sp, d_ep, work, count, d_stats); }
        if(use_chain)
        {
          if(packed) { LAUNCH_CHAIN(true, true); } else { LAUNCH_CHAIN(true, false); }
        }
        if(packed) { find_kernel<true, 1, true, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, d_stats, refill_at, last_work, last_count); }
        else { find_kernel<true, 1, false, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, d_stats, refill_at, last_work, last_count); }
      }
      else
      {
        if(packed) { LAUNCH_FAST_U(false, true); } else { LAUNCH_FAST_U(false, false); }
        if(use_quads) { find_quad_kernel<false><<<gridFor(n, index->sm_count, 8), 256, 0, stream>>>(v, L, quad_work, quad_count, d_sp, d_ep, work, count, nullptr); }
        if(use_chain)
        {
          if(packed) { LAUNCH_CHAIN(false, true); } else { LAUNCH_CHAIN(false, false); }
        }
        if(packed) { find_kernel<false, 4, true, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, nullptr, refill_at, last_work, last_count); }
        else { find_kernel<false, 4, false, true><<<slow_grid, 256, 0, stream>>>(v, d_chars, nullptr, 0, fixed_length, n, d_sp, d_ep, nullptr, refill_at, last_work, last_count); }
      }
      #undef LAUNCH_FAST_U
      #undef LAUNCH_CHAIN
      #undef LAUNCH_FAST
      e = cudaGetLastError();
    }
    cudaFreeAsync(buffer, stream);
    CUDA_TRY(e);
    g_fast_launches.fetch_add(1);
    return 0;
  }
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
  Host-buffer find.  The batch is cut into chunks that are pipelined over the slots of a HostPipe (the H2D copy of
  one chunk overlaps the kernel of another and the D2H copy of a third).  What bounds this entry point is the H2D
  copy of the patterns (32 pattern bytes in, 16 result bytes out per 32-mer), so fixed-length ACGT batches are also
  2-bit packed on the host (pack.cpp): 4x fewer bytes over the link -- for the chunks the host manages to pack.

  Raw copying and packing SHARE the batch.  The calling thread is the driver -- the only thread that talks to the
  CUDA runtime: it keeps a few raw chunks from the FRONT of the batch queued ahead of the copy engine and sends every
  packed chunk as soon as it is complete.  The other threads of its OpenMP team are packers: they work through the
  chunks the driver opens for them from the BACK of the batch, one sub-block of 8192 patterns at a time, into a ring
  of pinned staging buffers.  The two ends meet wherever the ratio of packing rate to link rate puts them (a packer as
  fast as the link leaves the batch at 0.57 of the raw transfer time, twice as fast at 0.4; with a slow host nearly
  everything goes raw).  Nobody waits for anybody: the first version had a helper thread for the raw copies that the
  packing team starved of a core, the second one packed and enqueued in turns on one thread (measured 7.1 and 5.5 ms
  per 10 M 32-mers, profiles/r02_bench_cfg2_*pack*.json).
  A chunk with any character other than ACGT/acgt is sent raw.
    GCSA_B200_HOST_PACK=0   no packing;   =N   N packing threads;
    unset or "auto"         all OpenMP threads (GCSA_B200_HOST_PACK_THREADS overrides the count).
*/
static int hostPackThreads()
{
  const char* e = std::getenv("GCSA_B200_HOST_PACK");
  if(e != nullptr && *e != 0 && std::strcmp(e, "auto") != 0) { return std::max(0, std::atoi(e)); }
  const char* t = std::getenv("GCSA_B200_HOST_PACK_THREADS");
  int threads = (t != nullptr && *t != 0 ? std::atoi(t) : omp_get_max_threads());
  return std::max(1, threads);
}

static inline void cpuRelax()
{
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}

// Measurement hook (not part of the C ABI): chunks of the last host-buffer find of this process that went packed, and all.
static std::atomic<unsigned long long> g_last_packed_chunks(0), g_last_chunks(0);
extern "C" void gcsa_b200_internal_pack_share(unsigned long long* packed, unsigned long long* total)
{
  if(packed) { *packed = g_last_packed_chunks.load(); }
  if(total) { *total = g_last_chunks.load(); }
}

static int findHost(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets, uint64_t fixed_length,
                    uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats, int pack_threads_override = -1)
{
  if(index == nullptr || (n > 0 && (chars == nullptr || sp == nullptr || ep == nullptr)))
  {
    return fail(GCSA_B200_ERR_INVALID, "find_host: null argument");
  }
  if(stats) { std::memset(stats, 0, sizeof(*stats)); stats->queries = n; }
  if(n == 0) { return 0; }
  DeviceGuard guard(index->device);

  const int pack_threads = (pack_threads_override >= 0 ? pack_threads_override : hostPackThreads());
  // (below three chunks of 128 k queries there is nothing to share)
  bool pack = (pack_threads > 0 && fixed_length > 0 && offsets == nullptr && stats == nullptr && n > (2u << 17));
  // No explicit policy: both ways are tried on the first large batches, then the faster one is used, and the other one
  // is tried again every 32nd batch (the load on the host changes).
  const char* policy_env = std::getenv("GCSA_B200_HOST_PACK");
  const bool auto_policy = pack && (policy_env == nullptr || *policy_env == 0 || std::strcmp(policy_env, "auto") == 0);
  if(auto_policy)
  {
    std::lock_guard<std::mutex> lock(index->policy_mutex);
    u64 call = index->policy_calls++;
    if(index->policy_seconds[1] == 0.0) { pack = true; }
    else if(index->policy_seconds[0] == 0.0) { pack = false; }
    else
    {
      bool best = (index->policy_seconds[1] <= index->policy_seconds[0]);
      pack = (call % 32 == 31 ? !best : best);
    }
  }
  const double policy_t0 = omp_get_wtime();
  const bool policy_packed = pack;
  // Chunks of >= 128 k queries (4 MB of 32-mers: the link is at its streaming rate), at most ~24 per batch (48 when
  // packing shares it): the H2D engine is the busy resource from the first byte on, so what the pipeline adds to
  // the transfer time is the kernel and the D2H of the LAST chunk -- the smaller the chunks, the smaller that tail.
  const u64 CHUNK = (pack ? std::max<u64>(1ull << 17, (n + 47) / 48) : std::max<u64>(1ull << 18, (n + 23) / 24));
  const u64 n_chunks = (n + CHUNK - 1) / CHUNK;
  const u64 words_per_pattern = (fixed_length + 31) / 32;
  const int SLOTS = HostPipe::SLOTS;

  HostPipe* pipe = index->takePipe();
  struct Return { const gcsa_b200_index* index; HostPipe* pipe; ~Return() { index->givePipe(pipe); } } give_back = { index, pipe };
  {
    cudaError_t e = pipe->init();
    if(e != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string("find_host: stream creation: ") + cudaGetErrorString(e)); }
  }
  FindStatsDev* d_stats = nullptr;
  if(stats)
  {
    if(cudaMalloc(&d_stats, sizeof(FindStatsDev)) != cudaSuccess || cudaMemset(d_stats, 0, sizeof(FindStatsDev)) != cudaSuccess)
    {
      if(d_stats) { cudaFree(d_stats); }
      return fail(GCSA_B200_ERR_CUDA, "find_host: out of device memory");
    }
  }

  int rc = 0;
  u64 issued = 0;                        // chunks enqueued so far: chunk number k uses slot k % SLOTS
  // One chunk through the next slot: H2D (raw bytes, or the words packed into the slot's staging buffer), kernel, D2H.
  auto enqueue = [&](u64 c, int staging_buffer) -> int
  {
    const bool packed = (staging_buffer >= 0);
    const int slot = (int)(issued % SLOTS);
    cudaStream_t st = pipe->stream[slot];
    u64 q0 = c * CHUNK, q1 = std::min(n, q0 + CHUNK), m = q1 - q0;
    u64 c0 = (offsets ? offsets[q0] : q0 * fixed_length), c1 = (offsets ? offsets[q1] : q1 * fixed_length);
    u64 bytes = (packed ? m * words_per_pattern * sizeof(u64) : c1 - c0);
    #define PIPE_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
      return fail(e_ == cudaErrorMemoryAllocation ? GCSA_B200_ERR_NOMEM : GCSA_B200_ERR_CUDA, std::string("find_host: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)
    if(pipe->used[slot]) { PIPE_TRY(cudaEventSynchronize(pipe->done[slot])); }       // the slot's previous chunk has left its buffers
    PIPE_TRY(HostPipe::grow(&pipe->d_in[slot], &pipe->in_bytes[slot], bytes + 16, false));
    PIPE_TRY(HostPipe::grow(&pipe->d_res[slot], &pipe->res_bytes[slot], 2 * m * sizeof(u64), false));
    if(offsets) { PIPE_TRY(HostPipe::grow(&pipe->d_off[slot], &pipe->off_bytes[slot], (m + 1) * sizeof(u64), false)); }
    u8* d_chars = (u8*)pipe->d_in[slot]; u64* d_off = (offsets ? (u64*)pipe->d_off[slot] : nullptr); u64* d_res = (u64*)pipe->d_res[slot];
    if(packed)
    {
      PIPE_TRY(cudaMemcpyAsync(d_chars, pipe->staging[staging_buffer], bytes, cudaMemcpyHostToDevice, st));
      PIPE_TRY(cudaEventRecord(pipe->staged[staging_buffer], st));       // the buffer may be packed into again
      pipe->staged_used[staging_buffer] = true;
    }
    else if(bytes) { PIPE_TRY(cudaMemcpyAsync(d_chars, chars + c0, bytes, cudaMemcpyHostToDevice, st)); }
    if(offsets) { PIPE_TRY(cudaMemcpyAsync(d_off, offsets + q0, (m + 1) * sizeof(u64), cudaMemcpyHostToDevice, st)); }
    PIPE_TRY(cudaEventRecord(pipe->copied[slot], st));
    int r = launchFind(index, d_chars, d_off, c0, fixed_length, m, d_res, d_res + m, d_stats, st, packed);
    if(r != 0) { return r; }
    PIPE_TRY(cudaMemcpyAsync(sp + q0, d_res, m * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PIPE_TRY(cudaMemcpyAsync(ep + q0, d_res + m, m * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PIPE_TRY(cudaEventRecord(pipe->done[slot], st));
    #undef PIPE_TRY
    pipe->used[slot] = true;
    issued++;
    return 0;
  };

  // Unclaimed chunks are [front, back): raw chunks are claimed from the front, packed ones from the back.  Only the
  // driver claims (it opens chunks for the packers), so front / back need no lock.
  u64 front = 0, back = n_chunks, packed_chunks = 0;
  std::vector<int> h2d_slots;            // slots of the chunks (raw or packed) whose H2D copy may still be queued, oldest first
  auto h2d_queued = [&]() -> size_t
  {
    while(!h2d_slots.empty() && cudaEventQuery(pipe->copied[h2d_slots.front()]) == cudaSuccess) { h2d_slots.erase(h2d_slots.begin()); }
    cudaGetLastError();                  // cudaErrorNotReady is not an error
    return h2d_slots.size();
  };
  auto send = [&](u64 c, int staging_buffer) -> int
  {
    const int slot = (int)(issued % SLOTS);
    int r = enqueue(c, staging_buffer);
    if(r == 0) { h2d_slots.push_back(slot); }
    return r;
  };
  auto send_raw = [&]() -> int
  {
    int r = send(front, -1);
    if(r == 0) { front++; }
    return r;
  };

  const int team = (pack ? std::min(pack_threads + 1, std::max(2, omp_get_max_threads())) : 1);     // the driver and the packers
  if(team < 2)
  {
    while(front < back && rc == 0) { rc = send_raw(); }
  }
  else
  {
    // Packed chunk j (the j-th from the back) is chunk n_chunks - 1 - j and uses staging buffer j % STAGING.
    const int STAGING = HostPipe::STAGING;
    const u64 SUB = 8192;                                            // patterns per work item
    const u64 subs_per_chunk = (CHUNK + SUB - 1) / SUB;
    std::vector<std::atomic<u32>> blocks_done(n_chunks), blocks_bad(n_chunks);
    for(u64 j = 0; j < n_chunks; j++) { blocks_done[j].store(0); blocks_bad[j].store(0); }
    std::atomic<u64> ticket(0), opened(0);
    std::atomic<bool> closing(false);
    for(int b = 0; b < STAGING && rc == 0; b++)
    {
      cudaError_t e = HostPipe::grow(&pipe->staging[b], &pipe->staging_bytes[b], CHUNK * words_per_pattern * sizeof(u64), true);
      if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("find_host: staging buffer: ") + cudaGetErrorString(e)); }
    }
    // A raw chunk is sent only when the copy engine is about to run dry (fewer than this many H2D copies queued,
    // packed ones included): every chunk the packers finish in time crosses the link at a quarter of the bytes, and
    // the raw chunks fill the gaps they leave.  (Keeping raw copies queued regardless gave the raw path half of the
    // batch however fast the packers were: profiles/r02_bench_cfg2_pipe3_pack_*.json.)
    const size_t feed_below = 2;

    #pragma omp parallel num_threads(team)
    {
      if(omp_get_thread_num() != 0)
      {
        // ---- packer: work items (chunk j, sub-block b) in order; wait until the driver has opened chunk j ----
        while(true)
        {
          u64 t = ticket.fetch_add(1), j = t / subs_per_chunk, b = t % subs_per_chunk;
          u32 spins = 0;
          while(j >= opened.load(std::memory_order_acquire) && !closing.load(std::memory_order_acquire))
          {
            if(++spins < 2000) { cpuRelax(); } else { std::this_thread::yield(); }
          }
          if(j >= opened.load(std::memory_order_acquire)) { break; }                       // closing: no more chunks
          u64 c = n_chunks - 1 - j, q0 = c * CHUNK, m = std::min(n, q0 + CHUNK) - q0;
          u64 first = b * SUB, last = std::min(m, first + SUB);
          if(first < last)
          {
            int good = gcsa_b200_internal_pack_range(chars + q0 * fixed_length, first, last, fixed_length, index->pack_code,
                                                     index->pack_default ? 1 : 0, (u64*)pipe->staging[j % STAGING]);
            if(!good) { blocks_bad[j].fetch_add(1, std::memory_order_relaxed); }
          }
          blocks_done[j].fetch_add(1, std::memory_order_release);
        }
      }
      else
      {
        // ---- driver ----
        u64 sent = 0;                                                // packed chunks handed to the copy engine
        while(rc == 0 && (front < back || sent < opened.load(std::memory_order_relaxed)))
        {
          bool progress = false;
          // a packed chunk is complete: send it (raw from the caller's buffer if it held another character)
          if(sent < opened.load(std::memory_order_relaxed) && blocks_done[sent].load(std::memory_order_acquire) == subs_per_chunk)
          {
            bool ok = (blocks_bad[sent].load() == 0);
            rc = send(n_chunks - 1 - sent, ok ? (int)(sent % STAGING) : -1);
            if(rc == 0 && ok) { packed_chunks++; }
            sent++;
            continue;
          }
          // open the next chunk for the packers: one being packed and one waiting is enough to keep them busy, and its
          // staging buffer must have been read by the copy engine (the packed chunk STAGING places before it)
          u64 open_now = opened.load(std::memory_order_relaxed);
          if(front < back && open_now - sent < 2)
          {
            const int buffer = (int)(open_now % STAGING);            // last used by packed chunk open_now - STAGING < sent
            bool free_buffer = true;
            if(pipe->staged_used[buffer])
            {
              if(cudaEventQuery(pipe->staged[buffer]) == cudaSuccess) { pipe->staged_used[buffer] = false; }
              else { free_buffer = false; cudaGetLastError(); }
            }
            if(free_buffer) { back--; opened.store(open_now + 1, std::memory_order_release); progress = true; }
          }
          // keep the copy engine fed
          if(front < back && h2d_queued() < feed_below) { rc = send_raw(); progress = true; }
          if(!progress) { cpuRelax(); }
        }
        closing.store(true, std::memory_order_release);
      }
    }
  }
  g_last_packed_chunks.store(packed_chunks); g_last_chunks.store(n_chunks);

  // everything that was enqueued must have left the caller's buffers before this returns, error or not
  cudaError_t err = cudaSuccess;
  for(int s = 0; s < SLOTS; s++)
  {
    if(pipe->stream[s]) { cudaError_t e = cudaStreamSynchronize(pipe->stream[s]); if(e != cudaSuccess) { err = e; } }
  }
  if(stats && err == cudaSuccess && rc == 0)
  {
    FindStatsDev h;
    err = cudaMemcpy(&h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    stats->found = h.found; stats->total_length = h.total_length; stats->lf_steps = h.lf_steps;
    stats->sector_probes = h.sector_probes; stats->table_hits = h.table_hits;
  }
  if(d_stats) { cudaFree(d_stats); }
  if(rc) { return rc; }
  if(err != cudaSuccess) { return fail(GCSA_B200_ERR_CUDA, std::string("find_host: ") + cudaGetErrorString(err)); }
  if(auto_policy)
  {
    // seconds per query of this batch; a moving average over the batches sent the same way
    double per_query = (omp_get_wtime() - policy_t0) / (double)n;
    std::lock_guard<std::mutex> lock(index->policy_mutex);
    double& slot = index->policy_seconds[policy_packed ? 1 : 0];
    slot = (slot == 0.0 ? per_query : 0.75 * slot + 0.25 * per_query);
  }
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

int gcsa_b200_find_fixed_stats_host(const gcsa_b200_index* index, const uint8_t* chars, uint64_t pattern_length,
                                    uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats)
{
  if(stats == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_fixed_stats_host: null stats"); }
  return findHost(index, chars, nullptr, pattern_length, n, sp, ep, stats);
}

int gcsa_b200_find_stats_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                              uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats)
{
  if(stats == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_stats_host: null stats"); }
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_stats_host: null offsets"); }
  return findHost(index, chars, offsets, 0, n, sp, ep, stats);
}

//------------------------------------------------------------------------------
// One process, several GPUs: the batch is cut into contiguous blocks, one per handle (each handle on its own
// device, the index replicated), and one host thread per handle runs the single-device pipeline on its block,
// writing straight into the caller's arrays.  This is what a caller that parallelises over queries with OpenMP
// threads in one process (src/algorithms.cpp:113, 409; vg) can use; there is no exchange between the devices.
//------------------------------------------------------------------------------


int gcsa_b200_find_fixed_host_multi(const gcsa_b200_index* const* indexes, int count, const uint8_t* chars, uint64_t pattern_length,
                                    uint64_t n, uint64_t* sp, uint64_t* ep)
{
  int rc = checkHandles(indexes, count, "find_fixed_host_multi");
  if(rc != 0) { return rc; }
  if(count == 1) { return findHost(indexes[0], chars, nullptr, pattern_length, n, sp, ep, nullptr); }
  const int per_handle = std::max(1, hostPackThreads() / count);       // the packing threads are shared out
  return runPerHandle(count, "find_fixed_host_multi", [&](int g) -> int
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    if(q0 == q1) { return 0; }
    return findHost(indexes[g], chars + q0 * pattern_length, nullptr, pattern_length, q1 - q0, sp + q0, ep + q0, nullptr,
                    hostPackThreads() == 0 ? 0 : per_handle);
  });
}

int gcsa_b200_find_host_multi(const gcsa_b200_index* const* indexes, int count, const uint8_t* chars, const uint64_t* offsets,
                              uint64_t n, uint64_t* sp, uint64_t* ep)
{
  int rc = checkHandles(indexes, count, "find_host_multi");
  if(rc != 0) { return rc; }
  if(n > 0 && offsets == nullptr) { return fail(GCSA_B200_ERR_INVALID, "find_host_multi: null offsets"); }
  if(count == 1) { return findHost(indexes[0], chars, offsets, 0, n, sp, ep, nullptr); }
  return runPerHandle(count, "find_host_multi", [&](int g) -> int
  {
    u64 q0, q1; shardBlock(n, count, g, &q0, &q1);
    if(q0 == q1) { return 0; }
    return findHost(indexes[g], chars, offsets + q0, 0, q1 - q0, sp + q0, ep + q0, nullptr);      // offsets stay batch-wide
  });
}

