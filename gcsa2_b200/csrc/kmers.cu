/*
  kmers.cu -- countKMers / compareKMers: frontier expansion over the trie of k-mers.
  One of the CUDA translation units of libgcsa2_b200.so (see engine.h); host side of the C ABI of include/gcsa2_b200.h,
  kernels in the device/*.cuh it includes.
*/
#include "engine.h"
#include "device/kmers.cuh"
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
  // Only the number is wanted: the frontier need not stay ordered, one kernel per level (kmer_level_kernel).  The
  // frontier of a level is at most `chars` times the one before, and the two buffers are sized for that.
  static const bool unordered_off = []() { const char* s_ = std::getenv("GCSA_B200_KMERS_ORDERED"); return (s_ != nullptr && std::atoi(s_) != 0); }();
  ulonglong2 *cur = nullptr, *next_buf = nullptr; unsigned long long* counter = nullptr;
  if(ranges == nullptr && !unordered_off)
  {
    u64 cur_capacity = 0, next_capacity = 0;
    {
      ulonglong2 root = make_ulonglong2(0, index->header.path_nodes - 1);
      KM_TRY(engineMallocAsync(&cur, sizeof(ulonglong2), st)); cur_capacity = 1;
      KM_TRY(engineMallocAsync(&counter, sizeof(unsigned long long), st));
      KM_TRY(cudaMemcpyAsync(cur, &root, sizeof(root), cudaMemcpyHostToDevice, st));
      for(u64 level = 0; level < k && n > 0; level++)
      {
        u64 want = n * chars;
        if(next_buf == nullptr || next_capacity < want)
        {
          if(next_buf) { cudaFreeAsync(next_buf, st); next_buf = nullptr; }
          KM_TRY(engineMallocAsync(&next_buf, want * sizeof(ulonglong2), st)); next_capacity = want;
        }
        KM_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
        kmer_level_kernel<<<gridFor(n, index->sm_count), 256, 0, st>>>(index->view, cur, n, chars, next_buf, counter, next_capacity);
        unsigned long long produced = 0;
        KM_TRY(cudaMemcpyAsync(&produced, counter, sizeof(produced), cudaMemcpyDeviceToHost, st));
        KM_TRY(cudaStreamSynchronize(st));
        std::swap(cur, next_buf); std::swap(cur_capacity, next_capacity);
        n = produced;
      }
      *result = n;
    }
    goto done;
  }
  {
    u64 root[2] = { 0, index->header.path_nodes - 1 };
    KM_TRY(engineMallocAsync(&sp, sizeof(u64), st)); KM_TRY(engineMallocAsync(&ep, sizeof(u64), st));
    KM_TRY(cudaMemcpyAsync(sp, &root[0], sizeof(u64), cudaMemcpyHostToDevice, st));
    KM_TRY(cudaMemcpyAsync(ep, &root[1], sizeof(u64), cudaMemcpyHostToDevice, st));
    for(u64 level = 0; level < k && n > 0; level++)
    {
      u64 total = n * chars;
      u64 *csp = nullptr, *cep = nullptr, *flag = nullptr, *pos = nullptr;
      KM_TRY(engineMallocAsync(&csp, total * sizeof(u64), st)); KM_TRY(engineMallocAsync(&cep, total * sizeof(u64), st));
      KM_TRY(engineMallocAsync(&flag, (total + 1) * sizeof(u64), st)); KM_TRY(engineMallocAsync(&pos, (total + 1) * sizeof(u64), st));
      KM_TRY(cudaMemsetAsync(flag + total, 0, sizeof(u64), st));
      kmer_expand_kernel<<<gridFor(total, index->sm_count), 256, 0, st>>>(index->view, sp, ep, n, chars, csp, cep, flag);
      rc = scanExclusive(flag, pos, total + 1, st);
      if(rc) { goto done; }
      u64 next = 0;
      KM_TRY(cudaMemcpyAsync(&next, pos + total, sizeof(u64), cudaMemcpyDeviceToHost, st));
      KM_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(sp, st); cudaFreeAsync(ep, st); sp = ep = nullptr;
      KM_TRY(engineMallocAsync(&sp, std::max<u64>(next, 1) * sizeof(u64), st)); KM_TRY(engineMallocAsync(&ep, std::max<u64>(next, 1) * sizeof(u64), st));
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
  if(cur) { cudaFreeAsync(cur, st); }
  if(next_buf) { cudaFreeAsync(next_buf, st); }
  if(counter) { cudaFreeAsync(counter, st); }
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
  u64* state = nullptr; u64* kmer = nullptr; unsigned long long* counter = nullptr;
  u64 stride = 1;                          // entries per array of `state`
  cudaError_t e = cudaSuccess;
  #define CK_TRY(expr) do { e = (expr); if(e != cudaSuccess) { rc = fail(GCSA_B200_ERR_CUDA, std::string("compare_kmers: " #expr ": ") + cudaGetErrorString(e)); goto done; } } while(0)
  #define CK_ALLOC(ptr, count) do { CK_TRY(engineMallocAsync((void**)&(ptr), std::max<u64>((count), 1) * sizeof(u64), st)); } while(0)
  {
    CK_ALLOC(state, 4); CK_TRY(cudaMemcpyAsync(state, root, sizeof(root), cudaMemcpyHostToDevice, st));
    if(want) { CK_ALLOC(kmer, 3); CK_TRY(cudaMemcpyAsync(kmer, zero_kmer, sizeof(zero_kmer), cudaMemcpyHostToDevice, st)); }
    CK_TRY(engineMallocAsync((void**)&counter, sizeof(unsigned long long), st));
    for(u64 level = 0; level < k && n > 0; level++)
    {
      // one kernel per level: the next frontier holds at most `chars` children per state
      u64 capacity = n * chars;
      u64 *new_state = nullptr, *new_kmer = nullptr;
      CK_ALLOC(new_state, 4 * capacity);
      if(want) { CK_ALLOC(new_kmer, 3 * capacity); }
      CK_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
      compare_level_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(left->view, right->view, state, n, stride, kmer, chars, level,
                                                                        new_state, capacity, new_kmer, counter);
      unsigned long long produced = 0;
      CK_TRY(cudaMemcpyAsync(&produced, counter, sizeof(produced), cudaMemcpyDeviceToHost, st));
      CK_TRY(cudaStreamSynchronize(st));
      cudaFreeAsync(state, st);
      if(kmer) { cudaFreeAsync(kmer, st); }
      state = new_state; kmer = new_kmer; n = produced; stride = capacity;
    }
    if(n > 0)
    {
      ull* counts = nullptr; u64 *lflag = nullptr, *rflag = nullptr, *lpos = nullptr, *rpos = nullptr;
      CK_TRY(engineMallocAsync((void**)&counts, 3 * sizeof(ull), st));
      CK_TRY(cudaMemsetAsync(counts, 0, 3 * sizeof(ull), st));
      if(want)
      {
        CK_ALLOC(lflag, n + 1); CK_ALLOC(rflag, n + 1); CK_ALLOC(lpos, n + 1); CK_ALLOC(rpos, n + 1);
        CK_TRY(cudaMemsetAsync(lflag + n, 0, sizeof(u64), st)); CK_TRY(cudaMemsetAsync(rflag + n, 0, sizeof(u64), st));
      }
      compare_classify_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(state, n, stride, counts, lflag, rflag);
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
          compare_emit_kernel<<<gridFor(n, left->sm_count), 256, 0, st>>>(state, stride, kmer, n, k, side == 0 ? lflag : rflag, side == 0 ? lpos : rpos, records);
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
  if(counter) { cudaFreeAsync(counter, st); }
  #undef CK_TRY
  #undef CK_ALLOC
  if(rc != 0)
  {
    if(left_kmers && *left_kmers) { std::free(*left_kmers); *left_kmers = nullptr; }
    if(right_kmers && *right_kmers) { std::free(*right_kmers); *right_kmers = nullptr; }
  }
  HOST_EPILOGUE("compare_kmers", rc);
}


/*
  compareKMers with KMerSearchParameters::output set (include/gcsa/algorithms.h:59-71, src/algorithms.cpp:556-613): the
  k-mers unique to either index are written to <output>.left / <output>.right as the reference writes them -- raw
  KMerComparisonState records, 64 bytes each (both ranges, k, the k-mer in three words), in no particular order.
*/
int gcsa_b200_compare_kmers_to_files(const gcsa_b200_index* left, const gcsa_b200_index* right, uint64_t k, int include_Ns,
                                     const char* output, uint64_t* result)
{
  if(output == nullptr || *output == 0) { return gcsa_b200_compare_kmers(left, right, k, include_Ns, result, nullptr, nullptr); }
  const std::string left_name = std::string(output) + ".left", right_name = std::string(output) + ".right";   // KMerSearchParameters::LEFT_EXTENSION / RIGHT_EXTENSION
  FILE* left_file = std::fopen(left_name.c_str(), "wb");
  if(left_file == nullptr) { return fail(GCSA_B200_ERR_INVALID, "compare_kmers: cannot open output file " + left_name); }
  FILE* right_file = std::fopen(right_name.c_str(), "wb");
  if(right_file == nullptr) { std::fclose(left_file); return fail(GCSA_B200_ERR_INVALID, "compare_kmers: cannot open output file " + right_name); }
  gcsa_b200_kmer_state *left_kmers = nullptr, *right_kmers = nullptr;
  int rc = gcsa_b200_compare_kmers(left, right, k, include_Ns, result, &left_kmers, &right_kmers);
  if(rc == 0)
  {
    bool ok = (result[1] == 0 || std::fwrite(left_kmers, sizeof(gcsa_b200_kmer_state), result[1], left_file) == result[1]);
    ok = ok && (result[2] == 0 || std::fwrite(right_kmers, sizeof(gcsa_b200_kmer_state), result[2], right_file) == result[2]);
    if(!ok) { rc = fail(GCSA_B200_ERR_INVALID, "compare_kmers: writing the output files failed"); }
  }
  std::free(left_kmers); std::free(right_kmers);
  if(std::fclose(left_file) != 0 && rc == 0) { rc = fail(GCSA_B200_ERR_INVALID, "compare_kmers: closing " + left_name + " failed"); }
  if(std::fclose(right_file) != 0 && rc == 0) { rc = fail(GCSA_B200_ERR_INVALID, "compare_kmers: closing " + right_name + " failed"); }
  return rc;
}
