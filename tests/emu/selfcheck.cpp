/*
  selfcheck.cpp -- TEST INFRASTRUCTURE: the host emulation of tests/emu/cuda_emu.h checked against what the CUDA
  constructs are defined to do, on kernels small enough to verify by hand: warp ballots with early-exited lanes,
  shuffles, a block-wide barrier with shared memory, atomics across blocks, the intrinsics whose semantics are easy to
  get wrong (__fns, __brev, __clzll, __vcmpltu4), the scan / segmented-sort stand-ins, and the order in which blocks
  and threads run with and without GCSA_EMU_SHUFFLE.  Prints "selfcheck OK" or the first failing line.
*/
#define EMU_DEFINE_SWITCH
#include "cuda_emu.h"

#include <cstdio>
#include <numeric>
#include <vector>

#define REQUIRE(cond) do { if(!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while(0)

// every lane votes; odd warps lose their upper half before the vote
__global__ void ballot_kernel(unsigned* out)
{
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = tid >> 5;
  if((warp & 1) && lane >= 16) { return; }
  unsigned votes = __ballot_sync(0xFFFFFFFFu, (lane % 3) == 0);
  out[tid] = votes;
}

__global__ void shuffle_kernel(unsigned long long* sums, unsigned* rotated)
{
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  unsigned long long x = tid;
  for(int d = 16; d > 0; d >>= 1) { x += __shfl_down_sync(0xFFFFFFFFu, x, d); }
  if(lane == 0) { sums[tid >> 5] = x; }
  rotated[tid] = __shfl_sync(0xFFFFFFFFu, tid, (lane + 1) & 31);
}

// reverse a block's values through shared memory
__global__ void barrier_kernel(const unsigned* in, unsigned* out)
{
  __shared__ unsigned buffer[128];
  buffer[threadIdx.x] = in[blockIdx.x * blockDim.x + threadIdx.x];
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = buffer[blockDim.x - 1 - threadIdx.x];
}

__global__ void order_kernel(unsigned* ticket, unsigned* order)
{
  order[atomicAdd(ticket, 1u)] = blockIdx.x * blockDim.x + threadIdx.x;
}

int main()
{
  // ---- ballots ----
  {
    const unsigned blocks = 3, threads = 96, n = blocks * threads;
    unsigned* out = nullptr; REQUIRE(cudaMalloc(&out, n * sizeof(unsigned)) == cudaSuccess);
    cudaMemset(out, 0xFF, n * sizeof(unsigned));
    emu::launch(emu::Cfg(blocks, threads), [&] { ballot_kernel(out); }, true);
    unsigned full = 0; for(unsigned l = 0; l < 32; l += 3) { full |= 1u << l; }
    for(unsigned t = 0; t < n; t++)
    {
      unsigned warp = t >> 5, lane = t & 31;
      if((warp & 1) && lane >= 16) { REQUIRE(out[t] == 0xFFFFFFFFu); }             // exited before the vote: nothing written
      else { REQUIRE(out[t] == ((warp & 1) ? (full & 0xFFFFu) : full)); }          // exited lanes do not vote
    }
    cudaFree(out);
  }
  // ---- shuffles ----
  {
    const unsigned blocks = 2, threads = 64, n = blocks * threads;
    unsigned long long* sums = nullptr; unsigned* rotated = nullptr;
    cudaMalloc(&sums, (n / 32) * sizeof(unsigned long long)); cudaMalloc(&rotated, n * sizeof(unsigned));
    emu::launch(emu::Cfg(blocks, threads), [&] { shuffle_kernel(sums, rotated); }, true);
    for(unsigned w = 0; w < n / 32; w++) { REQUIRE(sums[w] == 32ull * (32 * w) + 496); }
    for(unsigned t = 0; t < n; t++) { REQUIRE(rotated[t] == (t & ~31u) + ((t + 1) & 31)); }
    cudaFree(sums); cudaFree(rotated);
  }
  // ---- block barrier + shared memory ----
  {
    const unsigned blocks = 4, threads = 128, n = blocks * threads;
    unsigned *in = nullptr, *out = nullptr; cudaMalloc(&in, n * sizeof(unsigned)); cudaMalloc(&out, n * sizeof(unsigned));
    for(unsigned i = 0; i < n; i++) { in[i] = 7 * i + 1; }
    emu::launch(emu::Cfg(blocks, threads), [&] { barrier_kernel(in, out); }, true);
    for(unsigned i = 0; i < n; i++) { unsigned b = i / threads, t = i % threads; REQUIRE(out[i] == in[b * threads + (threads - 1 - t)]); }
    cudaFree(in); cudaFree(out);
  }
  // ---- execution order: ascending by default, a permutation under GCSA_EMU_SHUFFLE ----
  {
    const unsigned blocks = 5, threads = 64, n = blocks * threads;
    unsigned *ticket = nullptr, *order = nullptr; cudaMalloc(&ticket, sizeof(unsigned)); cudaMalloc(&order, n * sizeof(unsigned));
    *ticket = 0;
    emu::launch(emu::Cfg(blocks, threads), [&] { order_kernel(ticket, order); }, false);
    REQUIRE(*ticket == n);
    std::vector<unsigned> seen(order, order + n);
    bool ascending = true; for(unsigned i = 0; i < n; i++) { ascending = ascending && (seen[i] == i); }
    REQUIRE(ascending == !emu::shuffling());
    std::sort(seen.begin(), seen.end());
    for(unsigned i = 0; i < n; i++) { REQUIRE(seen[i] == i); }                      // every thread ran exactly once
    cudaFree(ticket); cudaFree(order);
  }
  // ---- intrinsics ----
  REQUIRE(__fns(0b10110u, 0, 1) == 1 && __fns(0b10110u, 0, 3) == 4 && __fns(0b10110u, 0, 4) == 0xFFFFFFFFu);
  REQUIRE(__brev(1u) == 0x80000000u && __brev(0x0000F00Fu) == 0xF00F0000u && __brevll(1ull) == 0x8000000000000000ull);
  REQUIRE(__clzll(1ll) == 63 && __clzll(0ll) == 64 && __ffsll(0ll) == 0 && __ffsll(8ll) == 4 && __popcll(0xF0F0ull) == 8);
  REQUIRE(__vcmpltu4(0x01FF7F00u, 0x02FE7F01u) == 0xFF0000FFu);
  // ---- scan and segmented sort stand-ins (in place, as engine.cu calls them) ----
  {
    unsigned long long v[6] = { 3, 0, 4, 1, 5, 0 }; size_t bytes = 0;
    REQUIRE(cub::DeviceScan::ExclusiveSum(nullptr, bytes, v, v, 6) == cudaSuccess && bytes > 0);
    REQUIRE(cub::DeviceScan::ExclusiveSum((void*)v, bytes, v, v, 6) == cudaSuccess);
    unsigned long long expect[6] = { 0, 3, 3, 7, 8, 13 };
    for(int i = 0; i < 6; i++) { REQUIRE(v[i] == expect[i]); }
    unsigned long long keys[7] = { 9, 2, 5, 7, 1, 8, 3 }, sorted[7]; unsigned long long seg[4] = { 0, 3, 3, 7 };
    REQUIRE(cub::DeviceSegmentedSort::SortKeys((void*)keys, bytes, keys, sorted, 7ll, 3ll, seg, seg + 1) == cudaSuccess);
    unsigned long long want[7] = { 2, 5, 9, 1, 3, 7, 8 };
    for(int i = 0; i < 7; i++) { REQUIRE(sorted[i] == want[i]); }
  }
  std::printf("selfcheck OK%s\n", emu::shuffling() ? " (shuffled)" : "");
  return 0;
}
