// Microbenchmark: dependent chains of random 32-byte sector reads (LDG.E.256) over a footprint
// of S bytes.  Prints probes/s; run under ncu to read dram__bytes_read per probe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ ulonglong4 ld256(const ulonglong4* p)
{ ulonglong4 r; asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__global__ void init(ulonglong4* a, u64 n) { for(u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) a[i] = make_ulonglong4(mix(i), i, 0, 0); }
__global__ void __launch_bounds__(256) chase(const ulonglong4* a, u64 n_sectors, int steps, u64 stride_sectors, u64* out)
{
  u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  u64 x = mix(tid * 0x9E3779B97F4A7C15ULL + 1);
  u64 acc = 0;
  for(int s = 0; s < steps; s++)
  {
    u64 idx = ((x % n_sectors) / stride_sectors) * stride_sectors;
    ulonglong4 q = ld256(a + idx);
    acc += q.y; x = mix(q.x + x);
  }
  out[tid] = acc;
}
int main(int argc, char** argv)
{
  double mb = (argc > 1 ? atof(argv[1]) : 512); int steps = (argc > 2 ? atoi(argv[2]) : 20);
  u64 stride = (argc > 3 ? atoll(argv[3]) : 1); int blocks_per_sm = (argc > 4 ? atoi(argv[4]) : 6);
  u64 n = (u64)(mb * 1024 * 1024 / 32);
  ulonglong4* a; u64* out; u64 threads = 148ull * blocks_per_sm * 256 * 8;
  cudaMalloc(&a, n * 32); cudaMalloc(&out, threads * 8);
  init<<<148 * 8, 256>>>(a, n); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int rep = 0; rep < 3; rep++)
  {
    cudaEventRecord(e0); chase<<<(unsigned)(threads / 256), 256>>>(a, n, steps, stride, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if(rep == 2) printf("footprint %8.0f MB stride %llu sectors: %6.2f G probes/s  (%.3f ms, %llu threads x %d steps) %s\n", mb, stride, threads * (double)steps / ms / 1e6, ms, threads, steps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
