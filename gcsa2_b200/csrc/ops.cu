/*
  ops.cu -- LF(range, c), LF(node), LF_fast / LF_all and count(): device- and host-buffer entry points.
  One of the CUDA translation units of libgcsa2_b200.so (see engine.h); host side of the C ABI of include/gcsa2_b200.h,
  kernels in the device/*.cuh it includes.
*/
#include "engine.h"
#include "device/lf_count.cuh"
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


