/* device/two_step.cuh -- construction of the two-step blocks (pairs of characters per probe) from the one-step blocks.
   Included by engine.cu only (construction kernels); sm_100a only. */
#ifndef GCSA2_B200_DEVICE_TWO_STEP_CUH
#define GCSA2_B200_DEVICE_TWO_STEP_CUH

//------------------------------------------------------------------------------
// Kernels: construction of the two-step blocks from the one-step blocks
//------------------------------------------------------------------------------


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

#endif
