// Micro-benchmark: tcgen05.ld throughput per SM (cycles per 32x32b.x32 load, i.e. per 4 KB) with 4 or 8 reading warps,
// with the wait after every load or after every fourth.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ic_gan_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace icgan;

template <int WAIT_EVERY>
__global__ void __launch_bounds__(320, 1) k(int iters, int nwarps, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  long long t0 = 0, t1 = 0;
  uint32_t acc = 0;
  if (warp >= 2 && warp < 2 + nwarps) {
    const uint32_t taddr = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp - 2) >> 2) * 64;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld32(taddr + (i & 1) * 32, r);
      if ((i % WAIT_EVERY) == WAIT_EVERY - 1) tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
    tmem_ld_wait();
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + warp - 2] = t1 - t0;
    if (acc == 0x12345678u) out[1000] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

int main() {
  long long* out; cudaMallocManaged(&out, 4096 * sizeof(long long));
  const int iters = 4096;
  for (int nw : {1, 4, 8}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int i = 0; i < 4096; ++i) out[i] = 0;
      if (mode == 0) k<1><<<148, 320>>>(iters, nw, out); else k<4><<<148, 320>>>(iters, nw, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long mx = 0; for (int w = 0; w < nw; ++w) mx = out[w] > mx ? out[w] : mx;
      printf("warps %d wait every %d: %.1f cycles per x32 load per warp; SM total %.1f B/clk\n", nw, mode ? 4 : 1,
             (double)mx / iters, (double)nw * iters * 4096.0 / mx);
    }
  }
  return 0;
}
