// Generic strided, batched CUDA-core GEMM with float32 accumulation:
//   C[b][m][n] = alpha * sum_k A[b][m][k] * B[b][k][n] (+ bias[n]) (+ beta * C[b][m][n])
// Every operand is addressed through explicit (row, col, batch) element strides, so NN/NT/TN products and NHWC views
// need no copies.  Used for the small dense layers of the hot path -- SNLinear (layers.py:164-165: ccbn gain/bias
// embeddings of the class+instance conditioning vector, shared_feat 2048->512, G.linear, D.linear/linear_feat) and
// their backward -- and for the attention products theta^T phi and g beta^T (layers.py:237-243) and their backward.
#include "common.cuh"

namespace icgan {

struct GemmParams {
  int M, N, K, batch;
  int64_t sam, sak, sab;
  int64_t sbk, sbn, sbb;
  int64_t scm, scn, scb;
  float alpha, beta;
  const float* alpha_dev;
  const float* bias;
  int splits, k_per_split;  // split-K (batch == 1, float32 C zeroed by the host, atomic accumulation)
};

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ Bm, TC* __restrict__ C, GemmParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int b = p.splits > 1 ? 0 : blockIdx.z, split = p.splits > 1 ? blockIdx.z : 0;
  const int kbeg = split * p.k_per_split, kend = p.splits > 1 ? min(p.K, kbeg + p.k_per_split) : p.K;
  const TA* Ab = A + b * p.sab;
  const TB* Bb = Bm + b * p.sbb;
  TC* Cb = C + b * p.scb;
  const int lr = tid / 4, lk = (tid % 4) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      const int m = m0 + lr, n = n0 + lr;
      As[lk + j][lr] = (m < p.M && k < kend) ? ld_as_float(Ab, m * p.sam + k * p.sak) : 0.f;
      Bs[lk + j][lr] = (n < p.N && k < kend) ? ld_as_float(Bb, k * p.sbk + n * p.sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float alpha = p.alpha_dev ? p.alpha * (*p.alpha_dev) : p.alpha;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = alpha * acc[i][j];
      if (p.bias && split == 0) v += p.bias[n];
      const int64_t off = m * p.scm + n * p.scn;
      if (p.splits > 1) {
        atomicAdd(reinterpret_cast<float*>(Cb) + off, v);
        continue;
      }
      if (p.beta != 0.f) v = fmaf(p.beta, ld_as_float(Cb, off), v);
      st_from_float(Cb, off, v);
    }
  }
}

}  // namespace icgan

using namespace icgan;

extern "C" int icgan_gemm(const void* A, const void* B, void* C, int M, int N, int K, int batch, int64_t sam,
                          int64_t sak, int64_t sab, int64_t sbk, int64_t sbn, int64_t sbb, int64_t scm, int64_t scn,
                          int64_t scb, float alpha, const float* alpha_dev, float beta, const float* bias, int a_dtype,
                          int b_dtype, int c_dtype, void* stream) {
  ICGAN_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "icgan_gemm: bad arguments");
  ICGAN_REQUIRE(batch <= 65535, "icgan_gemm: batch too large");
  GemmParams p{M, N, K, batch, sam, sak, sab, sbk, sbn, sbb, scm, scn, scb, alpha, beta, alpha_dev, bias, 1, K};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tiles = ((N + 63) / 64) * ((M + 63) / 64);
  // skinny products (the [B, 657] x [C, 657]^T ccbn embeddings): too few tiles to fill 148 SMs -> split K
  if (batch == 1 && beta == 0.f && c_dtype == ICGAN_F32 && K >= 256 && tiles < num_sms() && scn == 1 &&
      scm == static_cast<int64_t>(N)) {
    int splits = (2 * num_sms() + tiles - 1) / tiles;
    if (splits > (K + 127) / 128) splits = (K + 127) / 128;
    if (splits > 32) splits = 32;
    if (splits > 1) {
      p.k_per_split = ((K + splits - 1) / splits + 15) / 16 * 16;
      p.splits = (K + p.k_per_split - 1) / p.k_per_split;
      ICGAN_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * static_cast<size_t>(M) * N, s));
    }
  }
  dim3 grid(static_cast<unsigned>((N + 63) / 64), static_cast<unsigned>((M + 63) / 64),
            static_cast<unsigned>(p.splits > 1 ? p.splits : batch));
  typedef __nv_bfloat16 bf;
  const int key = (a_dtype == ICGAN_BF16) * 4 + (b_dtype == ICGAN_BF16) * 2 + (c_dtype == ICGAN_BF16);
  switch (key) {
    case 0: gemm_simt_kernel<float, float, float><<<grid, 256, 0, s>>>((const float*)A, (const float*)B, (float*)C, p); break;
    case 1: gemm_simt_kernel<float, float, bf><<<grid, 256, 0, s>>>((const float*)A, (const float*)B, (bf*)C, p); break;
    case 2: gemm_simt_kernel<float, bf, float><<<grid, 256, 0, s>>>((const float*)A, (const bf*)B, (float*)C, p); break;
    case 3: gemm_simt_kernel<float, bf, bf><<<grid, 256, 0, s>>>((const float*)A, (const bf*)B, (bf*)C, p); break;
    case 4: gemm_simt_kernel<bf, float, float><<<grid, 256, 0, s>>>((const bf*)A, (const float*)B, (float*)C, p); break;
    case 5: gemm_simt_kernel<bf, float, bf><<<grid, 256, 0, s>>>((const bf*)A, (const float*)B, (bf*)C, p); break;
    case 6: gemm_simt_kernel<bf, bf, float><<<grid, 256, 0, s>>>((const bf*)A, (const bf*)B, (float*)C, p); break;
    default: gemm_simt_kernel<bf, bf, bf><<<grid, 256, 0, s>>>((const bf*)A, (const bf*)B, (bf*)C, p); break;
  }
  ICGAN_LAUNCH_CHECK();
  return 0;
}
