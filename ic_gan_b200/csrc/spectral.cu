// Spectral normalisation (BigGAN_PyTorch/layers.py:39-61 power_iteration, :98-112 SN.W_) for all layers of a network
// in a handful of launches: one power-iteration step per forward (also in eval), sigma = u'^T W v, plus the kernels
// that turn the float32 OIHW master weight into the scaled operand copies the conv kernels read and that map the
// gradient of the scaled weight back onto the master weight (sigma is differentiated, layers.py:59).
#include "common.cuh"

namespace icgan {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void sn_zero_kernel(const IcganSnLayer* layers) {
  const IcganSnLayer L = layers[blockIdx.x];
  if (threadIdx.x < 2) L.scratch[threadIdx.x] = 0.f;
  for (int k = threadIdx.x; k < L.cols; k += blockDim.x) L.v[k] = 0.f;
}

constexpr int kSnRowChunk = 64;
// v_raw[k] += sum_{r in chunk} u[r] W[r][k].  Flat grid over all (layer, column block, row chunk) work items: block b
// finds its layer by scanning the per-layer item counts (<= a few dozen layers), so no block is launched idle.
__global__ void sn_wt_u_kernel(const IcganSnLayer* layers, int n_layers) {
  int item = blockIdx.x, li = 0, cbs = 0;
  for (; li < n_layers; ++li) {
    cbs = (layers[li].cols + 127) / 128;
    const int items = cbs * ((layers[li].rows + kSnRowChunk - 1) / kSnRowChunk);
    if (item < items) break;
    item -= items;
  }
  if (li >= n_layers) return;
  const IcganSnLayer L = layers[li];
  const int k = (item % cbs) * blockDim.x + threadIdx.x;
  const int r0 = (item / cbs) * kSnRowChunk;
  if (k >= L.cols || r0 >= L.rows) return;
  const int r1 = min(L.rows, r0 + kSnRowChunk);
  const float* w = L.W + k;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc = fmaf(L.u[r], w[static_cast<int64_t>(r) * L.cols], acc);
  atomicAdd(L.v + k, acc);
}

// scratch[0] = |v_raw|^2
__global__ void sn_vnorm_kernel(const IcganSnLayer* layers) {
  __shared__ float sh[8];
  const IcganSnLayer L = layers[blockIdx.x];
  float s = 0.f;
  for (int k = threadIdx.x; k < L.cols; k += blockDim.x) s = fmaf(L.v[k], L.v[k], s);
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) L.scratch[0] = t;
}

// t[r] = sum_k W[r][k] v_raw[k] / max(|v_raw|, eps) ; scratch[1] += t[r]^2     (one warp per row)
__global__ void sn_w_v_kernel(const IcganSnLayer* layers, float eps) {
  const IcganSnLayer L = layers[blockIdx.y];
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= L.rows) return;
  const int lane = threadIdx.x & 31;
  const float* w = L.W + static_cast<int64_t>(r) * L.cols;
  float acc = 0.f;
  for (int k = lane; k < L.cols; k += 32) acc = fmaf(w[k], L.v[k], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float t = acc / fmaxf(sqrtf(L.scratch[0]), eps);
    L.u_new[r] = t;
    atomicAdd(L.scratch + 1, t * t);
  }
}

__global__ void sn_finish_kernel(const IcganSnLayer* layers, float eps, int update_u) {
  const IcganSnLayer L = layers[blockIdx.x];
  const float nv = fmaxf(sqrtf(L.scratch[0]), eps);
  const float s1 = L.scratch[1];
  const float nt = fmaxf(sqrtf(s1), eps);
  for (int k = threadIdx.x; k < L.cols; k += blockDim.x) L.v[k] = L.v[k] / nv;
  for (int r = threadIdx.x; r < L.rows; r += blockDim.x) {
    const float un = L.u_new[r] / nt;
    L.u_new[r] = un;
    if (update_u) L.u[r] = un;
  }
  if (threadIdx.x == 0) {
    const float sigma = s1 / nt;  // (W v) . u'
    L.sigma[0] = sigma;
    L.sigma[1] = 1.f / sigma;
  }
}

// The three kernels below move weights between the OIHW float32 master ([Cout][Cin][k][k]) and the kernel layouts
// ([Cout][k][k][Cin] forward, [Cin][k'][k'][Cout] flipped for dgrad).  A block owns a 32 (co) x 32 (ci) x k*k tile and
// goes through shared memory, so that global reads and writes are contiguous runs on BOTH sides (the one-element-
// per-thread version wrote 2-byte elements Cin apart and ran at ~2 % of the HBM roofline).
static constexpr int kSnTile = 32;
static constexpr int kSnRow = kSnTile * 9 + 1;  // padded row: k*k <= 9

// OIHW float32 master -> scaled operand copies. fwd: [Cout][kh][kw][Cin]; dgrad: [Cin][k-1-kh][k-1-kw][Cout].
template <typename TO>
__global__ void __launch_bounds__(256)
sn_prepare_kernel(const float* __restrict__ W, const float* __restrict__ inv_sigma, TO* __restrict__ fwd,
                  TO* __restrict__ dgrad, int Cout, int Cin, int k) {
  __shared__ float sm[kSnTile][kSnRow];
  const int kk = k * k, ci0 = blockIdx.x * kSnTile, co0 = blockIdx.y * kSnTile;
  const int nci = min(kSnTile, Cin - ci0), nco = min(kSnTile, Cout - co0);
  const int rowlen = nci * kk;
  const float sc = inv_sigma ? *inv_sigma : 1.f;
  for (int idx = threadIdx.x; idx < nco * rowlen; idx += blockDim.x) {
    const int co_l = idx / rowlen, r = idx - co_l * rowlen;
    sm[co_l][r] = W[(static_cast<int64_t>(co0 + co_l) * Cin + ci0) * kk + r] * sc;
  }
  __syncthreads();
  if (fwd) {
    for (int idx = threadIdx.x; idx < nco * kk * kSnTile; idx += blockDim.x) {
      const int ci_l = idx % kSnTile, tap = (idx / kSnTile) % kk, co_l = idx / (kSnTile * kk);
      if (ci_l < nci)
        st_from_float(fwd, (static_cast<int64_t>(co0 + co_l) * kk + tap) * Cin + ci0 + ci_l, sm[co_l][ci_l * kk + tap]);
    }
  }
  if (dgrad) {
    for (int idx = threadIdx.x; idx < nci * kk * kSnTile; idx += blockDim.x) {
      const int co_l = idx % kSnTile, tap = (idx / kSnTile) % kk, ci_l = idx / (kSnTile * kk);
      if (co_l < nco)
        st_from_float(dgrad, (static_cast<int64_t>(ci0 + ci_l) * kk + (kk - 1 - tap)) * Cout + co0 + co_l,
                      sm[co_l][ci_l * kk + tap]);
    }
  }
}

// scratch[0] += <G, W> with G in kernel layout [Cout][k][k][Cin] and W in OIHW
__global__ void __launch_bounds__(256)
sn_grad_dot_kernel(const float* __restrict__ G, const float* __restrict__ W, float* scratch, int Cout, int Cin, int k) {
  __shared__ float sm[kSnTile][kSnRow];
  __shared__ float sh[8];
  const int kk = k * k, ci0 = blockIdx.x * kSnTile, co0 = blockIdx.y * kSnTile;
  const int nci = min(kSnTile, Cin - ci0), nco = min(kSnTile, Cout - co0);
  const int rowlen = nci * kk;
  for (int idx = threadIdx.x; idx < nco * rowlen; idx += blockDim.x) {
    const int co_l = idx / rowlen, r = idx - co_l * rowlen;
    sm[co_l][r] = W[(static_cast<int64_t>(co0 + co_l) * Cin + ci0) * kk + r];
  }
  __syncthreads();
  float s = 0.f;
  for (int idx = threadIdx.x; idx < nco * kk * kSnTile; idx += blockDim.x) {
    const int ci_l = idx % kSnTile, tap = (idx / kSnTile) % kk, co_l = idx / (kSnTile * kk);
    if (ci_l < nci)
      s = fmaf(G[(static_cast<int64_t>(co0 + co_l) * kk + tap) * Cin + ci0 + ci_l], sm[co_l][ci_l * kk + tap], s);
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(scratch, t);
}

// dW[i] = (G[map(i)] - (<G,W>/sigma) * u'[row] * v[col] / sigma ... ) see header; without SN: plain relayout
__global__ void __launch_bounds__(256)
sn_grad_apply_kernel(const float* __restrict__ G, const float* __restrict__ u_new, const float* __restrict__ v,
                     const float* __restrict__ sigma, const float* __restrict__ scratch, float* __restrict__ dW,
                     int Cout, int Cin, int k) {
  __shared__ float sm[kSnTile][kSnRow];
  const int kk = k * k, ci0 = blockIdx.x * kSnTile, co0 = blockIdx.y * kSnTile;
  const int nci = min(kSnTile, Cin - ci0), nco = min(kSnTile, Cout - co0);
  const int rowlen = nci * kk;
  const float inv = sigma ? sigma[1] : 1.f;
  const float coef = sigma ? scratch[0] * inv : 0.f;  // <G, W> / sigma = <G, W~>
  for (int idx = threadIdx.x; idx < nco * kk * kSnTile; idx += blockDim.x) {
    const int ci_l = idx % kSnTile, tap = (idx / kSnTile) % kk, co_l = idx / (kSnTile * kk);
    if (ci_l < nci) sm[co_l][ci_l * kk + tap] = G[(static_cast<int64_t>(co0 + co_l) * kk + tap) * Cin + ci0 + ci_l];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nco * rowlen; idx += blockDim.x) {
    const int co_l = idx / rowlen, r = idx - co_l * rowlen;
    float g = sm[co_l][r];
    if (sigma) g = (g - coef * u_new[co0 + co_l] * v[static_cast<int64_t>(ci0) * kk + r]) * inv;
    dW[(static_cast<int64_t>(co0 + co_l) * Cin + ci0) * kk + r] = g;
  }
}

}  // namespace icgan

using namespace icgan;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int icgan_sn_power_iteration(const IcganSnLayer* layers_dev, int n_layers, int max_rows, int max_cols,
                                        int64_t wt_u_items, float eps, int update_u, void* stream) {
  ICGAN_REQUIRE(layers_dev && n_layers > 0 && max_rows > 0 && max_cols > 0 && wt_u_items > 0,
                "icgan_sn_power_iteration: bad arguments");
  sn_zero_kernel<<<n_layers, 256, 0, STREAM>>>(layers_dev);
  sn_wt_u_kernel<<<static_cast<unsigned>(wt_u_items), 128, 0, STREAM>>>(layers_dev, n_layers);
  sn_vnorm_kernel<<<n_layers, 256, 0, STREAM>>>(layers_dev);
  sn_w_v_kernel<<<dim3((max_rows + 7) / 8, n_layers), 256, 0, STREAM>>>(layers_dev, eps);
  sn_finish_kernel<<<n_layers, 256, 0, STREAM>>>(layers_dev, eps, update_u);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_sn_prepare_weight(const float* W, const float* inv_sigma_dev, void* wk_fwd, void* wk_dgrad,
                                       int Cout, int Cin, int ksize, int out_dtype, void* stream) {
  ICGAN_REQUIRE(W && (wk_fwd || wk_dgrad), "icgan_sn_prepare_weight: null pointer");
  ICGAN_REQUIRE(ksize >= 1 && ksize <= 3, "icgan_sn_prepare_weight: ksize must be 1..3 (got %d)", ksize);
  const dim3 blocks(static_cast<unsigned>(ceil_div(Cin, kSnTile)), static_cast<unsigned>(ceil_div(Cout, kSnTile)));
  ICGAN_REQUIRE(blocks.y <= 65535u, "icgan_sn_prepare_weight: Cout too large");
  if (out_dtype == ICGAN_BF16)
    sn_prepare_kernel<__nv_bfloat16><<<blocks, 256, 0, STREAM>>>(W, inv_sigma_dev, static_cast<__nv_bfloat16*>(wk_fwd),
                                                                static_cast<__nv_bfloat16*>(wk_dgrad), Cout, Cin, ksize);
  else
    sn_prepare_kernel<float><<<blocks, 256, 0, STREAM>>>(W, inv_sigma_dev, static_cast<float*>(wk_fwd),
                                                        static_cast<float*>(wk_dgrad), Cout, Cin, ksize);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_sn_weight_grad(const float* G_k, const float* W, const float* u_new, const float* v,
                                    const float* sigma, float* scratch, float* dW, int Cout, int Cin, int ksize,
                                    void* stream) {
  ICGAN_REQUIRE(G_k && dW, "icgan_sn_weight_grad: null pointer");
  ICGAN_REQUIRE(ksize >= 1 && ksize <= 3, "icgan_sn_weight_grad: ksize must be 1..3 (got %d)", ksize);
  const dim3 blocks(static_cast<unsigned>(ceil_div(Cin, kSnTile)), static_cast<unsigned>(ceil_div(Cout, kSnTile)));
  ICGAN_REQUIRE(blocks.y <= 65535u, "icgan_sn_weight_grad: Cout too large");
  if (sigma) {
    ICGAN_REQUIRE(W && u_new && v && scratch, "icgan_sn_weight_grad: spectral-norm buffers missing");
    ICGAN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float), STREAM));
    sn_grad_dot_kernel<<<blocks, 256, 0, STREAM>>>(G_k, W, scratch, Cout, Cin, ksize);
  }
  sn_grad_apply_kernel<<<blocks, 256, 0, STREAM>>>(G_k, u_new, v, sigma, scratch, dW, Cout, Cin, ksize);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
