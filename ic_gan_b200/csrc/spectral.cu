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

// OIHW float32 master -> scaled operand copies. fwd: [Cout][kh][kw][Cin]; dgrad: [Cin][k-1-kh][k-1-kw][Cout].
template <typename TO>
__global__ void sn_prepare_kernel(const float* __restrict__ W, const float* __restrict__ inv_sigma,
                                  TO* __restrict__ fwd, TO* __restrict__ dgrad, int Cout, int Cin, int k) {
  const int64_t total = static_cast<int64_t>(Cout) * Cin * k * k;
  const float sc = inv_sigma ? *inv_sigma : 1.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int kw = static_cast<int>(i % k);
    const int kh = static_cast<int>((i / k) % k);
    const int ci = static_cast<int>((i / (k * k)) % Cin);
    const int co = static_cast<int>(i / (static_cast<int64_t>(k) * k * Cin));
    const float v = W[i] * sc;
    if (fwd) st_from_float(fwd, ((static_cast<int64_t>(co) * k + kh) * k + kw) * Cin + ci, v);
    if (dgrad) st_from_float(dgrad, ((static_cast<int64_t>(ci) * k + (k - 1 - kh)) * k + (k - 1 - kw)) * Cout + co, v);
  }
}

// scratch[0] += <G, W> with G in kernel layout [Cout][k][k][Cin] and W in OIHW
__global__ void sn_grad_dot_kernel(const float* __restrict__ G, const float* __restrict__ W, float* scratch, int Cout,
                                   int Cin, int k) {
  __shared__ float sh[8];
  const int64_t total = static_cast<int64_t>(Cout) * Cin * k * k;
  float s = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int kw = static_cast<int>(i % k);
    const int kh = static_cast<int>((i / k) % k);
    const int ci = static_cast<int>((i / (k * k)) % Cin);
    const int64_t co = i / (static_cast<int64_t>(k) * k * Cin);
    s = fmaf(G[((co * k + kh) * k + kw) * Cin + ci], W[i], s);
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(scratch, t);
}

// dW[i] = (G[map(i)] - (<G,W>/sigma) * u'[row] * v[col] / sigma ... ) see header; without SN: plain relayout
__global__ void sn_grad_apply_kernel(const float* __restrict__ G, const float* __restrict__ u_new,
                                     const float* __restrict__ v, const float* __restrict__ sigma,
                                     const float* __restrict__ scratch, float* __restrict__ dW, int Cout, int Cin,
                                     int k) {
  const int64_t total = static_cast<int64_t>(Cout) * Cin * k * k;
  const int64_t cols = static_cast<int64_t>(Cin) * k * k;
  const float inv = sigma ? sigma[1] : 1.f;
  const float coef = sigma ? scratch[0] * inv : 0.f;  // <G, W> / sigma = <G, W~>
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int kw = static_cast<int>(i % k);
    const int kh = static_cast<int>((i / k) % k);
    const int ci = static_cast<int>((i / (k * k)) % Cin);
    const int64_t co = i / cols;
    float g = G[((co * k + kh) * k + kw) * Cin + ci];
    if (sigma) g = (g - coef * u_new[co] * v[i % cols]) * inv;
    dW[i] = g;
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
  const int64_t total = static_cast<int64_t>(Cout) * Cin * ksize * ksize;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
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
  const int64_t total = static_cast<int64_t>(Cout) * Cin * ksize * ksize;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  if (sigma) {
    ICGAN_REQUIRE(W && u_new && v && scratch, "icgan_sn_weight_grad: spectral-norm buffers missing");
    ICGAN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float), STREAM));
    sn_grad_dot_kernel<<<blocks, 256, 0, STREAM>>>(G_k, W, scratch, Cout, Cin, ksize);
  }
  sn_grad_apply_kernel<<<blocks, 256, 0, STREAM>>>(G_k, u_new, v, sigma, scratch, dW, Cout, Cin, ksize);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
