// HBM-bound NHWC kernels around the convolutions: batch-norm statistics / apply / backward with the per-sample
// (class+instance conditional) affine of layers.ccbn (BigGAN_PyTorch/layers.py:398-437) and layers.bn (:485-503),
// ReLU, nearest-upsample x2 (BigGAN.py:260), 2x2 avg/sum/max pooling (BigGAN.py:528, layers.py:230-231),
// D's global sum pooling (BigGAN.py:624), tanh backward and small elementwise helpers.
// All index arithmetic is on the flat [pixels][C] view, threads run along C => fully coalesced.
#include "common.cuh"
#include "norm_act_vec.cuh"

namespace icgan {

#define DISPATCH_T(dt, T, ...)                     \
  if ((dt) == ICGAN_BF16) {                        \
    using T = __nv_bfloat16;                       \
    __VA_ARGS__                                    \
  } else {                                         \
    using T = float;                               \
    __VA_ARGS__                                    \
  }

static inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---------------------------------------------------------------- statistics
// pass 1: ws[c] += sum_p x[p][c];  pass 2: ws[C + c] += sum_p (x[p][c] - mean_c)^2  (two-pass => no cancellation)
template <typename T, int PASS>
__global__ void bn_stats_kernel(const T* __restrict__ x, float* __restrict__ ws, int64_t P, int C,
                                int64_t rows_per_block) {
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < P ? r0 + rows_per_block : P;
  const float invP = 1.0f / static_cast<float>(P);
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < C; c += gridDim.y * blockDim.x) {
    float s = 0.f;
    int64_t r = r0;
    if (PASS == 1) {
      for (; r + 4 <= r1; r += 4) {
        const float a = ld_as_float(x, r * C + c), b = ld_as_float(x, (r + 1) * C + c),
                    d = ld_as_float(x, (r + 2) * C + c), e = ld_as_float(x, (r + 3) * C + c);
        s += (a + b) + (d + e);
      }
      for (; r < r1; ++r) s += ld_as_float(x, r * C + c);
      atomicAdd(ws + c, s);
    } else {
      const float m = ws[c] * invP;
      for (; r + 4 <= r1; r += 4) {
        const float a = ld_as_float(x, r * C + c) - m, b = ld_as_float(x, (r + 1) * C + c) - m,
                    d = ld_as_float(x, (r + 2) * C + c) - m, e = ld_as_float(x, (r + 3) * C + c) - m;
        s += (a * a + b * b) + (d * d + e * e);
      }
      for (; r < r1; ++r) {
        const float d = ld_as_float(x, r * C + c) - m;
        s = fmaf(d, d, s);
      }
      atomicAdd(ws + C + c, s);
    }
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ ws, float* running_mean, float* running_var,
                                   float* __restrict__ mean, float* __restrict__ invstd, int64_t P, int C, float eps,
                                   float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = ws[c] / static_cast<float>(P);
  const float var = ws[C + c] / static_cast<float>(P);  // biased, used to normalise
  mean[c] = m;
  invstd[c] = rsqrtf(var + eps);
  if (running_mean) {
    const float unbiased = P > 1 ? var * (static_cast<float>(P) / static_cast<float>(P - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

__global__ void bn_from_sums_kernel(const float* __restrict__ sums, const float* __restrict__ shift, float* running_mean,
                                    float* running_var, float* __restrict__ mean, float* __restrict__ invstd, int64_t P,
                                    int C, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invP = 1.f / static_cast<float>(P);
  const float m0 = sums[c] * invP;                         // mean of (y - shift)
  const float var = fmaxf(sums[C + c] * invP - m0 * m0, 0.f);  // biased variance (shift-invariant)
  const float m = m0 + (shift ? shift[c] : 0.f);
  mean[c] = m;
  invstd[c] = rsqrtf(var + eps);
  if (running_mean) {
    const float unbiased = P > 1 ? var * (static_cast<float>(P) / static_cast<float>(P - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// ---------------------------------------------------------------- apply:  y = act(xhat * gain[n,c] + bias[n,c]) (+up2)
template <typename TI, typename TO>
__global__ void bn_apply_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gain,
                                const float* __restrict__ bias, int gstride, int B, int H, int W, int C, int relu,
                                int up) {
  const int64_t total = static_cast<int64_t>(B) * H * W * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t pix = i / C;
    const int n = static_cast<int>(pix / (static_cast<int64_t>(H) * W));
    const float xh = (ld_as_float(x, i) - mean[c]) * invstd[c];
    float v = xh * gain[static_cast<int64_t>(n) * gstride + c] + bias[static_cast<int64_t>(n) * gstride + c];
    if (relu) v = fmaxf(v, 0.f);
    if (!up) {
      st_from_float(y, i, v);
    } else {
      const int w = static_cast<int>(pix % W);
      const int h = static_cast<int>((pix / W) % H);
      const int64_t o = ((static_cast<int64_t>(n) * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
      const int64_t rs = static_cast<int64_t>(2) * W * C;
      st_from_float(y, o, v);
      st_from_float(y, o + C, v);
      st_from_float(y, o + rs, v);
      st_from_float(y, o + rs + C, v);
    }
  }
}

// gradient reaching the BN-affine output at input resolution (sums the 2x2 children when the forward upsampled,
// and applies the ReLU mask recomputed from x)
template <typename TX, typename TG>
__device__ __forceinline__ float bn_out_grad(const TX* x, const TG* dy, const float* mean, const float* invstd,
                                             const float* gain, const float* bias, int gstride, int n, int h, int w,
                                             int c, int H, int W, int C, int relu, int up, float& xhat) {
  const int64_t i = ((static_cast<int64_t>(n) * H + h) * W + w) * C + c;
  xhat = (ld_as_float(x, i) - mean[c]) * invstd[c];
  const int64_t gi = static_cast<int64_t>(n) * gstride + c;
  if (relu && !(xhat * gain[gi] + bias[gi] > 0.f)) return 0.f;
  if (!up) return ld_as_float(dy, i);
  const int64_t o = ((static_cast<int64_t>(n) * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
  const int64_t rs = static_cast<int64_t>(2) * W * C;
  return ld_as_float(dy, o) + ld_as_float(dy, o + C) + ld_as_float(dy, o + rs) + ld_as_float(dy, o + rs + C);
}

// s1[n,c] += sum_hw g ; s2[n,c] += sum_hw g * xhat      (grid: x = pixel slabs, y = n)
template <typename TX, typename TG>
__global__ void bn_bwd_reduce_kernel(const TX* __restrict__ x, const TG* __restrict__ dy,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ gain, const float* __restrict__ bias, int gstride,
                                     float* __restrict__ s1, float* __restrict__ s2, int H, int W, int C, int relu,
                                     int up, int pix_per_block) {
  const int n = blockIdx.y;
  const int HW = H * W;
  const int q0 = blockIdx.x * pix_per_block;
  const int q1 = min(HW, q0 + pix_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a1 = 0.f, a2 = 0.f;
    int q = q0;
    for (; q + 4 <= q1; q += 4) {  // 4 independent pixels in flight per thread
      float xh[4], g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        g[j] = bn_out_grad(x, dy, mean, invstd, gain, bias, gstride, n, (q + j) / W, (q + j) % W, c, H, W, C, relu, up,
                           xh[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a1 += g[j];
        a2 = fmaf(g[j], xh[j], a2);
      }
    }
    for (; q < q1; ++q) {
      float xh;
      const float g = bn_out_grad(x, dy, mean, invstd, gain, bias, gstride, n, q / W, q % W, c, H, W, C, relu, up, xh);
      a1 += g;
      a2 = fmaf(g, xh, a2);
    }
    atomicAdd(s1 + static_cast<int64_t>(n) * C + c, a1);
    atomicAdd(s2 + static_cast<int64_t>(n) * C + c, a2);
  }
}

// dx = invstd * (gain * g - m1 - xhat * m2)
template <typename TX, typename TG, typename TO>
__global__ void bn_bwd_apply_kernel(const TX* __restrict__ x, const TG* __restrict__ dy, TO* __restrict__ dx,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gain, const float* __restrict__ bias, int gstride,
                                    const float* __restrict__ m1, const float* __restrict__ m2, int B, int H, int W,
                                    int C, int relu, int up) {
  const int64_t total = static_cast<int64_t>(B) * H * W * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t pix = i / C;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<int64_t>(H) * W));
    float xh;
    const float g = bn_out_grad(x, dy, mean, invstd, gain, bias, gstride, n, h, w, c, H, W, C, relu, up, xh);
    const float v = invstd[c] * (gain[static_cast<int64_t>(n) * gstride + c] * g - m1[c] - xh * m2[c]);
    st_from_float(dx, i, v);
  }
}

// ---------------------------------------------------------------- elementwise
template <typename T>
__global__ void relu_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    st_from_float(y, i, fmaxf(ld_as_float(x, i), 0.f));
}
// dx = dy * (ref > 0)   (ref = forward input or output of the ReLU; identical mask)
template <typename T, typename TG>
__global__ void relu_bwd_kernel(const TG* __restrict__ dy, const T* __restrict__ ref, TG* __restrict__ dx, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    st_from_float(dx, i, ld_as_float(ref, i) > 0.f ? ld_as_float(dy, i) : 0.f);
}
// dx = dy * (1 - y^2)
template <typename T, typename TG>
__global__ void tanh_bwd_kernel(const TG* __restrict__ dy, const T* __restrict__ y, TG* __restrict__ dx, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float t = ld_as_float(y, i);
    st_from_float(dx, i, ld_as_float(dy, i) * (1.f - t * t));
  }
}
// out = alpha * a + beta * b   (alpha/beta read from device scalars when the pointers are non-null)
template <typename T>
__global__ void axpby_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, float alpha,
                             const float* alpha_p, float beta, const float* beta_p, int64_t n) {
  const float al = alpha_p ? *alpha_p : alpha, be = beta_p ? *beta_p : beta;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = al * ld_as_float(a, i);
    if (b) v = fmaf(be, ld_as_float(b, i), v);
    st_from_float(out, i, v);
  }
}
// out[0] += sum a*b
template <typename T>
__global__ void dot_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, int64_t n) {
  // double accumulation inside the block: the result (attention's d gamma = <dy, o>) is a heavily cancelling sum
  // (measured at real width: |sum| ~ 1e-4 of sum|.|), float partials lose it; HBM-bound either way
  double s = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s = fma(static_cast<double>(ld_as_float(a, i)), static_cast<double>(ld_as_float(b, i)), s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += part[i];
    atomicAdd(out, static_cast<float>(t));
  }
}

// ---------------------------------------------------------------- 2x2 pooling / upsampling
// mode 0: y = scale * sum_{2x2} x (+ add) ; mode 1: y = max_{2x2} x
template <typename T>
__global__ void pool2_kernel(const T* __restrict__ x, const T* __restrict__ add, T* __restrict__ y, int B, int Ho,
                             int Wo, int C, float scale, int mode) {
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t pix = i / C;
    const int w = static_cast<int>(pix % Wo);
    const int h = static_cast<int>((pix / Wo) % Ho);
    const int64_t n = pix / (static_cast<int64_t>(Ho) * Wo);
    const int64_t o = ((n * 2 * Ho + 2 * h) * 2 * Wo + 2 * w) * C + c;
    const int64_t rs = static_cast<int64_t>(2) * Wo * C;
    const float a = ld_as_float(x, o), b = ld_as_float(x, o + C), d = ld_as_float(x, o + rs),
                e = ld_as_float(x, o + rs + C);
    float v;
    if (mode == 0) {
      v = (a + b + d + e) * scale;
      if (add) v += ld_as_float(add, i);
    } else {
      v = fmaxf(fmaxf(a, b), fmaxf(d, e));
    }
    st_from_float(y, i, v);
  }
}
// mode 0: dx[2h+a,2w+b] = scale * dy[h,w] (nearest upsample; avg/sum-pool backward)
// mode 1: max-pool backward: the gradient goes to the first maximal element in (a,b) scan order (ATen's rule)
template <typename T, typename TG>
__global__ void unpool2_kernel(const TG* __restrict__ dy, const T* __restrict__ xref, TG* __restrict__ dx, int B,
                               int Ho, int Wo, int C, float scale, int mode) {
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t pix = i / C;
    const int w = static_cast<int>(pix % Wo);
    const int h = static_cast<int>((pix / Wo) % Ho);
    const int64_t n = pix / (static_cast<int64_t>(Ho) * Wo);
    const int64_t o = ((n * 2 * Ho + 2 * h) * 2 * Wo + 2 * w) * C + c;
    const int64_t rs = static_cast<int64_t>(2) * Wo * C;
    const float g = ld_as_float(dy, i) * scale;
    if (mode == 0) {
      st_from_float(dx, o, g);
      st_from_float(dx, o + C, g);
      st_from_float(dx, o + rs, g);
      st_from_float(dx, o + rs + C, g);
    } else {
      const float v[4] = {ld_as_float(xref, o), ld_as_float(xref, o + C), ld_as_float(xref, o + rs),
                          ld_as_float(xref, o + rs + C)};
      int best = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k] > v[best]) best = k;
      st_from_float(dx, o, best == 0 ? g : 0.f);
      st_from_float(dx, o + C, best == 1 ? g : 0.f);
      st_from_float(dx, o + rs, best == 2 ? g : 0.f);
      st_from_float(dx, o + rs + C, best == 3 ? g : 0.f);
    }
  }
}

// out[n,c] += sum_hw relu(x[n,hw,c])
template <typename T>
__global__ void relu_sumpool_kernel(const T* __restrict__ x, float* __restrict__ out, int HW, int C,
                                    int pix_per_block) {
  const int n = blockIdx.y;
  const int q0 = blockIdx.x * pix_per_block, q1 = min(HW, q0 + pix_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int q = q0; q < q1; ++q) s += fmaxf(ld_as_float(x, (static_cast<int64_t>(n) * HW + q) * C + c), 0.f);
    atomicAdd(out + static_cast<int64_t>(n) * C + c, s);
  }
}
// dx[n,hw,c] = dh[n,c] * (x > 0)
template <typename T>
__global__ void relu_sumpool_bwd_kernel(const T* __restrict__ x, const float* __restrict__ dh, T* __restrict__ dx,
                                        int B, int HW, int C) {
  const int64_t total = static_cast<int64_t>(B) * HW * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t n = i / (static_cast<int64_t>(HW) * C);
    st_from_float(dx, i, ld_as_float(x, i) > 0.f ? dh[n * C + c] : 0.f);
  }
}

// ---------------------------------------------------------------- softmax over rows (attention maps, layers.py:237)
template <typename TS, typename TP>
__global__ void softmax_rows_kernel(const TS* __restrict__ s, TP* __restrict__ p, int64_t rows, int cols) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TS* sr = s + row * cols;
  float mx = -INFINITY;
  for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, ld_as_float(sr, j));
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < cols; j += 32) sum += __expf(ld_as_float(sr, j) - mx);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  TP* pr = p + row * cols;
  for (int j = lane; j < cols; j += 32) st_from_float(pr, j, __expf(ld_as_float(sr, j) - mx) * inv);
}
// ds = p * (dp - sum_j dp_j p_j)
template <typename TP, typename TD, typename TO>
__global__ void softmax_rows_bwd_kernel(const TP* __restrict__ p, const TD* __restrict__ dp, TO* __restrict__ ds,
                                        int64_t rows, int cols) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TP* pr = p + row * cols;
  const TD* dr = dp + row * cols;
  float dot = 0.f;
  for (int j = lane; j < cols; j += 32) dot = fmaf(ld_as_float(pr, j), ld_as_float(dr, j), dot);
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  TO* or_ = ds + row * cols;
  for (int j = lane; j < cols; j += 32) st_from_float(or_, j, ld_as_float(pr, j) * (ld_as_float(dr, j) - dot));
}

}  // namespace icgan

using namespace icgan;
#define STREAM static_cast<cudaStream_t>(stream)

static int reduce_threads(int C) { return C >= 256 ? 256 : (C >= 128 ? 128 : (C >= 64 ? 64 : 32)); }

extern "C" int icgan_bn_train_stats(const void* x, int64_t P, int C, int dtype, float* ws, const float* shift,
                                    float* running_mean, float* running_var, float* mean, float* invstd, float eps,
                                    float momentum, void* stream) {
  ICGAN_REQUIRE(x && ws && mean && invstd && P > 0 && C > 0, "icgan_bn_train_stats: bad arguments");
  ICGAN_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, STREAM));
  if (dtype == ICGAN_BF16 && vec::ok(C)) {
    int64_t blocks = static_cast<int64_t>(num_sms()) * 8;
    if (blocks > (P + 127) / 128) blocks = (P + 127) / 128;
    const int64_t ppb = (P + blocks - 1) / blocks;
    blocks = (P + ppb - 1) / ppb;
    const vec::bf16* xb = static_cast<const vec::bf16*>(x);
    // one pass over the activation: shifted first and second moments (shift = last step's batch mean, or NULL)
    vec::shifted_moments_vec_kernel<<<static_cast<unsigned>(blocks), vec::kThreads, 0, STREAM>>>(xb, shift, ws, P, C, ppb);
    bn_from_sums_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(ws, shift, running_mean, running_var, mean, invstd, P, C,
                                                            eps, momentum);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  int row_blocks = static_cast<int>((P + 63) / 64);
  if (row_blocks > 24 * num_sms()) row_blocks = 24 * num_sms();
  const int64_t rpb = (P + row_blocks - 1) / row_blocks;
  row_blocks = static_cast<int>((P + rpb - 1) / rpb);
  const int threads = reduce_threads(C);
  dim3 grid(static_cast<unsigned>(row_blocks), static_cast<unsigned>((C + threads - 1) / threads));
  DISPATCH_T(dtype, T, {
    bn_stats_kernel<T, 1><<<grid, threads, 0, STREAM>>>(static_cast<const T*>(x), ws, P, C, rpb);
    bn_stats_kernel<T, 2><<<grid, threads, 0, STREAM>>>(static_cast<const T*>(x), ws, P, C, rpb);
  })
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(ws, running_mean, running_var, mean, invstd, P, C, eps,
                                                          momentum);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_bn_stats_from_sums(const float* sums, const float* shift, int64_t P, int C, float* running_mean,
                                        float* running_var, float* mean, float* invstd, float eps, float momentum,
                                        void* stream) {
  ICGAN_REQUIRE(sums && mean && invstd && P > 0 && C > 0, "icgan_bn_stats_from_sums: bad arguments");
  bn_from_sums_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(sums, shift, running_mean, running_var, mean, invstd, P, C,
                                                          eps, momentum);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_bn_apply(const void* x, void* y, const float* mean, const float* invstd, const float* gain,
                              const float* bias, int gain_stride, int B, int H, int W, int C, int relu, int up,
                              int in_dtype, int out_dtype, void* stream) {
  ICGAN_REQUIRE(x && y && mean && invstd && gain && bias, "icgan_bn_apply: null pointer");
  const int64_t total = static_cast<int64_t>(B) * H * W * C;
  if (in_dtype == ICGAN_BF16 && out_dtype == ICGAN_BF16 && vec::ok(C)) {
    int slabs;
    const int ppb = vec::slab_grid(B, H * W, &slabs);
    vec::bn_apply_vec_kernel<<<dim3(static_cast<unsigned>(slabs), static_cast<unsigned>(B)), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(x), static_cast<vec::bf16*>(y), mean, invstd, gain, bias, gain_stride, H, W, C,
        relu, up, ppb);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  const int blocks = ew_blocks(total);
  DISPATCH_T(in_dtype, TI, {DISPATCH_T(out_dtype, TO, {
    bn_apply_kernel<TI, TO><<<blocks, 256, 0, STREAM>>>(static_cast<const TI*>(x), static_cast<TO*>(y), mean, invstd,
                                                       gain, bias, gain_stride, B, H, W, C, relu, up);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_bn_bwd_reduce(const void* x, const void* dy, const float* mean, const float* invstd,
                                   const float* gain, const float* bias, int gain_stride, float* s1, float* s2, int B,
                                   int H, int W, int C, int relu, int up, int x_dtype, int dy_dtype, void* stream) {
  ICGAN_REQUIRE(x && dy && s1 && s2, "icgan_bn_bwd_reduce: null pointer");
  ICGAN_CUDA(cudaMemsetAsync(s1, 0, sizeof(float) * B * C, STREAM));
  ICGAN_CUDA(cudaMemsetAsync(s2, 0, sizeof(float) * B * C, STREAM));
  const int HW = H * W;
  if (x_dtype == ICGAN_BF16 && dy_dtype == ICGAN_BF16 && vec::ok(C)) {
    int vs = ceil_div(8 * num_sms(), B);
    if (vs > (HW + 31) / 32) vs = (HW + 31) / 32;
    if (vs < 1) vs = 1;
    const int vppb = (HW + vs - 1) / vs;
    vs = (HW + vppb - 1) / vppb;
    const dim3 vgrid(static_cast<unsigned>(vs), static_cast<unsigned>(B));
    if (up)
      vec::bn_bwd_reduce_vec_kernel<false><<<vgrid, vec::kThreads, 0, STREAM>>>(
          static_cast<const vec::bf16*>(x), static_cast<const vec::bf16*>(dy), mean, invstd, gain, bias, gain_stride, s1,
          s2, H, W, C, relu, up, vppb);
    else
      vec::bn_bwd_reduce_vec_kernel<true><<<vgrid, vec::kThreads, 0, STREAM>>>(
          static_cast<const vec::bf16*>(x), static_cast<const vec::bf16*>(dy), mean, invstd, gain, bias, gain_stride, s1,
          s2, H, W, C, relu, up, vppb);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  int slabs = ceil_div(24 * num_sms(), B);
  if (slabs > (HW + 15) / 16) slabs = (HW + 15) / 16;
  if (slabs < 1) slabs = 1;
  const int ppb = (HW + slabs - 1) / slabs;
  slabs = (HW + ppb - 1) / ppb;
  dim3 grid(static_cast<unsigned>(slabs), static_cast<unsigned>(B));
  const int threads = reduce_threads(C);
  DISPATCH_T(x_dtype, TX, {DISPATCH_T(dy_dtype, TG, {
    bn_bwd_reduce_kernel<TX, TG><<<grid, threads, 0, STREAM>>>(static_cast<const TX*>(x), static_cast<const TG*>(dy),
                                                              mean, invstd, gain, bias, gain_stride, s1, s2, H, W, C,
                                                              relu, up, ppb);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_bn_bwd_apply(const void* x, const void* dy, void* dx, const float* mean, const float* invstd,
                                  const float* gain, const float* bias, int gain_stride, const float* m1,
                                  const float* m2, int B, int H, int W, int C, int relu, int up, int x_dtype,
                                  int dy_dtype, void* stream) {
  ICGAN_REQUIRE(x && dy && dx && m1 && m2, "icgan_bn_bwd_apply: null pointer");
  const int64_t total = static_cast<int64_t>(B) * H * W * C;
  if (x_dtype == ICGAN_BF16 && dy_dtype == ICGAN_BF16 && vec::ok(C)) {
    int slabs;
    const int ppb = vec::slab_grid(B, H * W, &slabs);
    const dim3 vgrid(static_cast<unsigned>(slabs), static_cast<unsigned>(B));
    if (up)
      vec::bn_bwd_apply_vec_kernel<false><<<vgrid, vec::kThreads, 0, STREAM>>>(
          static_cast<const vec::bf16*>(x), static_cast<const vec::bf16*>(dy), static_cast<vec::bf16*>(dx), mean, invstd,
          gain, bias, gain_stride, m1, m2, H, W, C, relu, up, ppb);
    else
      vec::bn_bwd_apply_vec_kernel<true><<<vgrid, vec::kThreads, 0, STREAM>>>(
          static_cast<const vec::bf16*>(x), static_cast<const vec::bf16*>(dy), static_cast<vec::bf16*>(dx), mean, invstd,
          gain, bias, gain_stride, m1, m2, H, W, C, relu, up, ppb);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  const int blocks = ew_blocks(total);
  DISPATCH_T(x_dtype, TX, {DISPATCH_T(dy_dtype, TG, {
    bn_bwd_apply_kernel<TX, TG, TG><<<blocks, 256, 0, STREAM>>>(static_cast<const TX*>(x), static_cast<const TG*>(dy),
                                                               static_cast<TG*>(dx), mean, invstd, gain, bias,
                                                               gain_stride, m1, m2, B, H, W, C, relu, up);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_relu(const void* x, void* y, int64_t n, int dtype, void* stream) {
  ICGAN_REQUIRE(x && y && n > 0, "icgan_relu: bad arguments");
  if (dtype == ICGAN_BF16 && n % 8 == 0) {
    vec::ew_vec_kernel<0><<<vec::blocks_for(n / 8), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(x), nullptr, static_cast<vec::bf16*>(y), 0.f, 0.f, nullptr, nullptr, n / 8);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(dtype, T, { relu_kernel<T><<<ew_blocks(n), 256, 0, STREAM>>>(static_cast<const T*>(x), static_cast<T*>(y), n); })
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_relu_bwd(const void* dy, const void* ref, void* dx, int64_t n, int ref_dtype, int g_dtype,
                              void* stream) {
  ICGAN_REQUIRE(dy && ref && dx && n > 0, "icgan_relu_bwd: bad arguments");
  if (ref_dtype == ICGAN_BF16 && g_dtype == ICGAN_BF16 && n % 8 == 0) {
    vec::ew_vec_kernel<1><<<vec::blocks_for(n / 8), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(dy), static_cast<const vec::bf16*>(ref), static_cast<vec::bf16*>(dx), 0.f, 0.f,
        nullptr, nullptr, n / 8);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(ref_dtype, T, {DISPATCH_T(g_dtype, TG, {
    relu_bwd_kernel<T, TG><<<ew_blocks(n), 256, 0, STREAM>>>(static_cast<const TG*>(dy), static_cast<const T*>(ref),
                                                            static_cast<TG*>(dx), n);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, int y_dtype, int g_dtype,
                              void* stream) {
  ICGAN_REQUIRE(dy && y && dx && n > 0, "icgan_tanh_bwd: bad arguments");
  DISPATCH_T(y_dtype, T, {DISPATCH_T(g_dtype, TG, {
    tanh_bwd_kernel<T, TG><<<ew_blocks(n), 256, 0, STREAM>>>(static_cast<const TG*>(dy), static_cast<const T*>(y),
                                                            static_cast<TG*>(dx), n);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_axpby(const void* a, const void* b, void* out, float alpha, const float* alpha_dev, float beta,
                           const float* beta_dev, int64_t n, int dtype, void* stream) {
  ICGAN_REQUIRE(a && out && n > 0, "icgan_axpby: bad arguments");
  if (dtype == ICGAN_BF16 && n % 8 == 0) {
    vec::ew_vec_kernel<3><<<vec::blocks_for(n / 8), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(a), static_cast<const vec::bf16*>(b), static_cast<vec::bf16*>(out), alpha, beta,
        alpha_dev, beta_dev, n / 8);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(dtype, T, {
    axpby_kernel<T><<<ew_blocks(n), 256, 0, STREAM>>>(static_cast<const T*>(a), static_cast<const T*>(b),
                                                     static_cast<T*>(out), alpha, alpha_dev, beta, beta_dev, n);
  })
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_dot(const void* a, const void* b, float* out, int64_t n, int dtype, void* stream) {
  ICGAN_REQUIRE(a && b && out && n > 0, "icgan_dot: bad arguments");
  ICGAN_CUDA(cudaMemsetAsync(out, 0, sizeof(float), STREAM));
  int blocks = ew_blocks(n);
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  DISPATCH_T(dtype, T, { dot_kernel<T><<<blocks, 256, 0, STREAM>>>(static_cast<const T*>(a), static_cast<const T*>(b), out, n); })
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_pool2(const void* x, const void* add, void* y, int B, int Hout, int Wout, int C, float scale,
                           int mode, int dtype, void* stream) {
  ICGAN_REQUIRE(x && y, "icgan_pool2: null pointer");
  const int64_t total = static_cast<int64_t>(B) * Hout * Wout * C;
  if (dtype == ICGAN_BF16 && vec::ok(C)) {
    vec::pool2_vec_kernel<<<vec::blocks_for(total / 8), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(x), static_cast<const vec::bf16*>(add), static_cast<vec::bf16*>(y), B, Hout, Wout,
        C, scale, mode);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(dtype, T, {
    pool2_kernel<T><<<ew_blocks(total), 256, 0, STREAM>>>(static_cast<const T*>(x), static_cast<const T*>(add),
                                                         static_cast<T*>(y), B, Hout, Wout, C, scale, mode);
  })
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_unpool2(const void* dy, const void* xref, void* dx, int B, int Hout, int Wout, int C, float scale,
                             int mode, int ref_dtype, int g_dtype, void* stream) {
  ICGAN_REQUIRE(dy && dx && (mode == 0 || xref), "icgan_unpool2: null pointer");
  const int64_t total = static_cast<int64_t>(B) * Hout * Wout * C;
  if (g_dtype == ICGAN_BF16 && (mode == 0 || ref_dtype == ICGAN_BF16) && vec::ok(C)) {
    vec::unpool2_vec_kernel<<<vec::blocks_for(total / 8), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(dy), static_cast<const vec::bf16*>(xref), static_cast<vec::bf16*>(dx), B, Hout,
        Wout, C, scale, mode);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(ref_dtype, T, {DISPATCH_T(g_dtype, TG, {
    unpool2_kernel<T, TG><<<ew_blocks(total), 256, 0, STREAM>>>(static_cast<const TG*>(dy),
                                                               static_cast<const T*>(xref), static_cast<TG*>(dx), B,
                                                               Hout, Wout, C, scale, mode);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_relu_sumpool(const void* x, float* out, int B, int HW, int C, int dtype, void* stream) {
  ICGAN_REQUIRE(x && out, "icgan_relu_sumpool: null pointer");
  ICGAN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * B * C, STREAM));
  if (dtype == ICGAN_BF16 && vec::ok(C)) {
    int vs = ceil_div(4 * num_sms(), B);
    if (vs > (HW + 15) / 16) vs = (HW + 15) / 16;
    if (vs < 1) vs = 1;
    const int vppb = (HW + vs - 1) / vs;
    vs = (HW + vppb - 1) / vppb;
    vec::relu_sumpool_vec_kernel<<<dim3(static_cast<unsigned>(vs), static_cast<unsigned>(B)), vec::kThreads, 0, STREAM>>>(
        static_cast<const vec::bf16*>(x), out, HW, C, vppb);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  int slabs = ceil_div(2 * num_sms(), B);
  if (slabs > HW) slabs = HW;
  if (slabs < 1) slabs = 1;
  const int ppb = (HW + slabs - 1) / slabs;
  slabs = (HW + ppb - 1) / ppb;
  dim3 grid(static_cast<unsigned>(slabs), static_cast<unsigned>(B));
  DISPATCH_T(dtype, T, {
    relu_sumpool_kernel<T><<<grid, reduce_threads(C), 0, STREAM>>>(static_cast<const T*>(x), out, HW, C, ppb);
  })
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_relu_sumpool_bwd(const void* x, const float* dh, void* dx, int B, int HW, int C, int dtype,
                                      void* stream) {
  ICGAN_REQUIRE(x && dh && dx, "icgan_relu_sumpool_bwd: null pointer");
  const int64_t total = static_cast<int64_t>(B) * HW * C;
  DISPATCH_T(dtype, T, {
    relu_sumpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, STREAM>>>(static_cast<const T*>(x), dh, static_cast<T*>(dx),
                                                                    B, HW, C);
  })
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_softmax_rows(const void* s, void* p, int64_t rows, int cols, int s_dtype, int p_dtype,
                                  void* stream) {
  ICGAN_REQUIRE(s && p && rows > 0 && cols > 0, "icgan_softmax_rows: bad arguments");
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  DISPATCH_T(s_dtype, TS, {DISPATCH_T(p_dtype, TP, {
    softmax_rows_kernel<TS, TP><<<blocks, 256, 0, STREAM>>>(static_cast<const TS*>(s), static_cast<TP*>(p), rows, cols);
  })})
  ICGAN_LAUNCH_CHECK();
  return 0;
}
extern "C" int icgan_softmax_rows_bwd(const void* p, const void* dp, void* ds, int64_t rows, int cols, int p_dtype,
                                      int dp_dtype, int ds_dtype, void* stream) {
  ICGAN_REQUIRE(p && dp && ds && rows > 0 && cols > 0, "icgan_softmax_rows_bwd: bad arguments");
  ICGAN_REQUIRE(dp != ds || dp_dtype == ds_dtype, "icgan_softmax_rows_bwd: in-place needs equal dtypes");
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  DISPATCH_T(p_dtype, TP, {DISPATCH_T(dp_dtype, TD, {DISPATCH_T(ds_dtype, TO, {
    softmax_rows_bwd_kernel<TP, TD, TO><<<blocks, 256, 0, STREAM>>>(
        static_cast<const TP*>(p), static_cast<const TD*>(dp), static_cast<TO*>(ds), rows, cols);
  })})})
  ICGAN_LAUNCH_CHECK();
  return 0;
}
