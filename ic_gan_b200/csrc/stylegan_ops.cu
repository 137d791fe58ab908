// StyleGAN2-ADA's two native ops on B200 (the only CUDA code of the reference, 840 lines under
// stylegan2_ada_pytorch/torch_utils/ops/): bias_act (bias_act.cu:26-150, py :131-171) and upfirdn2d
// (upfirdn2d.cu:32-203, py :145-193).  Both are HBM-bound elementwise/stencil passes; they are restated here with
// 128-bit-friendly flat indexing, any of float32 / bfloat16 / float16 storage and float32 arithmetic.
#include <cuda_fp16.h>

#include "common.cuh"

namespace icgan {

__device__ __forceinline__ float ld_as_float(const __half* p, int64_t i) { return __half2float(p[i]); }
__device__ __forceinline__ void st_from_float(__half* p, int64_t i, float v) { p[i] = __float2half_rn(v); }

struct BiasActParams {
  int64_t n;
  int64_t step_b;  // product of the dims after `dim`
  int size_b;      // x.shape[dim]
  int grad, act;
  float alpha, gain, clamp;
};

// y = clamp(act(x + b) * gain)                                   (grad 0)
// y = dy-like input `x` times act'(.) * gain, masked by the clamp (grad 1; xref = forward input, yref = forward output)
// y = second-order term, multiplied by `dy`                      (grad 2)
// Activation ids and formulas follow bias_act.py:26-99 / bias_act.cu:60-131.
template <typename T>
__global__ void bias_act_kernel(const T* __restrict__ x, const T* __restrict__ b, const T* __restrict__ xref,
                                const T* __restrict__ yref, const T* __restrict__ dy, T* __restrict__ y,
                                BiasActParams p) {
  const float selu_scale = 1.0507009873554804934193349852946f, selu_alpha = 1.6732632423543772848170429916717f;
  const float exp_range = 80.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float xv = ld_as_float(x, i);
    const float bv = b ? ld_as_float(b, (i / p.step_b) % p.size_b) : 0.f;
    float xr = xref ? ld_as_float(xref, i) : 0.f;
    float yr = yref ? ld_as_float(yref, i) : 0.f;
    const float dyv = dy ? ld_as_float(dy, i) : 1.f;
    const float yy = p.gain != 0.f ? yr / p.gain : 0.f;
    const int G = p.grad;
    if (G == 0) xv += bv; else xr += bv;
    float out = 0.f;
    switch (p.act) {
      case 1: if (G <= 1) out = xv; break;                                                       // linear
      case 2: if (G == 0) out = xv > 0.f ? xv : 0.f; else if (G == 1) out = yy > 0.f ? xv : 0.f; break;  // relu
      case 3:                                                                                    // lrelu
        if (G == 0) out = xv > 0.f ? xv : xv * p.alpha;
        else if (G == 1) out = yy > 0.f ? xv : xv * p.alpha;
        break;
      case 4:                                                                                    // tanh
        if (G == 0) { const float c = expf(xv), d = 1.f / c; out = xv < -exp_range ? -1.f : (xv > exp_range ? 1.f : (c - d) / (c + d)); }
        else if (G == 1) out = xv * (1.f - yy * yy);
        else out = xv * (1.f - yy * yy) * (-2.f * yy);
        break;
      case 5:                                                                                    // sigmoid
        if (G == 0) out = xv < -exp_range ? 0.f : 1.f / (expf(-xv) + 1.f);
        else if (G == 1) out = xv * yy * (1.f - yy);
        else out = xv * yy * (1.f - yy) * (1.f - 2.f * yy);
        break;
      case 6:                                                                                    // elu
        if (G == 0) out = xv >= 0.f ? xv : expf(xv) - 1.f;
        else if (G == 1) out = yy >= 0.f ? xv : xv * (yy + 1.f);
        else out = yy >= 0.f ? 0.f : xv * (yy + 1.f);
        break;
      case 7:                                                                                    // selu
        if (G == 0) out = xv >= 0.f ? selu_scale * xv : (selu_scale * selu_alpha) * (expf(xv) - 1.f);
        else if (G == 1) out = yy >= 0.f ? xv * selu_scale : xv * (yy + selu_scale * selu_alpha);
        else out = yy >= 0.f ? 0.f : xv * (yy + selu_scale * selu_alpha);
        break;
      case 8:                                                                                    // softplus
        if (G == 0) out = xv > exp_range ? xv : logf(expf(xv) + 1.f);
        else if (G == 1) out = xv * (1.f - expf(-yy));
        else { const float c = expf(-yy); out = xv * c * (1.f - c); }
        break;
      case 9:                                                                                    // swish
        if (G == 0) {
          out = xv < -exp_range ? 0.f : xv / (expf(-xv) + 1.f);
        } else {
          const float c = expf(xr), d = c + 1.f;
          if (G == 1) out = xr > 40.f ? xv : xv * c * (xr + d) / (d * d);
          else out = xr > 40.f ? 0.f : xv * c * (xr * (2.f - d) + 2.f * d) / (d * d * d);
          yr = xr < -exp_range ? 0.f : xr / (expf(-xr) + 1.f) * p.gain;
        }
        break;
      default: break;
    }
    out *= p.gain * dyv;
    if (p.clamp >= 0.f) {
      if (G == 0) out = (out > -p.clamp && out < p.clamp) ? out : (out >= 0.f ? p.clamp : -p.clamp);
      else out = (yr > -p.clamp && yr < p.clamp) ? out : 0.f;
    }
    st_from_float(y, i, out);
  }
}

struct UpfirdnParams {
  int N, C, inH, inW, outH, outW;
  int upx, upy, downx, downy, padx0, pady0, fw, fh, flip;
  float gain;
  int channels_last;
};

// y[n,c,oy,ox] = gain * sum_{ty,tx} f'[ty,tx] * u[oy*downy + ty, ox*downx + tx],  u = zero-upsampled, padded x,
// f' = f flipped in both axes unless `flip` (true convolution by default, upfirdn2d.py:231-234).
template <typename T>
__global__ void upfirdn2d_kernel(const T* __restrict__ x, const float* __restrict__ f, T* __restrict__ y,
                                 UpfirdnParams p) {
  const int64_t total = static_cast<int64_t>(p.N) * p.C * p.outH * p.outW;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int ox, oy, c, n;
    if (!p.channels_last) {
      ox = static_cast<int>(i % p.outW);
      oy = static_cast<int>((i / p.outW) % p.outH);
      c = static_cast<int>((i / (static_cast<int64_t>(p.outW) * p.outH)) % p.C);
      n = static_cast<int>(i / (static_cast<int64_t>(p.outW) * p.outH * p.C));
    } else {
      c = static_cast<int>(i % p.C);
      ox = static_cast<int>((i / p.C) % p.outW);
      oy = static_cast<int>((i / (static_cast<int64_t>(p.C) * p.outW)) % p.outH);
      n = static_cast<int>(i / (static_cast<int64_t>(p.C) * p.outW * p.outH));
    }
    // first tap whose source position is a multiple of `up`:  (o*down + t - pad0) % up == 0
    const int basey = oy * p.downy - p.pady0, basex = ox * p.downx - p.padx0;
    int ty0 = ((-basey) % p.upy + p.upy) % p.upy, tx0 = ((-basex) % p.upx + p.upx) % p.upx;
    float acc = 0.f;
    for (int ty = ty0; ty < p.fh; ty += p.upy) {
      const int sy = (basey + ty) / p.upy;
      if (sy < 0 || sy >= p.inH) continue;
      const int fy = p.flip ? ty : p.fh - 1 - ty;
      for (int tx = tx0; tx < p.fw; tx += p.upx) {
        const int sx = (basex + tx) / p.upx;
        if (sx < 0 || sx >= p.inW) continue;
        const int fx = p.flip ? tx : p.fw - 1 - tx;
        const int64_t xi = !p.channels_last
                               ? ((static_cast<int64_t>(n) * p.C + c) * p.inH + sy) * p.inW + sx
                               : ((static_cast<int64_t>(n) * p.inH + sy) * p.inW + sx) * p.C + c;
        acc = fmaf(ld_as_float(x, xi), f[fy * p.fw + fx], acc);
      }
    }
    st_from_float(y, i, acc * p.gain);
  }
}

}  // namespace icgan

using namespace icgan;
#define STREAM static_cast<cudaStream_t>(stream)

static int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int icgan_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                              int64_t n, int64_t step_b, int size_b, int grad, int act, float alpha, float gain,
                              float clamp, int dtype, void* stream) {
  ICGAN_REQUIRE(x && y && n > 0, "icgan_bias_act: bad arguments");
  ICGAN_REQUIRE(act >= 1 && act <= 9 && grad >= 0 && grad <= 2, "icgan_bias_act: act in 1..9, grad in 0..2");
  ICGAN_REQUIRE(!b || (step_b > 0 && size_b > 0), "icgan_bias_act: bad bias geometry");
  BiasActParams p{n, step_b > 0 ? step_b : 1, size_b > 0 ? size_b : 1, grad, act, alpha, gain, clamp};
  const int blocks = ew_grid(n);
  if (dtype == ICGAN_BF16)
    bias_act_kernel<__nv_bfloat16><<<blocks, 256, 0, STREAM>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(b), static_cast<const __nv_bfloat16*>(xref),
        static_cast<const __nv_bfloat16*>(yref), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(y), p);
  else if (dtype == ICGAN_F16)
    bias_act_kernel<__half><<<blocks, 256, 0, STREAM>>>(static_cast<const __half*>(x), static_cast<const __half*>(b),
                                                        static_cast<const __half*>(xref), static_cast<const __half*>(yref),
                                                        static_cast<const __half*>(dy), static_cast<__half*>(y), p);
  else
    bias_act_kernel<float><<<blocks, 256, 0, STREAM>>>(static_cast<const float*>(x), static_cast<const float*>(b),
                                                       static_cast<const float*>(xref), static_cast<const float*>(yref),
                                                       static_cast<const float*>(dy), static_cast<float*>(y), p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_upfirdn2d(const void* x, const float* f, void* y, int N, int C, int inH, int inW, int fh, int fw,
                               int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                               int flip, float gain, int channels_last, int dtype, void* stream) {
  ICGAN_REQUIRE(x && f && y, "icgan_upfirdn2d: null pointer");
  ICGAN_REQUIRE(upx >= 1 && upy >= 1 && downx >= 1 && downy >= 1 && fh >= 1 && fw >= 1, "icgan_upfirdn2d: bad factors");
  UpfirdnParams p{};
  p.N = N; p.C = C; p.inH = inH; p.inW = inW;
  p.outW = (inW * upx + padx0 + padx1 - fw + downx) / downx;  // upfirdn2d.cpp:35-36
  p.outH = (inH * upy + pady0 + pady1 - fh + downy) / downy;
  ICGAN_REQUIRE(p.outW >= 1 && p.outH >= 1, "icgan_upfirdn2d: empty output");
  p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.fw = fw; p.fh = fh;
  p.flip = flip; p.gain = gain; p.channels_last = channels_last;
  const int64_t total = static_cast<int64_t>(N) * C * p.outH * p.outW;
  const int blocks = ew_grid(total);
  if (dtype == ICGAN_BF16)
    upfirdn2d_kernel<__nv_bfloat16><<<blocks, 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x), f,
                                                                static_cast<__nv_bfloat16*>(y), p);
  else if (dtype == ICGAN_F16)
    upfirdn2d_kernel<__half><<<blocks, 256, 0, STREAM>>>(static_cast<const __half*>(x), f, static_cast<__half*>(y), p);
  else
    upfirdn2d_kernel<float><<<blocks, 256, 0, STREAM>>>(static_cast<const float*>(x), f, static_cast<float*>(y), p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
