// bf16 fast paths of the HBM-bound NHWC kernels (norm_act.cu): every thread owns 8 consecutive channels of one pixel,
// i.e. one 16-byte load/store per tensor access; the C/8 threads of a pixel and consecutive pixels are contiguous, so
// warps issue full 128-byte lines.  Reductions keep 8 (or 16) float accumulators per thread, combine the pixel lanes of
// a block through shared memory and finish with one float atomic per channel and block.  Requires C % 8 == 0.
#pragma once
#include "common.cuh"

namespace icgan {
namespace {  // internal linkage: this header is included by more than one translation unit
namespace vec {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void cvt8(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
constexpr int kUnroll = 4;  // pixels per thread and loop trip in the streaming kernels (memory-level parallelism)
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ void ldf8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

constexpr int kThreads = 256;

// block-level: sum acc[NA] over the pixel lanes (py) of the block, then atomicAdd to dst[j][c]: the partials go to
// shared memory as [py][C]; thread c adds up column c (conflict-free, one channel per thread) and issues one atomic.
template <int NA>
__device__ __forceinline__ void block_reduce_atomic(float (&acc)[NA][8], float* const (&dst)[NA], int /*c0*/, int cg, int py,
                                                    int CG, int PY, bool active) {
  __shared__ float sm[kThreads * 8];
  const int C = CG * 8;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[py * C + cg * 8 + i] = acc[j][i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      float s = 0.f;
#pragma unroll 4
      for (int y = 0; y < PY; ++y) s += sm[y * C + c];
      atomicAdd(dst[j] + c, s);
    }
  }
}

// ---- single pass: ws[c] += sum (x - shift_c), ws[C+c] += sum (x - shift_c)^2   (shift ~ the channel mean keeps
//      E[d^2] - E[d]^2 well conditioned in float32; NULL = 0)
__global__ void __launch_bounds__(kThreads)
shifted_moments_vec_kernel(const bf16* __restrict__ x, const float* __restrict__ shift, float* __restrict__ ws,
                           int64_t P, int C, int64_t pix_per_block) {
  const int CG = C >> 3, PY = kThreads / CG;
  const int cg = threadIdx.x % CG, py = threadIdx.x / CG;
  const bool active = py < PY;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < P ? p0 + pix_per_block : P;
  float acc[2][8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = sh[i] = 0.f;
  if (shift) ldf8(shift + cg * 8, sh);
  if (active) {
    const bf16* xc = x + cg * 8;
    for (int64_t p = p0 + py; p < p1; p += kUnroll * PY) {  // kUnroll independent 16-byte loads in flight per thread
      uint4 raw[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (p + u * PY < p1) raw[u] = *reinterpret_cast<const uint4*>(xc + (p + u * PY) * C);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (p + u * PY >= p1) break;
        float v[8];
        cvt8(raw[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[i] - sh[i];
          acc[0][i] += d;
          acc[1][i] = fmaf(d, d, acc[1][i]);
        }
      }
    }
  }
  float* const dst[2] = {ws, ws + C};
  block_reduce_atomic<2>(acc, dst, cg * 8, cg, py, CG, PY, active);
}

// ---- column sums / centred squares over [P][C]:  PASS 1: ws[c] += sum x ; PASS 2: ws[C+c] += sum (x-mean)^2 ;
//      PASS 0: out[c] += sum x (bias gradients)
template <int PASS>
__global__ void __launch_bounds__(kThreads)
colsum_vec_kernel(const bf16* __restrict__ x, float* __restrict__ ws, int64_t P, int C, int64_t pix_per_block) {
  const int CG = C >> 3, PY = kThreads / CG;
  const int cg = threadIdx.x % CG, py = threadIdx.x / CG;
  const bool active = py < PY;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < P ? p0 + pix_per_block : P;
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.f;
  float mean[8];
  if (PASS == 2) {
    ldf8(ws + cg * 8, mean);
    const float invP = 1.f / static_cast<float>(P);
#pragma unroll
    for (int i = 0; i < 8; ++i) mean[i] *= invP;
  }
  if (active) {
    const bf16* xc = x + cg * 8;
    for (int64_t p = p0 + py; p < p1; p += kUnroll * PY) {
      uint4 raw[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (p + u * PY < p1) raw[u] = *reinterpret_cast<const uint4*>(xc + (p + u * PY) * C);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (p + u * PY >= p1) break;
        float v[8];
        cvt8(raw[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (PASS == 2) {
            const float d = v[i] - mean[i];
            acc[0][i] = fmaf(d, d, acc[0][i]);
          } else {
            acc[0][i] += v[i];
          }
        }
      }
    }
  }
  float* const dst[1] = {PASS == 2 ? ws + C : ws};
  block_reduce_atomic<1>(acc, dst, cg * 8, cg, py, CG, PY, active);
}

// ---- y = [up2]([relu](xhat * gain[n,c] + bias[n,c]))      grid: (pixel slabs, n)
// A thread keeps one 8-channel group (its mean / invstd / gain / bias live in registers: re-reading them per element
// made the kernel L1-bandwidth bound at ~3.6 TB/s) and walks the pixels of its slab, kUnroll loads in flight.
__global__ void __launch_bounds__(kThreads)
bn_apply_vec_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gain, const float* __restrict__ bias,
                    int gstride, int H, int W, int C, int relu, int up, int pix_per_block) {
  const int CG = C >> 3, PY = kThreads / CG;
  const int cg = threadIdx.x % CG, py = threadIdx.x / CG;
  if (py >= PY) return;
  const int n = blockIdx.y, c0 = cg * 8;
  const int HW = H * W;
  const int q0 = blockIdx.x * pix_per_block, q1 = min(HW, q0 + pix_per_block);
  float m[8], is[8], g[8], b[8];
  ldf8(mean + c0, m);
  ldf8(invstd + c0, is);
  ldf8(gain + static_cast<int64_t>(n) * gstride + c0, g);
  ldf8(bias + static_cast<int64_t>(n) * gstride + c0, b);
  const bf16* xn = x + static_cast<int64_t>(n) * HW * C + c0;
  bf16* yn = y + static_cast<int64_t>(n) * HW * C * (up ? 4 : 1) + c0;
  for (int q = q0 + py; q < q1; q += kUnroll * PY) {
    uint4 raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (q + u * PY < q1) raw[u] = *reinterpret_cast<const uint4*>(xn + static_cast<int64_t>(q + u * PY) * C);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int qq = q + u * PY;
      if (qq >= q1) break;
      float v[8];
      cvt8(raw[u], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float o = (v[i] - m[i]) * is[i] * g[i] + b[i];
        v[i] = relu ? fmaxf(o, 0.f) : o;
      }
      if (!up) {
        st8(yn + static_cast<int64_t>(qq) * C, v);
      } else {
        const int h = qq / W, w = qq - h * W;
        bf16* o = yn + (static_cast<int64_t>(2 * h) * 2 * W + 2 * w) * C;
        const int64_t rs = static_cast<int64_t>(2) * W * C;
        st8(o, v);
        st8(o + C, v);
        st8(o + rs, v);
        st8(o + rs + C, v);
      }
    }
  }
}

// gradient at the BN-affine output for 8 channels (ReLU mask recomputed, 2x2 children summed when upsampled)
__device__ __forceinline__ void bn_grad8(const bf16* x, const bf16* dy, const float (&m)[8], const float (&is)[8],
                                         const float (&g)[8], const float (&b)[8], int n, int h, int w, int c0, int H,
                                         int W, int C, int relu, int up, float (&xh)[8], float (&gr)[8]) {
  float v[8];
  ld8(x + ((static_cast<int64_t>(n) * H + h) * W + w) * C + c0, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) xh[i] = (v[i] - m[i]) * is[i];
  if (!up) {
    ld8(dy + ((static_cast<int64_t>(n) * H + h) * W + w) * C + c0, gr);
  } else {
    const bf16* o = dy + ((static_cast<int64_t>(n) * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c0;
    const int64_t rs = static_cast<int64_t>(2) * W * C;
    float a[8], bb[8], cc[8], dd[8];
    ld8(o, a); ld8(o + C, bb); ld8(o + rs, cc); ld8(o + rs + C, dd);
#pragma unroll
    for (int i = 0; i < 8; ++i) gr[i] = (a[i] + bb[i]) + (cc[i] + dd[i]);
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!(xh[i] * g[i] + b[i] > 0.f)) gr[i] = 0.f;
  }
}

// s1[n,c] += sum_hw g ; s2[n,c] += sum_hw g*xhat    grid: (pixel slabs, n)
// kPlain: dy has the resolution of x (host passes up == 0); upsampled layers sum 2x2 children of dy per pixel.
template <bool kPlain>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_vec_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gain,
                         const float* __restrict__ bias, int gstride, float* __restrict__ s1, float* __restrict__ s2,
                         int H, int W, int C, int relu, int up, int pix_per_block) {
  const int CG = C >> 3, PY = kThreads / CG;
  const int cg = threadIdx.x % CG, py = threadIdx.x / CG;
  const bool active = py < PY;
  const int n = blockIdx.y, c0 = cg * 8;
  const int HW = H * W;
  const int q0 = blockIdx.x * pix_per_block, q1 = min(HW, q0 + pix_per_block);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    float m[8], is[8], g[8], b[8];
    ldf8(mean + c0, m);
    ldf8(invstd + c0, is);
    ldf8(gain + static_cast<int64_t>(n) * gstride + c0, g);
    ldf8(bias + static_cast<int64_t>(n) * gstride + c0, b);
    if constexpr (kPlain) {
      // plain (not upsampled) layer: kUnroll pixels per trip; every load is issued (index clamped, so unconditional)
      // before anything is consumed -- 2 * kUnroll 16-byte loads in flight per thread.
      const bf16* xn = x + static_cast<int64_t>(n) * HW * C + c0;
      const bf16* dn = dy + static_cast<int64_t>(n) * HW * C + c0;
      for (int q = q0 + py; q < q1; q += kUnroll * PY) {
        uint4 rx[kUnroll], rd[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int64_t o = static_cast<int64_t>(min(q + u * PY, q1 - 1)) * C;
          rx[u] = *reinterpret_cast<const uint4*>(xn + o);
          rd[u] = *reinterpret_cast<const uint4*>(dn + o);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (q + u * PY >= q1) break;
          float v[8], gr[8];
          cvt8(rx[u], v);
          cvt8(rd[u], gr);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = (v[i] - m[i]) * is[i];
            if (relu && !(xh * g[i] + b[i] > 0.f)) gr[i] = 0.f;
            acc[0][i] += gr[i];
            acc[1][i] = fmaf(gr[i], xh, acc[1][i]);
          }
        }
      }
    } else {
      for (int q = q0 + py; q < q1; q += PY) {
        float xh[8], gr[8];
        bn_grad8(x, dy, m, is, g, b, n, q / W, q % W, c0, H, W, C, relu, up, xh, gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += gr[i];
          acc[1][i] = fmaf(gr[i], xh[i], acc[1][i]);
        }
      }
    }
  }
  float* const dst[2] = {s1 + static_cast<int64_t>(n) * C, s2 + static_cast<int64_t>(n) * C};
  block_reduce_atomic<2>(acc, dst, c0, cg, py, CG, PY, active);
}

// dx = invstd * (gain * g - m1 - xhat * m2)      grid: (pixel slabs, n); per-channel parameters in registers
template <bool kPlain>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_vec_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gain, const float* __restrict__ bias, int gstride,
                        const float* __restrict__ m1, const float* __restrict__ m2, int H, int W, int C, int relu,
                        int up, int pix_per_block) {
  const int CG = C >> 3, PY = kThreads / CG;
  const int cg = threadIdx.x % CG, py = threadIdx.x / CG;
  if (py >= PY) return;
  const int n = blockIdx.y, c0 = cg * 8;
  const int HW = H * W;
  const int q0 = blockIdx.x * pix_per_block, q1 = min(HW, q0 + pix_per_block);
  float m[8], is[8], g[8], b[8], a1[8], a2[8];
  ldf8(mean + c0, m);
  ldf8(invstd + c0, is);
  ldf8(gain + static_cast<int64_t>(n) * gstride + c0, g);
  ldf8(bias + static_cast<int64_t>(n) * gstride + c0, b);
  ldf8(m1 + c0, a1);
  ldf8(m2 + c0, a2);
  bf16* dxn = dx + static_cast<int64_t>(n) * HW * C + c0;
  if constexpr (kPlain) {
    const bf16* xn = x + static_cast<int64_t>(n) * HW * C + c0;
    const bf16* dn = dy + static_cast<int64_t>(n) * HW * C + c0;
    for (int q = q0 + py; q < q1; q += kUnroll * PY) {
      uint4 rx[kUnroll], rd[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t o = static_cast<int64_t>(min(q + u * PY, q1 - 1)) * C;
        rx[u] = *reinterpret_cast<const uint4*>(xn + o);
        rd[u] = *reinterpret_cast<const uint4*>(dn + o);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int qq = q + u * PY;
        if (qq >= q1) break;
        float v[8], gr[8], o[8];
        cvt8(rx[u], v);
        cvt8(rd[u], gr);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (v[i] - m[i]) * is[i];
          if (relu && !(xh * g[i] + b[i] > 0.f)) gr[i] = 0.f;
          o[i] = is[i] * (g[i] * gr[i] - a1[i] - xh * a2[i]);
        }
        st8(dxn + static_cast<int64_t>(qq) * C, o);
      }
    }
  } else {
    for (int q = q0 + py; q < q1; q += PY) {
      float xh[8], gr[8], o[8];
      bn_grad8(x, dy, m, is, g, b, n, q / W, q % W, c0, H, W, C, relu, up, xh, gr);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = is[i] * (g[i] * gr[i] - a1[i] - xh[i] * a2[i]);
      st8(dxn + static_cast<int64_t>(q) * C, o);
    }
  }
}

// ---- flat elementwise (n % 8 == 0): op 0 relu, 1 relu_bwd(dy, ref), 2 tanh_bwd(dy, y), 3 axpby(alpha*a + beta*b)
template <int OP>
__global__ void __launch_bounds__(kThreads)
ew_vec_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, float alpha, float beta,
              const float* alpha_p, const float* beta_p, int64_t n8) {
  const float al = alpha_p ? *alpha_p : alpha, be = beta_p ? *beta_p : beta;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n8;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float u[8], v[8];
    ld8(a + t * 8, u);
    if (OP != 0 && (OP != 3 || b)) ld8(b + t * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) u[i] = fmaxf(u[i], 0.f);
      else if (OP == 1) u[i] = v[i] > 0.f ? u[i] : 0.f;
      else if (OP == 2) u[i] = u[i] * (1.f - v[i] * v[i]);
      else u[i] = b ? fmaf(be, v[i], al * u[i]) : al * u[i];
    }
    st8(out + t * 8, u);
  }
}

// ---- 2x2 pooling: mode 0 scale*sum(+add), mode 1 max
__global__ void __launch_bounds__(kThreads)
pool2_vec_kernel(const bf16* __restrict__ x, const bf16* __restrict__ add, bf16* __restrict__ y, int B, int Ho, int Wo,
                 int C, float scale, int mode) {
  const int CG = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * CG;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(t % CG) * 8;
    const int64_t pix = t / CG;
    const int w = static_cast<int>(pix % Wo);
    const int h = static_cast<int>((pix / Wo) % Ho);
    const int64_t n = pix / (static_cast<int64_t>(Ho) * Wo);
    const bf16* o = x + ((n * 2 * Ho + 2 * h) * 2 * Wo + 2 * w) * C + c0;
    const int64_t rs = static_cast<int64_t>(2) * Wo * C;
    float a[8], b[8], c[8], d[8], r[8];
    ld8(o, a); ld8(o + C, b); ld8(o + rs, c); ld8(o + rs + C, d);
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = ((a[i] + b[i]) + (c[i] + d[i])) * scale;
      if (add) {
        float e[8];
        ld8(add + pix * C + c0, e);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] += e[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaxf(fmaxf(a[i], b[i]), fmaxf(c[i], d[i]));
    }
    st8(y + pix * C + c0, r);
  }
}
// mode 0: dx = scale * up2(dy) ; mode 1: max-pool backward (first maximal element in scan order)
__global__ void __launch_bounds__(kThreads)
unpool2_vec_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ xref, bf16* __restrict__ dx, int B, int Ho,
                   int Wo, int C, float scale, int mode) {
  const int CG = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * CG;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(t % CG) * 8;
    const int64_t pix = t / CG;
    const int w = static_cast<int>(pix % Wo);
    const int h = static_cast<int>((pix / Wo) % Ho);
    const int64_t n = pix / (static_cast<int64_t>(Ho) * Wo);
    const int64_t off = ((n * 2 * Ho + 2 * h) * 2 * Wo + 2 * w) * C + c0;
    const int64_t rs = static_cast<int64_t>(2) * Wo * C;
    float g[8];
    ld8(dy + pix * C + c0, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] *= scale;
    if (mode == 0) {
      st8(dx + off, g);
      st8(dx + off + C, g);
      st8(dx + off + rs, g);
      st8(dx + off + rs + C, g);
    } else {
      float v[4][8], o[4][8];
      ld8(xref + off, v[0]); ld8(xref + off + C, v[1]); ld8(xref + off + rs, v[2]); ld8(xref + off + rs + C, v[3]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int best = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][i] > v[best][i]) best = k;
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][i] = best == k ? g[i] : 0.f;
      }
      st8(dx + off, o[0]);
      st8(dx + off + C, o[1]);
      st8(dx + off + rs, o[2]);
      st8(dx + off + rs + C, o[3]);
    }
  }
}

// out[n,c] += sum_hw relu(x)     grid: (pixel slabs, n)
__global__ void __launch_bounds__(kThreads)
relu_sumpool_vec_kernel(const bf16* __restrict__ x, float* __restrict__ out, int HW, int C, int pix_per_block) {
  const int CG = C >> 3, PY = kThreads / CG;
  const int cg = threadIdx.x % CG, py = threadIdx.x / CG;
  const bool active = py < PY;
  const int n = blockIdx.y;
  const int q0 = blockIdx.x * pix_per_block, q1 = min(HW, q0 + pix_per_block);
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.f;
  if (active) {
    for (int q = q0 + py; q < q1; q += PY) {
      float v[8];
      ld8(x + (static_cast<int64_t>(n) * HW + q) * C + cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[0][i] += fmaxf(v[i], 0.f);
    }
  }
  float* const dst[1] = {out + static_cast<int64_t>(n) * C};
  block_reduce_atomic<1>(acc, dst, cg * 8, cg, py, CG, PY, active);
}

inline int blocks_for(int64_t work_items) {
  int64_t b = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}
inline bool ok(int C) { return C % 8 == 0 && C / 8 <= kThreads; }
// (pixel slabs, samples) grid of ~16 blocks per SM; returns pixels per block
inline int slab_grid(int B, int HW, int* slabs_out) {
  int slabs = (16 * num_sms() + B - 1) / B;
  if (slabs > (HW + 31) / 32) slabs = (HW + 31) / 32;
  if (slabs < 1) slabs = 1;
  const int ppb = (HW + slabs - 1) / slabs;
  *slabs_out = (HW + ppb - 1) / ppb;
  return ppb;
}

}  // namespace vec
}  // namespace
}  // namespace icgan
