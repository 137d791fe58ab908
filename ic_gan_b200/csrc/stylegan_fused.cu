// StyleGAN2 hot-path kernels in NHWC (channels-last), written for B200's HBM: 128-bit accesses, one pass per tensor.
//
//  * upfirdn2d_tiled_kernel -- the reference's upfirdn2d_kernel_small (stylegan2_ada_pytorch/torch_utils/ops/
//    upfirdn2d.cu:100-203) re-designed for channels-last: a CTA stages the input patch of an output tile once in shared
//    memory (128 bytes = 64 bf16 / 32 fp32 channels per pixel), each thread walks a vertical strip of outputs with a
//    sliding register window (4x fewer shared-memory reads than one gather per tap), and the layer's epilogue
//    (demodulation coefficient, noise, bias, leaky ReLU, gain, clamp -- networks.py:77-95,441-444, bias_act.cu:26-150)
//    plus the NEXT layer's style modulation (second output y*s2[n,c]) can ride in the same pass.
//  * modulate_kernel / chan_dot_kernel -- x*s[n,c] and its adjoint sum_hw a*b -> [n,c] (the two bilinear pieces every
//    derivative of the modulation is made of).
//  * bias_act_vec_kernel -- bias_act for channels-last tensors, 8 elements per thread, no per-element div/mod, with the
//    same optional per-sample scale and noise inputs.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace icgan {
namespace {

template <typename T> struct Pack;  // 16 bytes of T <-> floats
template <> struct Pack<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
  __device__ static void unpack(const uint4& raw, float (&v)[4]) {
    v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
  }
};
template <> struct Pack<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& raw, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) { unpack(*reinterpret_cast<const uint4*>(p), v); }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
};
template <> struct Pack<__half> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& raw, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static void load(const __half* p, float (&v)[8]) { unpack(*reinterpret_cast<const uint4*>(p), v); }
  __device__ static void store(__half* p, const float (&v)[8]) {
    uint4 raw;
    __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
};

// 8 consecutive elements <-> floats with 128-bit accesses (one uint4 for 16-bit types, two float4 for float32)
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]) { Pack<T>::load(p, v); }
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]) { Pack<T>::store(p, v); }
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

struct PostArgs {          // optional epilogue: y = clamp(act((v * pre[n,c]) + noise[n,oy,ox]*ns + bias[c]) * act_gain); y2 = y*s2[n,c]
  const float* pre;        // [N, C] per-sample scale (demodulation coefficients), may be null
  const float* noise;      // [N or 1, outH, outW] float32, may be null
  const float* noise_strength;  // device scalar (may be null = 1)
  const float* bias;       // [C] float32, may be null
  const float* s2;         // [N, C]: second output y2 = y * s2 (the next layer's modulated input), may be null
  void* y2;
  int noise_per_sample, act;  // act: bias_act ids 1 (linear) / 3 (lrelu); 0 = no epilogue at all
  float alpha, act_gain, clamp;
};

struct UpfirdnTiledParams {
  int N, C, inH, inW, outH, outW, pad0x, pad0y, flip;
  float gain;
  PostArgs post;
  // rank-1 filters (f = fy (x) fx, e.g. the [1,3,3,1] outer product of every StyleGAN2 resampling filter): the kernel
  // filters each row horizontally as it enters the sliding window and vertically per output -- 8 instead of 16
  // tap-FMAs (and bf16 unpacks) per output vector, which is what makes the pass HBM-bound instead of issue-bound
  int separable;
  float fx[4], fy[4];  // already flipped / gain-scaled like fs[] below (gain on fy)
};

// epilogue + store of one output vector (shared by the tiled and the direct kernel)
template <typename T>
__device__ __forceinline__ void emit_output(const UpfirdnTiledParams& p, T* __restrict__ y, int n, int oy, int ox, int c0, float ns,
                                            float (&acc)[Pack<T>::N]) {
  constexpr int NV = Pack<T>::N;
  const int64_t pix = (static_cast<int64_t>(n) * p.outH + oy) * p.outW + ox;
  if (p.post.act) {
    const PostArgs& q = p.post;
    float nz = 0.f;
    if (q.noise) nz = ns * q.noise[(q.noise_per_sample ? static_cast<int64_t>(n) * p.outH * p.outW : 0) + static_cast<int64_t>(oy) * p.outW + ox];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float v = acc[k];
      if (q.pre) v *= q.pre[static_cast<int64_t>(n) * p.C + c0 + k];
      v += nz;
      if (q.bias) v += q.bias[c0 + k];
      if (q.act == 3) v = v > 0.f ? v : v * q.alpha;
      v *= q.act_gain;
      if (q.clamp >= 0.f) v = fminf(fmaxf(v, -q.clamp), q.clamp);
      acc[k] = v;
    }
  }
  Pack<T>::store(y + pix * p.C + c0, acc);
  if (p.post.y2) {
    float v2[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v2[k] = acc[k] * p.post.s2[static_cast<int64_t>(n) * p.C + c0 + k];
    Pack<T>::store(static_cast<T*>(p.post.y2) + pix * p.C + c0, v2);
  }
}

// Separable filters, no shared memory: a thread owns output column ox (8 lanes x 16 bytes of channels) and walks ROWS
// output rows with a sliding window of horizontally pre-filtered rows; its four 16-byte loads per new row come straight
// from global memory through L1 (neighbouring columns share three of them), so loads stay in flight the whole time --
// the tiled kernel alternates load and compute phases behind a barrier and reached only ~0.36 of HBM bandwidth on
// 0.5 GB tensors (profiles/r02_kernel_bandwidth_before_vectorising.txt).
template <typename T, int UP, int DOWN, int ROWS>
__global__ void __launch_bounds__(256)
upfirdn2d_direct_kernel(const T* __restrict__ x, T* __restrict__ y, const UpfirdnTiledParams p) {
  constexpr int NV = Pack<T>::N, CB = 8 * NV, F = 4, TOW = 32;
  const int lane = threadIdx.x & 7, col = threadIdx.x >> 3;
  int bid = blockIdx.x;
  const int cblocks = (p.C + CB - 1) / CB;
  const int cb = bid % cblocks; bid /= cblocks;
  const int tiles_w = (p.outW + TOW - 1) / TOW, tiles_h = (p.outH + ROWS - 1) / ROWS;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h;
  const int n = bid / tiles_h;
  const int ox = tw * TOW + col, oy0 = th * ROWS;
  const int c0 = cb * CB + lane * NV;
  const bool live = c0 < p.C && ox < p.outW;
  const int ux = ox * DOWN - p.pad0x;
  int u_cur = oy0 * DOWN - p.pad0y;
  const T* xn = x + static_cast<int64_t>(n) * p.inH * p.inW * p.C + c0;
  // column offsets and validity of the four horizontal taps never change while the thread walks down its strip
  int col_off[F];
  bool col_ok[F];
#pragma unroll
  for (int t = 0; t < F; ++t) {
    const int u = ux + t;
    int ix = u;
    bool ok = live;
    if (UP == 2) { ok = ok && ((u & 1) == 0); ix = u >> 1; }
    col_ok[t] = ok && ix >= 0 && ix < p.inW;
    col_off[t] = ix * p.C;
  }
  const int row_pitch = p.inW * p.C;
  auto load_hrow = [&](int u_row, float (&dst)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) dst[k] = 0.f;
    int iy = u_row;
    bool row_ok = true;
    if (UP == 2) { row_ok = (u_row & 1) == 0; iy = u_row >> 1; }
    row_ok = row_ok && iy >= 0 && iy < p.inH;
    const T* rowp = xn + static_cast<int64_t>(iy) * row_pitch;
    uint4 raw[F];
#pragma unroll
    for (int t = 0; t < F; ++t)
      raw[t] = (row_ok && col_ok[t]) ? __ldg(reinterpret_cast<const uint4*>(rowp + col_off[t])) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int t = 0; t < F; ++t) {
      float xv[NV];
      Pack<T>::unpack(raw[t], xv);
      const float w = p.fx[t];
#pragma unroll
      for (int k = 0; k < NV; ++k) dst[k] = fmaf(xv[k], w, dst[k]);
    }
  };
  float hwin[F][NV];
#pragma unroll
  for (int r = 0; r < F; ++r) load_hrow(u_cur + r, hwin[r]);
  const float ns = p.post.noise_strength ? *p.post.noise_strength : 1.f;
#pragma unroll 4  // a multiple of the window depth: the row rotation below becomes register renaming, not moves
  for (int ro = 0; ro < ROWS; ++ro) {
    const int oy = oy0 + ro;
    if (oy >= p.outH) break;
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
#pragma unroll
    for (int r = 0; r < F; ++r) {
      const float w = p.fy[r];
#pragma unroll
      for (int k = 0; k < NV; ++k) acc[k] = fmaf(hwin[r][k], w, acc[k]);
    }
    if (live) emit_output<T>(p, y, n, oy, ox, c0, ns, acc);
    u_cur += DOWN;
#pragma unroll
    for (int r = 0; r + DOWN < F; ++r)
#pragma unroll
      for (int k = 0; k < NV; ++k) hwin[r][k] = hwin[r + DOWN][k];
#pragma unroll
    for (int r = F - DOWN; r < F; ++r) load_hrow(u_cur + r, hwin[r]);
  }
}

// Square up/down factors, fh = fw = 4 (every resampling filter StyleGAN2 uses: [1,3,3,1] outer product).
// Tile: TOH x TOW outputs x (8 lanes x 16 bytes) channels.  Thread = (lane, column, row group).
template <typename T, int UP, int DOWN, int TOH, int TOW, bool SEP>
__global__ void __launch_bounds__(256, 2)
upfirdn2d_tiled_kernel(const T* __restrict__ x, const float* __restrict__ f, T* __restrict__ y, const UpfirdnTiledParams p) {
  constexpr int NV = Pack<T>::N;          // channels per lane
  constexpr int CB = 8 * NV;              // channels per CTA
  constexpr int F = 4;
  constexpr int RG = 32 / TOW;            // row groups
  constexpr int ROWS = TOH / RG;          // outputs per thread
  // patch extents in the input domain (upper bounds)
  constexpr int PH = (TOH * DOWN + F - 1 + UP - 1) / UP + 1;
  constexpr int PW = (TOW * DOWN + F - 1 + UP - 1) / UP + 1;
  extern __shared__ uint4 patch[];  // [PH][PW][8 lanes] x 16 bytes
  __shared__ float fs[F * F];

  const int tid = threadIdx.x;
  const int lane = tid & 7, col = (tid >> 3) % TOW, rg = (tid >> 3) / TOW;
  int bid = blockIdx.x;
  const int cblocks = (p.C + CB - 1) / CB;
  const int cb = bid % cblocks; bid /= cblocks;
  const int tiles_w = (p.outW + TOW - 1) / TOW, tiles_h = (p.outH + TOH - 1) / TOH;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h;
  const int n = bid / tiles_h;
  const int oy0 = th * TOH, ox0 = tw * TOW;
  const int c0 = cb * CB + lane * NV;
  const bool c_ok = c0 < p.C;  // C % NV == 0 is required by the launcher

  if (tid < F * F) {
    const int ty = tid / F, tx = tid % F;
    fs[tid] = f[(p.flip ? ty : F - 1 - ty) * F + (p.flip ? tx : F - 1 - tx)] * p.gain;
  }
  // zero-upsampled domain coordinate of the first tap of output o:  u = o*DOWN - pad0 (+ tap)
  const int uy0 = oy0 * DOWN - p.pad0y, ux0 = ox0 * DOWN - p.pad0x;
  // first input row/col the tile can touch: ceil(u / UP)
  const int iy0 = (uy0 >= 0 ? (uy0 + UP - 1) / UP : -((-uy0) / UP));
  const int ix0 = (ux0 >= 0 ? (ux0 + UP - 1) / UP : -((-ux0) / UP));
  for (int i = tid; i < PH * PW * 8; i += 256) {
    const int l = i & 7, pw = (i >> 3) % PW, ph = (i >> 3) / PW;
    const int iy = iy0 + ph, ix = ix0 + pw, c = cb * CB + l * NV;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < p.inH && ix >= 0 && ix < p.inW && c < p.C)
      v = *reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(n) * p.inH + iy) * p.inW + ix) * p.C + c);
    patch[i] = v;
  }
  __syncthreads();

  const int ox = ox0 + col;
  const int ux = ux0 + col * DOWN;  // zero-upsampled column of tap 0
  // sliding window over the zero-upsampled domain: win[r][t] = the 16 raw bytes at (row cursor + r, ux + t); kept packed
  // (64 registers instead of 128 for bf16) so that two or three CTAs fit an SM
  uint4 win[F][F];
  auto load_row = [&](int u_row, uint4 (&dst)[F]) {
#pragma unroll
    for (int t = 0; t < F; ++t) {
      const int u = ux + t;
      bool ok = c_ok;
      int iy = 0, ix = 0;
      if (UP == 1) { iy = u_row - iy0; ix = u - ix0; }
      else {
        ok = ok && ((u_row & (UP - 1)) == 0) && ((u & (UP - 1)) == 0);
        iy = (u_row >> 1) - iy0; ix = (u >> 1) - ix0;  // UP == 2; arithmetic shift = floor for negatives
      }
      ok = ok && iy >= 0 && iy < PH && ix >= 0 && ix < PW;
      dst[t] = ok ? patch[(iy * PW + ix) * 8 + lane] : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  const int row_first = rg * ROWS;
  int u_cur = uy0 + row_first * DOWN;
  // separable form: hwin[r] = horizontally filtered row (float), filled as rows enter the window
  float hwin[F][NV];
  auto load_hrow = [&](int u_row, float (&dst)[NV]) {
    uint4 raw[F];
    load_row(u_row, raw);
#pragma unroll
    for (int k = 0; k < NV; ++k) dst[k] = 0.f;
#pragma unroll
    for (int t = 0; t < F; ++t) {
      float xv[NV];
      Pack<T>::unpack(raw[t], xv);
      const float w = p.fx[t];
#pragma unroll
      for (int k = 0; k < NV; ++k) dst[k] = fmaf(xv[k], w, dst[k]);
    }
  };
  if constexpr (SEP) {
#pragma unroll
    for (int r = 0; r < F; ++r) load_hrow(u_cur + r, hwin[r]);
  } else {
#pragma unroll
    for (int r = 0; r < F; ++r) load_row(u_cur + r, win[r]);
  }
  const float ns = p.post.noise_strength ? *p.post.noise_strength : 1.f;
#pragma unroll
  for (int ro = 0; ro < ROWS; ++ro) {
    const int oy = oy0 + row_first + ro;
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
    if constexpr (SEP) {
#pragma unroll
      for (int r = 0; r < F; ++r) {
        const float w = p.fy[r];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = fmaf(hwin[r][k], w, acc[k]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < F; ++r)
#pragma unroll
        for (int t = 0; t < F; ++t) {
          const float w = fs[r * F + t];
          float xv[NV];
          Pack<T>::unpack(win[r][t], xv);
#pragma unroll
          for (int k = 0; k < NV; ++k) acc[k] = fmaf(xv[k], w, acc[k]);
        }
    }
    if (c_ok && oy < p.outH && ox < p.outW) emit_output<T>(p, y, n, oy, ox, c0, ns, acc);
    // advance the window by DOWN rows
    if (ro + 1 < ROWS) {
      u_cur += DOWN;
      if constexpr (SEP) {
#pragma unroll
        for (int r = 0; r + DOWN < F; ++r)
#pragma unroll
          for (int k = 0; k < NV; ++k) hwin[r][k] = hwin[r + DOWN][k];
#pragma unroll
        for (int r = F - DOWN; r < F; ++r) load_hrow(u_cur + r, hwin[r]);
      } else {
#pragma unroll
        for (int r = 0; r + DOWN < F; ++r)
#pragma unroll
          for (int t = 0; t < F; ++t) win[r][t] = win[r + DOWN][t];
#pragma unroll
        for (int r = F - DOWN; r < F; ++r) load_row(u_cur + r, win[r]);
      }
    }
  }
}

// y[n,p,c] = x[n,p,c] * s[n,c]   (optionally cast: TI -> TO)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
modulate_kernel(const TI* __restrict__ x, const float* __restrict__ s, TO* __restrict__ y, int64_t hw, int C, int64_t total8) {
  const int cv = C / 8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 8;
    const int64_t n = (i / cv) / hw;
    float v[8];
    load8(x + i * 8, v);
    const float4 s0 = *reinterpret_cast<const float4*>(s + n * C + c), s1 = *reinterpret_cast<const float4*>(s + n * C + c + 4);
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    store8(y + i * 8, v);
  }
}

// out[n,c] += sum over pixels of a[n,p,c] * b[n,p,c]; grid = (pixel slabs, N); blockDim 256 = 8 rows x 32 channel-vector lanes
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
chan_dot_kernel(const TA* __restrict__ a, const TB* __restrict__ b, float* __restrict__ out, int64_t hw, int C, int slab) {
  const int n = blockIdx.y;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * slab, p1 = min(hw, p0 + slab);
  const int cv = C / 8;
  __shared__ float red[8][33 * 8];
  for (int v0 = 0; v0 < cv; v0 += 32) {
    const int v = v0 + (threadIdx.x & 31);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (v < cv)
      for (int64_t px = p0 + (threadIdx.x >> 5); px < p1; px += 8) {
        const int64_t o = (static_cast<int64_t>(n) * hw + px) * C + v * 8;
        float av[8], bv[8];
        load8(a + o, av);
        load8(b + o, bv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(av[k], bv[k], acc[k]);
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x >> 5][(threadIdx.x & 31) * 8 + k] = acc[k];
    __syncthreads();
    if (threadIdx.x < 256) {
      const int idx = threadIdx.x;  // 256 = 32 lanes x 8 channels
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += red[r][idx];
      const int c = (v0 + idx / 8) * 8 + (idx & 7);
      if (c < C) atomicAdd(out + static_cast<int64_t>(n) * C + c, t);
    }
    __syncthreads();
  }
}

// channels-last bias_act (lrelu / linear, grad 0 or 1) with optional per-sample scale and noise, 8 elements per thread.
//   grad 0: y = clamp(act(x*pre[n,c] + noise[n,p]*ns + b[c]) * gain)
//   grad 1: y = x(=dy) * gain * act'(yref) masked by the clamp            (bias_act.cu:60-131, activations 1 and 3)
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_vec_kernel(const T* __restrict__ x, const T* __restrict__ yref, T* __restrict__ y, const float* __restrict__ bias,
                    const float* __restrict__ pre, const float* __restrict__ noise, const float* __restrict__ noise_strength,
                    int noise_per_sample, int64_t hw, int C, int64_t total8, int grad, int act, float alpha, float gain,
                    float clamp) {
  const int cv = C / 8;
  const float ns = noise_strength ? *noise_strength : 1.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 8;
    const int64_t pix = i / cv;
    const int64_t n = pix / hw;
    float v[8], yr[8], pv[8], bv[8];
    load8(x + i * 8, v);
    if (pre) load8(pre + n * C + c, pv);
    if (grad == 0) {
      float nz = 0.f;
      if (noise) nz = ns * noise[noise_per_sample ? pix : pix - n * hw];
      if (bias) load8(bias + c, bv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = v[k];
        if (pre) t *= pv[k];
        t += nz;
        if (bias) t += bv[k];
        if (act == 3) t = t > 0.f ? t : t * alpha;
        t *= gain;
        if (clamp >= 0.f) t = fminf(fmaxf(t, -clamp), clamp);
        v[k] = t;
      }
    } else {
      if (yref) load8(yref + i * 8, yr);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) yr[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = v[k];
        if (act == 3) t = yr[k] > 0.f ? t : t * alpha;  // sign(y) = sign(pre-activation) since gain > 0
        t *= gain;
        if (clamp >= 0.f) t = (yr[k] > -clamp && yr[k] < clamp) ? t : 0.f;
        if (pre) t *= pv[k];  // chain rule through the per-sample scale: d/dx
        v[k] = t;
      }
    }
    store8(y + i * 8, v);
  }
}

// bias_act / demodulation epilogue with the per-sample and per-channel parameters resident in registers: a thread owns one
// 8-channel vector of one sample and walks pixels (grid = (pixel slabs, N), 256 threads = 256/cv pixel rows x cv vectors).
// The flat kernel above re-loads pre[n,c..c+8) and bias[c..c+8) per element vector -- 64 bytes of L1 requests next to 16
// bytes of payload, which capped it at 0.46 of the HBM peak on 0.5 GB tensors (profiles/r02_kernel_bandwidth.txt).
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_slab_kernel(const T* __restrict__ x, const T* __restrict__ yref, T* __restrict__ y, const float* __restrict__ bias,
                     const float* __restrict__ pre, const float* __restrict__ noise, const float* __restrict__ noise_strength,
                     int noise_per_sample, int64_t hw, int C, int slab, int grad, int act, float alpha, float gain, float clamp) {
  const int cv = C / 8, pp = 256 / cv;
  const int lc = threadIdx.x % cv, prow = threadIdx.x / cv;
  const int n = blockIdx.y, c = lc * 8;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * slab, p1 = min(hw, p0 + slab);
  float pv[8], bv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { pv[k] = 1.f; bv[k] = 0.f; }
  if (pre) load8(pre + static_cast<int64_t>(n) * C + c, pv);
  if (bias && grad == 0) load8(bias + c, bv);
  const float ns = noise_strength ? *noise_strength : 1.f;
  const float* nz_row = noise ? noise + (noise_per_sample ? static_cast<int64_t>(n) * hw : 0) : nullptr;
  const T* xb = x + static_cast<int64_t>(n) * hw * C + c;
  const T* yb = yref ? yref + static_cast<int64_t>(n) * hw * C + c : nullptr;
  T* ob = y + static_cast<int64_t>(n) * hw * C + c;
  for (int64_t px = p0 + prow; px < p1; px += 2 * pp) {  // two independent pixels per trip: more loads in flight
    const int64_t px2 = px + pp;
    const bool two = px2 < p1;
    float v0[8], v1[8], r0[8], r1[8];
    load8(xb + px * C, v0);
    if (two) load8(xb + px2 * C, v1);
    if (grad == 1 && yb) {
      load8(yb + px * C, r0);
      if (two) load8(yb + px2 * C, r1);
    }
    float nz0 = 0.f, nz1 = 0.f;
    if (nz_row && grad == 0) {
      nz0 = ns * nz_row[px];
      if (two) nz1 = ns * nz_row[px2];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (grad == 0) {
        float t0 = fmaf(v0[k], pv[k], nz0 + bv[k]), t1 = fmaf(v1[k], pv[k], nz1 + bv[k]);
        if (act == 3) { t0 = t0 > 0.f ? t0 : t0 * alpha; t1 = t1 > 0.f ? t1 : t1 * alpha; }
        t0 *= gain; t1 *= gain;
        if (clamp >= 0.f) { t0 = fminf(fmaxf(t0, -clamp), clamp); t1 = fminf(fmaxf(t1, -clamp), clamp); }
        v0[k] = t0; v1[k] = t1;
      } else {
        float t0 = v0[k], t1 = v1[k];
        const float y0 = yb ? r0[k] : 0.f, y1 = yb ? r1[k] : 0.f;
        if (act == 3) { t0 = y0 > 0.f ? t0 : t0 * alpha; t1 = y1 > 0.f ? t1 : t1 * alpha; }
        t0 *= gain; t1 *= gain;
        if (clamp >= 0.f) { t0 = (y0 > -clamp && y0 < clamp) ? t0 : 0.f; t1 = (y1 > -clamp && y1 < clamp) ? t1 : 0.f; }
        v0[k] = t0 * pv[k]; v1[k] = t1 * pv[k];
      }
    }
    store8(ob + px * C, v0);
    if (two) store8(ob + px2 * C, v1);
  }
}

// out[n,c] += sum_p a*b with every lane busy: 256 threads = 256/cv pixel rows x cv channel vectors (the first version
// mapped 32 lanes to vectors and left 24 of them idle at C = 64).
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
chan_dot_slab_kernel(const TA* __restrict__ a, const TB* __restrict__ b, float* __restrict__ out, int64_t hw, int C, int slab) {
  const int cv = C / 8, pp = 256 / cv;
  const int lc = threadIdx.x % cv, prow = threadIdx.x / cv;
  const int n = blockIdx.y, c = lc * 8;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * slab, p1 = min(hw, p0 + slab);
  const TA* ab = a + static_cast<int64_t>(n) * hw * C + c;
  const TB* bb = b + static_cast<int64_t>(n) * hw * C + c;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t px = p0 + prow; px < p1; px += 2 * pp) {
    const int64_t px2 = px + pp;
    float a0[8], b0[8], a1[8], b1[8];
    load8(ab + px * C, a0);
    load8(bb + px * C, b0);
    if (px2 < p1) { load8(ab + px2 * C, a1); load8(bb + px2 * C, b1); }
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) { a1[k] = 0.f; b1[k] = 0.f; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(a1[k], b1[k], fmaf(a0[k], b0[k], acc[k]));
  }
  __shared__ float red[256];
  for (int k = 0; k < 8; ++k) {
    __syncthreads();
    red[threadIdx.x] = acc[k];
    __syncthreads();
    if (prow == 0) {
      float t = 0.f;
      for (int r = 0; r < pp; ++r) t += red[r * cv + lc];
      atomicAdd(out + static_cast<int64_t>(n) * C + c + k, t);
    }
  }
}

// First-order backward of y = clamp(act(x*pre[n,c] + noise[n,p] + bias[c]) * gain) in ONE pass:
//   t = dy * gain * act'(y) * [|y| < clamp];  dx = t * pre;  dpre[n,c] += sum_p t*x;  dbias_n[n,c] += sum_p t;
//   dnoise[n,p] = sum_c t
// (the composed form is five passes: activation gradient, modulate, chan_dot and two reductions).
// grid = (pixel slabs, N); 256 threads = (256/cv pixels) x (cv = C/8 channel vectors); needs 256 % cv == 0.
template <typename T>
__global__ void __launch_bounds__(256)
mod_bias_act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                        const float* __restrict__ pre, T* __restrict__ dx, float* __restrict__ dpre,
                        float* __restrict__ dbias_n, float* __restrict__ dnoise, int64_t hw, int C, int slab, int act,
                        float alpha, float gain, float clamp) {
  const int cv = C / 8, pp = 256 / cv;
  const int lc = threadIdx.x % cv, prow = threadIdx.x / cv;
  const int n = blockIdx.y;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * slab, p1 = min(hw, p0 + slab);
  const int c = lc * 8;
  float pv[8], apre[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { pv[k] = pre ? pre[static_cast<int64_t>(n) * C + c + k] : 1.f; apre[k] = 0.f; ab[k] = 0.f; }
  const int red_w = cv < 32 ? cv : 32;  // lanes of a warp that share a pixel
  for (int64_t pb = p0; pb < p1; pb += pp) {  // uniform trip count: the shuffles below need every lane of the warp
    const int64_t px = pb + prow;
    const bool live = px < p1;
    const int64_t o = (static_cast<int64_t>(n) * hw + (live ? px : p0)) * C + c;
    float g[8], yr[8], xv[8];
    Pack<T>::load(dy + o, g);
    Pack<T>::load(y + o, yr);
    if (x) Pack<T>::load(x + o, xv);
    if (!live) {
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = 0.f;
    }
    float s = 0.f, d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = g[k];
      if (act == 3) t = yr[k] > 0.f ? t : t * alpha;
      t *= gain;
      if (clamp >= 0.f) t = (yr[k] > -clamp && yr[k] < clamp) ? t : 0.f;
      s += t;
      ab[k] += t;
      if (x) apre[k] = fmaf(t, xv[k], apre[k]);
      d[k] = t * pv[k];
    }
    if (live) Pack<T>::store(dx + o, d);
    if (dnoise) {
      for (int off = red_w >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (live && (threadIdx.x & (red_w - 1)) == 0) {
        if (cv <= 32) dnoise[static_cast<int64_t>(n) * hw + px] = s;
        else atomicAdd(dnoise + static_cast<int64_t>(n) * hw + px, s);
      }
    }
  }
  // reduce the per-thread channel sums over the pixel rows of the block
  __shared__ float red[2][256 * 8 / 2 + 8];  // [which][ (prow, lc, k) ] folded: processed in two halves below
  float* r0 = &red[0][0];
  for (int which = 0; which < 2; ++which) {
    float* acc = which == 0 ? apre : ab;
    float* dst = which == 0 ? dpre : dbias_n;
    if (!dst) continue;
    // tree over prow using shared memory in 8-channel slices
    for (int k = 0; k < 8; ++k) {
      __syncthreads();
      r0[threadIdx.x] = acc[k];
      __syncthreads();
      if (prow == 0) {
        float t = 0.f;
        for (int r = 0; r < pp; ++r) t += r0[r * cv + lc];
        atomicAdd(dst + static_cast<int64_t>(n) * C + c + k, t);
      }
    }
  }
}

}  // namespace
}  // namespace icgan

using namespace icgan;
#define STREAM static_cast<cudaStream_t>(stream)

static int grid_for(int64_t work) {
  int64_t b = (work + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

template <typename T, int UP, int DOWN, int TOH, int TOW, bool SEP>
static int launch_upfirdn_tiled(const void* x, const float* f, void* y, const UpfirdnTiledParams& p, cudaStream_t st) {
  constexpr int NV = 16 / sizeof(T), CB = 8 * NV;
  constexpr int PH = (TOH * DOWN + 4 - 1 + UP - 1) / UP + 1, PW = (TOW * DOWN + 4 - 1 + UP - 1) / UP + 1;
  constexpr size_t smem = static_cast<size_t>(PH) * PW * 8 * 16;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(upfirdn2d_tiled_kernel<T, UP, DOWN, TOH, TOW, SEP>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const int64_t blocks = static_cast<int64_t>(p.N) * ((p.outH + TOH - 1) / TOH) * ((p.outW + TOW - 1) / TOW) *
                         ((p.C + CB - 1) / CB);
  ICGAN_REQUIRE(blocks > 0 && blocks < (1ll << 31), "icgan_upfirdn2d_nhwc: grid too large");
  upfirdn2d_tiled_kernel<T, UP, DOWN, TOH, TOW, SEP><<<static_cast<unsigned>(blocks), 256, smem, st>>>(
      static_cast<const T*>(x), f, static_cast<T*>(y), p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

template <typename T, int UP, int DOWN, int ROWS>
static int launch_upfirdn_direct(const void* x, void* y, const UpfirdnTiledParams& p, cudaStream_t st) {
  constexpr int NV = 16 / sizeof(T), CB = 8 * NV;
  const int64_t blocks = static_cast<int64_t>(p.N) * ((p.outH + ROWS - 1) / ROWS) * ((p.outW + 31) / 32) * ((p.C + CB - 1) / CB);
  ICGAN_REQUIRE(blocks > 0 && blocks < (1ll << 31), "icgan_upfirdn2d_nhwc: grid too large");
  upfirdn2d_direct_kernel<T, UP, DOWN, ROWS><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(y), p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int dispatch_upfirdn_tiled(const void* x, const float* f, void* y, const UpfirdnTiledParams& p, int up, int down,
                                  cudaStream_t st) {
  static const bool use_direct = []() { const char* e = getenv("ICGAN_FIR_DIRECT"); return e ? atoi(e) != 0 : true; }();
  if (p.separable && use_direct) {
    if (up == 1 && down == 1) return launch_upfirdn_direct<T, 1, 1, 32>(x, y, p, st);
    if (up == 2 && down == 1) return launch_upfirdn_direct<T, 2, 1, 32>(x, y, p, st);
    if (up == 1 && down == 2) return launch_upfirdn_direct<T, 1, 2, 16>(x, y, p, st);
  }
  if (p.separable) {
    if (up == 1 && down == 1) return launch_upfirdn_tiled<T, 1, 1, 8, 32, true>(x, f, y, p, st);
    if (up == 2 && down == 1) return launch_upfirdn_tiled<T, 2, 1, 8, 32, true>(x, f, y, p, st);
    if (up == 1 && down == 2) return launch_upfirdn_tiled<T, 1, 2, 8, 16, true>(x, f, y, p, st);
  } else {
    if (up == 1 && down == 1) return launch_upfirdn_tiled<T, 1, 1, 8, 32, false>(x, f, y, p, st);
    if (up == 2 && down == 1) return launch_upfirdn_tiled<T, 2, 1, 8, 32, false>(x, f, y, p, st);
    if (up == 1 && down == 2) return launch_upfirdn_tiled<T, 1, 2, 8, 16, false>(x, f, y, p, st);
  }
  icgan::set_error("icgan_upfirdn2d_nhwc: (up, down) must be (1,1), (2,1) or (1,2)");
  return -1;
}

extern "C" int icgan_upfirdn2d_nhwc(const void* x, const float* f4x4, void* y, int N, int C, int inH, int inW, int up,
                                    int down, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                                    const float* pre_scale, const float* noise, const float* noise_strength,
                                    int noise_per_sample, const float* bias, int act, float alpha, float act_gain,
                                    float clamp, const float* s2, void* y2, const float* fx_host, const float* fy_host,
                                    int dtype, void* stream) {
  ICGAN_REQUIRE(x && f4x4 && y && N > 0 && C > 0, "icgan_upfirdn2d_nhwc: bad arguments");
  ICGAN_REQUIRE((fx_host == nullptr) == (fy_host == nullptr), "icgan_upfirdn2d_nhwc: fx_host and fy_host go together");
  ICGAN_REQUIRE(act == 0 || act == 1 || act == 3, "icgan_upfirdn2d_nhwc: epilogue activation must be 0 (none), 1 or 3");
  ICGAN_REQUIRE((s2 == nullptr) == (y2 == nullptr), "icgan_upfirdn2d_nhwc: s2 and y2 go together");
  UpfirdnTiledParams p{};
  p.N = N; p.C = C; p.inH = inH; p.inW = inW;
  p.outW = (inW * up + padx0 + padx1 - 4 + down) / down;  // upfirdn2d.cpp:35-36
  p.outH = (inH * up + pady0 + pady1 - 4 + down) / down;
  ICGAN_REQUIRE(p.outW >= 1 && p.outH >= 1, "icgan_upfirdn2d_nhwc: empty output");
  p.pad0x = padx0; p.pad0y = pady0; p.flip = flip; p.gain = gain;
  p.separable = fx_host != nullptr;
  for (int t = 0; t < 4; ++t) {  // same orientation and gain as the 2-D table the kernel builds from f4x4
    p.fx[t] = p.separable ? fx_host[flip ? t : 3 - t] : 0.f;
    p.fy[t] = p.separable ? fy_host[flip ? t : 3 - t] * gain : 0.f;
  }
  p.post = PostArgs{pre_scale, noise, noise_strength, bias, s2, y2, noise_per_sample, act, alpha, act_gain, clamp};
  if (!act && (pre_scale || noise || bias)) { icgan::set_error("icgan_upfirdn2d_nhwc: epilogue inputs need act != 0"); return -1; }
  const int nv = dtype == ICGAN_F32 ? 4 : 8;
  ICGAN_REQUIRE(C % nv == 0, "icgan_upfirdn2d_nhwc: C must be a multiple of %d (got %d)", nv, C);
  if (dtype == ICGAN_BF16) return dispatch_upfirdn_tiled<__nv_bfloat16>(x, f4x4, y, p, up, down, STREAM);
  if (dtype == ICGAN_F16) return dispatch_upfirdn_tiled<__half>(x, f4x4, y, p, up, down, STREAM);
  return dispatch_upfirdn_tiled<float>(x, f4x4, y, p, up, down, STREAM);
}

extern "C" int icgan_modulate(const void* x, const float* s, void* y, int N, int64_t hw, int C, int in_dtype, int out_dtype,
                              void* stream) {
  ICGAN_REQUIRE(x && s && y && N > 0 && hw > 0 && C % 8 == 0, "icgan_modulate: bad arguments (C %% 8 == 0 required)");
  const int64_t total8 = static_cast<int64_t>(N) * hw * C / 8;
  const int g = grid_for(total8);
#define ICGAN_MOD(TI, TO) modulate_kernel<TI, TO><<<g, 256, 0, STREAM>>>(static_cast<const TI*>(x), s, static_cast<TO*>(y), hw, C, total8)
  if (in_dtype == ICGAN_BF16 && out_dtype == ICGAN_BF16) ICGAN_MOD(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == ICGAN_F32 && out_dtype == ICGAN_F32) ICGAN_MOD(float, float);
  else if (in_dtype == ICGAN_F32 && out_dtype == ICGAN_BF16) ICGAN_MOD(float, __nv_bfloat16);
  else if (in_dtype == ICGAN_BF16 && out_dtype == ICGAN_F32) ICGAN_MOD(__nv_bfloat16, float);
  else { icgan::set_error("icgan_modulate: dtypes must be float32 / bfloat16"); return -1; }
#undef ICGAN_MOD
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_chan_dot(const void* a, const void* b, float* out, int N, int64_t hw, int C, int a_dtype, int b_dtype,
                              void* stream) {
  ICGAN_REQUIRE(a && b && out && N > 0 && hw > 0 && C % 8 == 0, "icgan_chan_dot: bad arguments (C %% 8 == 0 required)");
  ICGAN_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(N) * C * sizeof(float), STREAM));
  // enough CTAs to fill the machine: slabs of pixels per sample
  int slabs = static_cast<int>((static_cast<int64_t>(num_sms()) * 4 + N - 1) / N);
  if (slabs > hw / 8) slabs = static_cast<int>(hw / 8 > 0 ? hw / 8 : 1);
  if (slabs < 1) slabs = 1;
  const int slab = static_cast<int>((hw + slabs - 1) / slabs);
  dim3 grid(static_cast<unsigned>((hw + slab - 1) / slab), static_cast<unsigned>(N));
  const bool slabbed = C / 8 <= 256 && 256 % (C / 8) == 0;
#define ICGAN_CD(TA, TB)                                                                                                        \
  do {                                                                                                                          \
    if (slabbed) chan_dot_slab_kernel<TA, TB><<<grid, 256, 0, STREAM>>>(static_cast<const TA*>(a), static_cast<const TB*>(b), out, hw, C, slab); \
    else chan_dot_kernel<TA, TB><<<grid, 256, 0, STREAM>>>(static_cast<const TA*>(a), static_cast<const TB*>(b), out, hw, C, slab);             \
  } while (0)
  if (a_dtype == ICGAN_BF16 && b_dtype == ICGAN_BF16) ICGAN_CD(__nv_bfloat16, __nv_bfloat16);
  else if (a_dtype == ICGAN_F32 && b_dtype == ICGAN_F32) ICGAN_CD(float, float);
  else if (a_dtype == ICGAN_F32 && b_dtype == ICGAN_BF16) ICGAN_CD(float, __nv_bfloat16);
  else if (a_dtype == ICGAN_BF16 && b_dtype == ICGAN_F32) ICGAN_CD(__nv_bfloat16, float);
  else { icgan::set_error("icgan_chan_dot: dtypes must be float32 / bfloat16"); return -1; }
#undef ICGAN_CD
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_mod_bias_act_bwd(const void* dy, const void* y, const void* x, const float* pre_scale, void* dx,
                                      float* dpre, float* dbias_n, float* dnoise, int N, int64_t hw, int C, int act,
                                      float alpha, float gain, float clamp, int dtype, void* stream) {
  ICGAN_REQUIRE(dy && y && dx && N > 0 && hw > 0, "icgan_mod_bias_act_bwd: bad arguments");
  ICGAN_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "icgan_mod_bias_act_bwd: C/8 must divide 256 (got C=%d)", C);
  ICGAN_REQUIRE(act == 1 || act == 3, "icgan_mod_bias_act_bwd: linear / lrelu only");
  ICGAN_REQUIRE(!dpre || x, "icgan_mod_bias_act_bwd: dpre needs x");
  if (dpre) ICGAN_CUDA(cudaMemsetAsync(dpre, 0, static_cast<size_t>(N) * C * 4, STREAM));
  if (dbias_n) ICGAN_CUDA(cudaMemsetAsync(dbias_n, 0, static_cast<size_t>(N) * C * 4, STREAM));
  if (dnoise && C / 8 > 32) ICGAN_CUDA(cudaMemsetAsync(dnoise, 0, static_cast<size_t>(N) * hw * 4, STREAM));
  const int pp = 256 / (C / 8);
  int slabs = static_cast<int>((static_cast<int64_t>(num_sms()) * 6 + N - 1) / N);
  const int64_t max_slabs = (hw + pp - 1) / pp;
  if (slabs > max_slabs) slabs = static_cast<int>(max_slabs);
  if (slabs < 1) slabs = 1;
  int slab = static_cast<int>((hw + slabs - 1) / slabs);
  slab = (slab + pp - 1) / pp * pp;
  dim3 grid(static_cast<unsigned>((hw + slab - 1) / slab), static_cast<unsigned>(N));
#define ICGAN_MB(T) mod_bias_act_bwd_kernel<T><<<grid, 256, 0, STREAM>>>(static_cast<const T*>(dy), static_cast<const T*>(y), static_cast<const T*>(x), pre_scale, static_cast<T*>(dx), dpre, dbias_n, dnoise, hw, C, slab, act, alpha, gain, clamp)
  if (dtype == ICGAN_BF16) ICGAN_MB(__nv_bfloat16);
  else { icgan::set_error("icgan_mod_bias_act_bwd: bfloat16 activations only"); return -1; }
#undef ICGAN_MB
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_bias_act_nhwc(const void* x, const void* yref, void* y, const float* bias, const float* pre_scale,
                                   const float* noise, const float* noise_strength, int noise_per_sample, int N,
                                   int64_t hw, int C, int grad, int act, float alpha, float gain, float clamp, int dtype,
                                   void* stream) {
  ICGAN_REQUIRE(x && y && N > 0 && hw > 0 && C % 8 == 0, "icgan_bias_act_nhwc: bad arguments (C %% 8 == 0 required)");
  ICGAN_REQUIRE((act == 1 || act == 3) && (grad == 0 || grad == 1), "icgan_bias_act_nhwc: linear/lrelu, grad 0/1 only");
  ICGAN_REQUIRE(grad == 0 || yref || (act == 1 && clamp < 0.f), "icgan_bias_act_nhwc: grad 1 needs the forward output");
  const int64_t total8 = static_cast<int64_t>(N) * hw * C / 8;
  const int g = grid_for(total8);
  const int cvv = C / 8;
  const bool slabbed = cvv <= 256 && 256 % cvv == 0 && hw >= 64;
  const int pp = slabbed ? 256 / cvv : 1;
  int slabs = static_cast<int>((static_cast<int64_t>(num_sms()) * 8 + N - 1) / N);
  const int64_t max_slabs = (hw + 2 * pp - 1) / (2 * pp);
  if (slabs > max_slabs) slabs = static_cast<int>(max_slabs);
  if (slabs < 1) slabs = 1;
  const int slab = static_cast<int>((hw + slabs - 1) / slabs);
  const dim3 sgrid(static_cast<unsigned>((hw + slab - 1) / slab), static_cast<unsigned>(N));
#define ICGAN_BA(T)                                                                                                             \
  do {                                                                                                                          \
    if (slabbed)                                                                                                                \
      bias_act_slab_kernel<T><<<sgrid, 256, 0, STREAM>>>(static_cast<const T*>(x), static_cast<const T*>(yref), static_cast<T*>(y), bias, pre_scale, noise, noise_strength, noise_per_sample, hw, C, slab, grad, act, alpha, gain, clamp); \
    else                                                                                                                        \
      bias_act_vec_kernel<T><<<g, 256, 0, STREAM>>>(static_cast<const T*>(x), static_cast<const T*>(yref), static_cast<T*>(y), bias, pre_scale, noise, noise_strength, noise_per_sample, hw, C, total8, grad, act, alpha, gain, clamp); \
  } while (0)
  if (dtype == ICGAN_BF16) ICGAN_BA(__nv_bfloat16);
  else if (dtype == ICGAN_F32) ICGAN_BA(float);
  else { icgan::set_error("icgan_bias_act_nhwc: dtype must be float32 / bfloat16"); return -1; }
#undef ICGAN_BA
  ICGAN_LAUNCH_CHECK();
  return 0;
}
