// Image-side convolutions: one of Cin/Cout is tiny (RGB = 3).  These layers (D.blocks.0 conv1 3->ch and its 1x1
// shortcut, G.output_layer 96->3; BigGAN.py:283-291, :493-530) carry <0.1% of the FLOPs but touch full-resolution
// activations, so they are HBM-bound streaming kernels, not GEMMs: each output element needs <= 27 MACs per small
// channel.  The generic 64x64-tile CUDA-core kernel wastes 95% of its tile on them (27% of the step in
// profiles/launches_r01); these kernels keep threads along the wide channel dimension (coalesced 16-byte accesses)
// and the small dimension in registers.
#include "common.cuh"

namespace icgan {

constexpr int kMaxSmall = 4;

struct SmallConvParams {
  int B, H, W, Cin, Cout, ksz, pad, act;
};

// ---- Cin small: y[p][co] = act(sum_{tap,ci<CS} x[p+tap][ci] * w[co][tap][ci] + bias[co]); thread = 8 out channels
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
conv_small_cin_kernel(const TI* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ alpha_p,
                      const float* __restrict__ bias, TO* __restrict__ y, SmallConvParams p) {
  const float alpha = alpha_p ? *alpha_p : 1.f;
  extern __shared__ float wsm[];  // [tap][ci][j<8][group]: lanes with consecutive groups hit consecutive banks
  const int taps = p.ksz * p.ksz;
  const int groups = (p.Cout + 7) / 8;
  for (int i = threadIdx.x; i < groups * 8 * taps * p.Cin; i += blockDim.x) {
    const int ci = i % p.Cin, tap = (i / p.Cin) % taps, co = i / (p.Cin * taps);
    wsm[((tap * p.Cin + ci) * 8 + (co & 7)) * groups + (co >> 3)] = co < p.Cout ? wk[i] : 0.f;
  }
  __syncthreads();
  const int64_t total = static_cast<int64_t>(p.B) * p.H * p.W * groups;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(t % groups);
    const int64_t pix = t / groups;
    const int w = static_cast<int>(pix % p.W);
    const int h = static_cast<int>((pix / p.W) % p.H);
    const int64_t n = pix / (static_cast<int64_t>(p.H) * p.W);
    const int co0 = g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < taps; ++tap) {
      const int ih = h + tap / p.ksz - p.pad, iw = w + tap % p.ksz - p.pad;
      if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) continue;
      const int64_t xo = ((n * p.H + ih) * p.W + iw) * p.Cin;
      for (int ci = 0; ci < p.Cin; ++ci) {
        const float xv = ld_as_float(x, xo + ci);
        const float* wr = wsm + (tap * p.Cin + ci) * 8 * groups + g;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wr[j * groups], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (co0 + j >= p.Cout) break;
      float v = alpha * acc[j] + (bias ? bias[co0 + j] : 0.f);
      if (p.act == ICGAN_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == ICGAN_ACT_TANH) v = tanhf(v);
      st_from_float(y, pix * p.Cout + co0 + j, v);
    }
  }
}

// ---- Cout small: one warp per output pixel, lanes stride over Cin, warp-shuffle reduction of the CS sums
template <typename TI, typename TO, int CS>
__global__ void __launch_bounds__(256)
conv_small_cout_kernel(const TI* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ alpha_p,
                       const float* __restrict__ bias, TO* __restrict__ y, SmallConvParams p) {
  const float alpha = alpha_p ? *alpha_p : 1.f;
  extern __shared__ float wsm[];  // [co][tap][ci] (as given)
  const int taps = p.ksz * p.ksz;
  for (int i = threadIdx.x; i < CS * taps * p.Cin; i += blockDim.x) wsm[i] = i < p.Cout * taps * p.Cin ? wk[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t P = static_cast<int64_t>(p.B) * p.H * p.W;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t pix = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; pix < P; pix += warps) {
    const int w = static_cast<int>(pix % p.W);
    const int h = static_cast<int>((pix / p.W) % p.H);
    const int64_t n = pix / (static_cast<int64_t>(p.H) * p.W);
    float acc[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) acc[c] = 0.f;
    for (int tap = 0; tap < taps; ++tap) {
      const int ih = h + tap / p.ksz - p.pad, iw = w + tap % p.ksz - p.pad;
      if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) continue;
      const int64_t xo = ((n * p.H + ih) * p.W + iw) * p.Cin;
      for (int ci = lane; ci < p.Cin; ci += 32) {
        const float xv = ld_as_float(x, xo + ci);
#pragma unroll
        for (int c = 0; c < CS; ++c) acc[c] = fmaf(xv, wsm[(c * taps + tap) * p.Cin + ci], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < CS; ++c)
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane < p.Cout) {
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c)
        if (lane == c) v = alpha * acc[c];
      if (bias) v += bias[lane];
      if (p.act == ICGAN_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == ICGAN_ACT_TANH) v = tanhf(v);
      st_from_float(y, pix * p.Cout + lane, v);
    }
  }
}

// ---- weight gradient with one tiny channel dimension:
//   out(cb, tap, cs) += sum_p big[p][cb] * small[p + sgn*tap][cs]
// small_is_x=1 (Cin small):  big = dy, small = x shifted by +tap, dwk[cb][tap][cs]
// small_is_x=0 (Cout small): big = x,  small = dy shifted by -tap, dwk[cs][tap][cb]
template <typename TB, typename TS, int CS>
__global__ void wgrad_small_kernel(const TB* __restrict__ big, const TS* __restrict__ small, float* __restrict__ dwk,
                                   int B, int H, int W, int Cb, int Cs, int ksz, int small_is_x,
                                   int64_t pix_per_block) {
  const int taps = ksz * ksz, pad = ksz / 2;
  const int64_t P = static_cast<int64_t>(B) * H * W;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < P ? p0 + pix_per_block : P;
  const int sgn = small_is_x ? 1 : -1;
  for (int cb = threadIdx.x; cb < Cb; cb += blockDim.x) {
    float acc[9][CS];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < CS; ++c) acc[t][c] = 0.f;
    for (int64_t pix = p0; pix < p1; ++pix) {
      const float bv = ld_as_float(big, pix * Cb + cb);
      const int w = static_cast<int>(pix % W);
      const int h = static_cast<int>((pix / W) % H);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (t >= taps) break;
        const int dh = sgn * (t / ksz - pad), dw = sgn * (t % ksz - pad);
        const int ih = h + dh, iw = w + dw;
        if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
        const int64_t so = (pix + static_cast<int64_t>(dh) * W + dw) * Cs;
#pragma unroll
        for (int c = 0; c < CS; ++c)
          if (c < Cs) acc[t][c] = fmaf(bv, ld_as_float(small, so + c), acc[t][c]);
      }
    }
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < Cs; ++c) {
        float* dst = small_is_x ? dwk + (static_cast<int64_t>(cb) * taps + t) * Cs + c
                                : dwk + (static_cast<int64_t>(c) * taps + t) * Cb + cb;
        atomicAdd(dst, acc[t][c]);
      }
  }
}


// ------------------------------------------------------------------------------------------------ 1x1, 128-bit accesses
// The RGB-side layers of StyleGAN2 (toRGB C->3, fromRGB 3->C; networks.py:451-485, :786-790) are 1x1: pure streaming
// over a full-resolution activation.  The generic kernels above read 2-byte elements (measured ~1 TB/s on 64-channel
// 256x256 tensors); these read / write 16 bytes per lane.
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// y[p][co<CS] = alpha * sum_c x[p][c] * w[co][c] + bias[co];  LP (power of two <= 32) lanes share a pixel, each lane owns
// NCH fixed 8-channel chunks whose CS x 8 weights stay in REGISTERS (the first version re-read 24 weights from shared
// memory per 16-byte activation load and reached 0.22 of the HBM peak)
template <typename TO, int CS, int NCH>
__global__ void __launch_bounds__(256)
conv1x1_to_small_vec_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ alpha_p,
                            const float* __restrict__ bias, TO* __restrict__ y, int64_t P, int C, int Cout, int LP) {
  const float alpha = alpha_p ? *alpha_p : 1.f;
  const int lg = threadIdx.x % LP;
  float wr[NCH][CS][8];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int cc = (lg + j * LP) * 8;
#pragma unroll
    for (int c = 0; c < CS; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) wr[j][c][k] = (c < Cout && cc + k < C) ? wk[c * C + cc + k] : 0.f;
  }
  const int64_t groups = (static_cast<int64_t>(gridDim.x) * blockDim.x) / LP;
  const int64_t iters = (P + groups - 1) / groups;
  const int64_t g0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LP;
  for (int64_t it = 0; it < iters; ++it) {  // uniform trip count (shuffles below)
    const int64_t pix = g0 + it * groups;
    const bool live = pix < P;
    float acc[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) acc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int cc = (lg + j * LP) * 8;
      if (live && cc < C) {
        float v[8];
        unpack8(*reinterpret_cast<const uint4*>(x + pix * C + cc), v);
#pragma unroll
        for (int c = 0; c < CS; ++c)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[c] = fmaf(v[k], wr[j][c][k], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < CS; ++c)
      for (int o = LP >> 1; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (live && lg == 0)
#pragma unroll
      for (int c = 0; c < CS; ++c)
        if (c < Cout) st_from_float(y, pix * Cout + c, alpha * acc[c] + (bias ? bias[c] : 0.f));
  }
}

// y[p][co0..co0+8) = alpha * sum_{ci<Cin<=4} x[p][ci] * w[co][ci] + bias: one 16-byte store per thread
template <typename TI>
__global__ void __launch_bounds__(256)
conv1x1_from_small_vec_kernel(const TI* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ alpha_p,
                              const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int64_t P, int Cin, int Cout) {
  extern __shared__ float wsm[];  // [Cout][Cin]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) wsm[i] = wk[i];
  __syncthreads();
  const float alpha = alpha_p ? *alpha_p : 1.f;
  const int groups = Cout / 8;
  const int64_t total = P * groups;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(t % groups);
    const int64_t pix = t / groups;
    float xv[kMaxSmall];
#pragma unroll
    for (int ci = 0; ci < kMaxSmall; ++ci) xv[ci] = ci < Cin ? ld_as_float(x, pix * Cin + ci) : 0.f;
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o2[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int co = g * 8 + 2 * j + e;
        float a = 0.f;
#pragma unroll
        for (int ci = 0; ci < kMaxSmall; ++ci)
          if (ci < Cin) a = fmaf(xv[ci], wsm[co * Cin + ci], a);
        o2[e] = alpha * a + (bias ? bias[co] : 0.f);
      }
      h[j] = __floats2bfloat162_rn(o2[0], o2[1]);
    }
    *reinterpret_cast<uint4*>(y + pix * Cout + g * 8) = pk;
  }
}

// out(cs, c) += sum_p small[p][cs] * big[p][c]   (1x1 weight gradient with one tiny channel count)
// block = (256 / LV pixel rows) x (LV = C/8 channel vectors); blockIdx.x = pixel slab
template <typename TS, int CS>
__global__ void __launch_bounds__(256)
wgrad1x1_small_vec_kernel(const __nv_bfloat16* __restrict__ big, const TS* __restrict__ small, float* __restrict__ dwk,
                          int64_t P, int Cb, int Cs, int small_is_x, int64_t slab) {
  const int LV = Cb / 8, rows = 256 / LV;
  const int lv = threadIdx.x % LV, prow = threadIdx.x / LV;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * slab, p1 = p0 + slab < P ? p0 + slab : P;
  float acc[CS][8];
#pragma unroll
  for (int c = 0; c < CS; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;
  if (prow < rows)
    for (int64_t pix = p0 + prow; pix < p1; pix += rows) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(big + pix * Cb + lv * 8), v);
#pragma unroll
      for (int c = 0; c < CS; ++c) {
        const float sv = c < Cs ? ld_as_float(small, pix * Cs + c) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[c][k] = fmaf(sv, v[k], acc[c][k]);
      }
    }
  __shared__ float red[256];
  for (int c = 0; c < CS; ++c) {
    if (c >= Cs) break;
    for (int k = 0; k < 8; ++k) {
      __syncthreads();
      red[threadIdx.x] = acc[c][k];
      __syncthreads();
      if (prow == 0 && threadIdx.x < LV) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += red[r * LV + lv];
        const int cb = lv * 8 + k;
        float* dst = small_is_x ? dwk + static_cast<int64_t>(cb) * Cs + c : dwk + static_cast<int64_t>(c) * Cb + cb;
        atomicAdd(dst, t);
      }
    }
  }
}

// out[p][j] = x[p + tap(j)][ci(j)] for j = tap*Cs + ci < k*k*Cs, else 0   (bf16 out; thread = 8 output columns)
template <typename TI>
__global__ void im2col_small_kernel(const TI* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                    int Cs, int ksz, int KP) {
  const int groups = KP >> 3, pad = ksz / 2, valid = ksz * ksz * Cs;
  const int64_t total = static_cast<int64_t>(B) * H * W * groups;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(t % groups);
    const int64_t pix = t / groups;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int64_t n = pix / (static_cast<int64_t>(H) * W);
    uint4 pk;
    __nv_bfloat16* v = reinterpret_cast<__nv_bfloat16*>(&pk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = g * 8 + i;
      float val = 0.f;
      if (j < valid) {
        const int tap = j / Cs, ci = j - tap * Cs;
        const int ih = h + tap / ksz - pad, iw = w + tap % ksz - pad;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = ld_as_float(x, ((n * H + ih) * W + iw) * Cs + ci);
      }
      v[i] = __float2bfloat16_rn(val);
    }
    *reinterpret_cast<uint4*>(out + pix * KP + g * 8) = pk;
  }
}

}  // namespace icgan

using namespace icgan;
#define STREAM static_cast<cudaStream_t>(stream)
#define DISPATCH_T(dt, T, ...)   \
  if ((dt) == ICGAN_BF16) {      \
    using T = __nv_bfloat16;     \
    __VA_ARGS__                  \
  } else {                       \
    using T = float;             \
    __VA_ARGS__                  \
  }

extern "C" int icgan_conv2d_small(const void* x, const float* wk, const float* alpha_dev, const float* bias, void* y,
                                  int B, int H, int W, int Cin, int Cout, int ksize, int in_dtype, int out_dtype, int act,
                                  void* stream) {
  ICGAN_REQUIRE(x && wk && y, "icgan_conv2d_small: null pointer");
  ICGAN_REQUIRE(ksize == 1 || ksize == 3, "icgan_conv2d_small: ksize must be 1 or 3");
  ICGAN_REQUIRE(Cin <= kMaxSmall || Cout <= kMaxSmall, "icgan_conv2d_small: needs Cin<=4 or Cout<=4 (got %d, %d)", Cin,
                Cout);
  SmallConvParams p{B, H, W, Cin, Cout, ksize, ksize / 2, act};
  const int taps = ksize * ksize;
  const int64_t P = static_cast<int64_t>(B) * H * W;
  if (ksize == 1 && act == ICGAN_ACT_NONE && Cout <= kMaxSmall && Cin % 8 == 0 && Cin <= 512 && in_dtype == ICGAN_BF16) {
    int LP = 1;  // C -> RGB, 128-bit loads, weights in registers
    while (LP < 32 && LP * 8 < Cin) LP *= 2;
    int64_t blocks = (P * LP + 255) / 256;
    if (blocks > static_cast<int64_t>(num_sms()) * 16) blocks = static_cast<int64_t>(num_sms()) * 16;
    DISPATCH_T(out_dtype, TO, {
      if (Cin <= LP * 8)
        conv1x1_to_small_vec_kernel<TO, kMaxSmall, 1><<<static_cast<unsigned>(blocks), 256, 0, STREAM>>>(
            static_cast<const __nv_bfloat16*>(x), wk, alpha_dev, bias, static_cast<TO*>(y), P, Cin, Cout, LP);
      else
        conv1x1_to_small_vec_kernel<TO, kMaxSmall, 2><<<static_cast<unsigned>(blocks), 256, 0, STREAM>>>(
            static_cast<const __nv_bfloat16*>(x), wk, alpha_dev, bias, static_cast<TO*>(y), P, Cin, Cout, LP);
    })
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  if (ksize == 1 && act == ICGAN_ACT_NONE && Cin <= kMaxSmall && Cout % 8 == 0 && out_dtype == ICGAN_BF16 &&
      static_cast<size_t>(Cout) * Cin * 4 <= 48 * 1024) {  // RGB -> C, 128-bit stores
    const int64_t total = P * (Cout / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > static_cast<int64_t>(num_sms()) * 32) blocks = static_cast<int64_t>(num_sms()) * 32;
    const size_t smem = sizeof(float) * Cout * Cin;
    DISPATCH_T(in_dtype, TI, {
      conv1x1_from_small_vec_kernel<TI><<<static_cast<unsigned>(blocks), 256, smem, STREAM>>>(
          static_cast<const TI*>(x), wk, alpha_dev, bias, static_cast<__nv_bfloat16*>(y), P, Cin, Cout);
    })
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  if (Cin <= kMaxSmall) {
    const size_t smem = sizeof(float) * static_cast<size_t>((Cout + 7) / 8 * 8) * taps * Cin;
    ICGAN_REQUIRE(smem <= 48 * 1024, "icgan_conv2d_small: weights do not fit shared memory");
    const int64_t total = P * ((Cout + 7) / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > static_cast<int64_t>(num_sms()) * 32) blocks = static_cast<int64_t>(num_sms()) * 32;
    DISPATCH_T(in_dtype, TI, {DISPATCH_T(out_dtype, TO, {
      conv_small_cin_kernel<TI, TO><<<static_cast<unsigned>(blocks), 256, smem, STREAM>>>(
          static_cast<const TI*>(x), wk, alpha_dev, bias, static_cast<TO*>(y), p);
    })})
  } else {
    const size_t smem = sizeof(float) * static_cast<size_t>(kMaxSmall) * taps * Cin;
    ICGAN_REQUIRE(smem <= 48 * 1024, "icgan_conv2d_small: weights do not fit shared memory");
    int64_t blocks = (P + 7) / 8;
    if (blocks > static_cast<int64_t>(num_sms()) * 32) blocks = static_cast<int64_t>(num_sms()) * 32;
    DISPATCH_T(in_dtype, TI, {DISPATCH_T(out_dtype, TO, {
      conv_small_cout_kernel<TI, TO, kMaxSmall><<<static_cast<unsigned>(blocks), 256, smem, STREAM>>>(
          static_cast<const TI*>(x), wk, alpha_dev, bias, static_cast<TO*>(y), p);
    })})
  }
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_conv2d_wgrad_small(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin,
                                        int Cout, int ksize, int x_dtype, int dy_dtype, void* stream) {
  ICGAN_REQUIRE(x && dy && dwk, "icgan_conv2d_wgrad_small: null pointer");
  ICGAN_REQUIRE(ksize == 1 || ksize == 3, "icgan_conv2d_wgrad_small: ksize must be 1 or 3");
  ICGAN_REQUIRE(Cin <= kMaxSmall || Cout <= kMaxSmall, "icgan_conv2d_wgrad_small: needs Cin<=4 or Cout<=4");
  const int64_t P = static_cast<int64_t>(B) * H * W;
  const int small_is_x = Cin <= kMaxSmall ? 1 : 0;
  const int Cb = small_is_x ? Cout : Cin, Cs = small_is_x ? Cin : Cout;
  const int big_dtype = small_is_x ? dy_dtype : x_dtype, small_dtype = small_is_x ? x_dtype : dy_dtype;
  if (ksize == 1 && big_dtype == ICGAN_BF16 && Cb % 8 == 0 && Cb / 8 <= 256) {  // 1x1: 128-bit loads of the wide operand
    int64_t blocks = static_cast<int64_t>(num_sms()) * 8;
    const int rows = 256 / (Cb / 8);
    if (blocks > (P + rows - 1) / rows) blocks = (P + rows - 1) / rows;
    if (blocks < 1) blocks = 1;
    const int64_t slab = (P + blocks - 1) / blocks;
    blocks = (P + slab - 1) / slab;
    const void* bigp = small_is_x ? dy : x;
    const void* smallp = small_is_x ? x : dy;
    DISPATCH_T(small_dtype, TS, {
      wgrad1x1_small_vec_kernel<TS, kMaxSmall><<<static_cast<unsigned>(blocks), 256, 0, STREAM>>>(
          static_cast<const __nv_bfloat16*>(bigp), static_cast<const TS*>(smallp), dwk, P, Cb, Cs, small_is_x, slab);
    })
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  const int threads = Cb >= 256 ? 256 : ((Cb + 31) / 32) * 32;
  int64_t blocks = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > (P + 63) / 64) blocks = (P + 63) / 64;
  if (blocks < 1) blocks = 1;
  const int64_t ppb = (P + blocks - 1) / blocks;
  blocks = (P + ppb - 1) / ppb;
  if (small_is_x) {
    DISPATCH_T(dy_dtype, TB, {DISPATCH_T(x_dtype, TS, {
      wgrad_small_kernel<TB, TS, kMaxSmall><<<static_cast<unsigned>(blocks), threads, 0, STREAM>>>(
          static_cast<const TB*>(dy), static_cast<const TS*>(x), dwk, B, H, W, Cb, Cs, ksize, 1, ppb);
    })})
  } else {
    DISPATCH_T(x_dtype, TB, {DISPATCH_T(dy_dtype, TS, {
      wgrad_small_kernel<TB, TS, kMaxSmall><<<static_cast<unsigned>(blocks), threads, 0, STREAM>>>(
          static_cast<const TB*>(x), static_cast<const TS*>(dy), dwk, B, H, W, Cb, Cs, ksize, 0, ppb);
    })})
  }
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_im2col_small(const void* x, void* out, int B, int H, int W, int Cs, int ksize, int KP, int in_dtype,
                                  void* stream) {
  ICGAN_REQUIRE(x && out && Cs >= 1 && Cs <= kMaxSmall && KP % 8 == 0 && KP >= ksize * ksize * Cs,
                "icgan_im2col_small: bad arguments (Cs=%d KP=%d)", Cs, KP);
  const int64_t total = static_cast<int64_t>(B) * H * W * (KP / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > static_cast<int64_t>(num_sms()) * 32) blocks = static_cast<int64_t>(num_sms()) * 32;
  DISPATCH_T(in_dtype, TI, {
    im2col_small_kernel<TI><<<static_cast<unsigned>(blocks), 256, 0, STREAM>>>(
        static_cast<const TI*>(x), static_cast<__nv_bfloat16*>(out), B, H, W, Cs, ksize, KP);
  })
  ICGAN_LAUNCH_CHECK();
  return 0;
}
