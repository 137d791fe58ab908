// CUDA-core convolution kernels (float32 accumulate), NHWC.  These are the exact-fp32 "parity mode" path and the
// fallback for shapes the tensor-core kernels do not take (Cin=3 / Cout=3 image-side layers, 4x4 feature maps in
// wgrad, strided StyleGAN2 convs).  Reference: F.conv2d behind layers.SNConv2d.forward (BigGAN_PyTorch/layers.py:144-153)
// and conv2d_gradfix (stylegan2_ada_pytorch/torch_utils/ops/conv2d_gradfix.py:126-272).
#include "common.cuh"
#include "norm_act_vec.cuh"

namespace icgan {

constexpr int TP = 64;  // output pixels per block
constexpr int TC = 64;  // output channels per block
constexpr int TK = 16;  // reduction chunk

struct SimtConvParams {
  int B, Hin, Win, Hout, Wout, Cin, Cout, ksz, stride, pad;
  int res_shift, act;
};

template <typename TIn, typename TOut, typename TRes>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const TIn* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ alpha_p,
                 const float* __restrict__ bias, const TRes* __restrict__ res, TOut* __restrict__ y, SimtConvParams p) {
  const float alpha = alpha_p ? *alpha_p : 1.f;
  __shared__ float As[TK][TP + 4];
  __shared__ float Bs[TK][TC + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> channels, ty -> pixels
  const int64_t P = static_cast<int64_t>(p.B) * p.Hout * p.Wout;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int co0 = blockIdx.y * TC;
  const int taps = p.ksz * p.ksz;

  // loader assignment: 64 rows x 16 k  -> thread loads row (tid / 4), k-quad (tid % 4) * 4
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const int64_t lp = p0 + lrow;
  int ln = 0, lh = 0, lw = 0;
  const bool lvalid = lp < P;
  if (lvalid) {
    lw = static_cast<int>(lp % p.Wout);
    const int64_t t = lp / p.Wout;
    lh = static_cast<int>(t % p.Hout);
    ln = static_cast<int>(t / p.Hout);
  }
  const int lco = co0 + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps; ++tap) {
    const int kh = tap / p.ksz, kw = tap % p.ksz;
    const int ih = lh * p.stride + kh - p.pad, iw = lw * p.stride + kw - p.pad;
    const bool in_ok = lvalid && ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win;
    const int64_t xoff = ((static_cast<int64_t>(ln) * p.Hin + ih) * p.Win + iw) * p.Cin;
    const int64_t woff = (static_cast<int64_t>(lco) * taps + tap) * p.Cin;
    for (int c0 = 0; c0 < p.Cin; c0 += TK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + lk + j;
        As[lk + j][lrow] = (in_ok && c < p.Cin) ? ld_as_float(x, xoff + c) : 0.f;
        Bs[lk + j][lrow] = (lco < p.Cout && c < p.Cin) ? wk[woff + c] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pp = p0 + ty * 4 + i;
    if (pp >= P) continue;
    int64_t rp = pp;
    if (res != nullptr && p.res_shift) {
      const int w = static_cast<int>(pp % p.Wout);
      const int64_t t = pp / p.Wout;
      const int h = static_cast<int>(t % p.Hout);
      const int n = static_cast<int>(t / p.Hout);
      rp = (static_cast<int64_t>(n) * (p.Hout >> 1) + (h >> 1)) * (p.Wout >> 1) + (w >> 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= p.Cout) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[co];
      if (res) v += ld_as_float(res, rp * p.Cout + co);
      if (p.act == ICGAN_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == ICGAN_ACT_TANH) v = tanhf(v);
      st_from_float(y, pp * p.Cout + co, v);
    }
  }
}

// dwk[co][tap][ci] += sum over a slab of output pixels of dy[p][co] * x[p (+) tap][ci]
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dwk, SimtConvParams p,
                  int ci_tiles, int64_t pix_per_split) {
  __shared__ float As[TK][TC + 4];  // dy  [pixel][co]
  __shared__ float Bs[TK][TC + 4];  // x   [pixel][ci]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> ci, ty -> co
  const int co0 = (blockIdx.x / ci_tiles) * TC, ci0 = (blockIdx.x % ci_tiles) * TC;
  const int tap = blockIdx.y;
  const int kh = tap / p.ksz, kw = tap % p.ksz;
  const int taps = p.ksz * p.ksz;
  const int64_t P = static_cast<int64_t>(p.B) * p.Hout * p.Wout;
  const int64_t pbeg = static_cast<int64_t>(blockIdx.z) * pix_per_split;
  const int64_t pend = pbeg + pix_per_split < P ? pbeg + pix_per_split : P;

  // loader: 16 pixels x 64 channels -> thread loads pixel (tid / 16), channels (tid % 16) * 4 .. +3
  const int lpix = tid / 16, lc = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t pc = pbeg; pc < pend; pc += TK) {
    const int64_t pp = pc + lpix;
    const bool pv = pp < pend;
    int n = 0, h = 0, w = 0;
    if (pv) {
      w = static_cast<int>(pp % p.Wout);
      const int64_t t = pp / p.Wout;
      h = static_cast<int>(t % p.Hout);
      n = static_cast<int>(t / p.Hout);
    }
    const int ih = h * p.stride + kh - p.pad, iw = w * p.stride + kw - p.pad;
    const bool xin = pv && ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win;
    const int64_t xoff = ((static_cast<int64_t>(n) * p.Hin + ih) * p.Win + iw) * p.Cin;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + lc + j, ci = ci0 + lc + j;
      As[lpix][lc + j] = (pv && co < p.Cout) ? ld_as_float(dy, pp * p.Cout + co) : 0.f;
      Bs[lpix][lc + j] = (xin && ci < p.Cin) ? ld_as_float(x, xoff + ci) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < p.Cin) atomicAdd(dwk + (static_cast<int64_t>(co) * taps + tap) * p.Cin + ci, acc[i][j]);
    }
  }
}

// out[c] += sum_p x[p][c]   (bias gradients; NHWC column sums)
template <typename T>
__global__ void channel_sum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t P, int C,
                                   int64_t rows_per_block) {
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < P ? r0 + rows_per_block : P;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < C; c += gridDim.y * blockDim.x) {
    float s = 0.f;
    int64_t r = r0;
    for (; r + 4 <= r1; r += 4) {
      const float a = ld_as_float(x, r * C + c), b = ld_as_float(x, (r + 1) * C + c),
                  d = ld_as_float(x, (r + 2) * C + c), e = ld_as_float(x, (r + 3) * C + c);
      s += (a + b) + (d + e);
    }
    for (; r < r1; ++r) s += ld_as_float(x, r * C + c);
    atomicAdd(out + c, s);
  }
}

template <typename TIn, typename TOut, typename TRes>
static int launch_conv_simt(const void* x, const float* wk, const float* alpha, const float* bias, const void* res,
                            void* y, const SimtConvParams& p, cudaStream_t s) {
  const int64_t P = static_cast<int64_t>(p.B) * p.Hout * p.Wout;
  dim3 grid(static_cast<unsigned>((P + TP - 1) / TP), static_cast<unsigned>((p.Cout + TC - 1) / TC));
  conv_simt_kernel<TIn, TOut, TRes><<<grid, 256, 0, s>>>(static_cast<const TIn*>(x), wk, alpha, bias,
                                                         static_cast<const TRes*>(res), static_cast<TOut*>(y), p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

}  // namespace icgan

using namespace icgan;

extern "C" int icgan_conv2d_simt(const void* x, const float* wk, const float* alpha_dev, const float* bias,
                                 const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize,
                                 int stride, int pad, int in_dtype, int out_dtype, int res_dtype, int res_shift, int act,
                                 void* stream) {
  ICGAN_REQUIRE(x && wk && y, "icgan_conv2d_simt: null pointer");
  ICGAN_REQUIRE(ksize >= 1 && ksize <= 7 && stride >= 1 && pad >= 0, "icgan_conv2d_simt: bad ksize/stride/pad");
  SimtConvParams p{};
  p.B = B; p.Hin = H; p.Win = W; p.Cin = Cin; p.Cout = Cout; p.ksz = ksize; p.stride = stride; p.pad = pad;
  p.Hout = (H + 2 * pad - ksize) / stride + 1;
  p.Wout = (W + 2 * pad - ksize) / stride + 1;
  ICGAN_REQUIRE(p.Hout > 0 && p.Wout > 0, "icgan_conv2d_simt: empty output");
  p.res_shift = res_shift; p.act = act;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ib = in_dtype == ICGAN_BF16, ob = out_dtype == ICGAN_BF16, rb = res_dtype == ICGAN_BF16;
  typedef __nv_bfloat16 bf;
  if (!ib && !ob && !rb) return launch_conv_simt<float, float, float>(x, wk, alpha_dev, bias, residual, y, p, s);
  if (ib && ob && rb) return launch_conv_simt<bf, bf, bf>(x, wk, alpha_dev, bias, residual, y, p, s);
  if (ib && !ob && !rb) return launch_conv_simt<bf, float, float>(x, wk, alpha_dev, bias, residual, y, p, s);
  if (ib && !ob && rb) return launch_conv_simt<bf, float, bf>(x, wk, alpha_dev, bias, residual, y, p, s);
  if (ib && ob && !rb) return launch_conv_simt<bf, bf, float>(x, wk, alpha_dev, bias, residual, y, p, s);
  if (!ib && ob && !rb) return launch_conv_simt<float, bf, float>(x, wk, alpha_dev, bias, residual, y, p, s);
  if (!ib && ob && rb) return launch_conv_simt<float, bf, bf>(x, wk, alpha_dev, bias, residual, y, p, s);
  return launch_conv_simt<float, float, bf>(x, wk, alpha_dev, bias, residual, y, p, s);
}

extern "C" int icgan_conv2d_wgrad_simt(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin,
                                       int Cout, int ksize, int stride, int pad, int in_dtype, void* stream) {
  ICGAN_REQUIRE(x && dy && dwk, "icgan_conv2d_wgrad_simt: null pointer");
  SimtConvParams p{};
  p.B = B; p.Hin = H; p.Win = W; p.Cin = Cin; p.Cout = Cout; p.ksz = ksize; p.stride = stride; p.pad = pad;
  p.Hout = (H + 2 * pad - ksize) / stride + 1;
  p.Wout = (W + 2 * pad - ksize) / stride + 1;
  const int64_t P = static_cast<int64_t>(B) * p.Hout * p.Wout;
  const int co_tiles = ceil_div(Cout, TC), ci_tiles = ceil_div(Cin, TC);
  const int taps = ksize * ksize;
  int splits = ceil_div(4 * num_sms(), static_cast<int64_t>(co_tiles) * ci_tiles * taps);
  const int64_t max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = static_cast<int>(max_splits);
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t per = (P + splits - 1) / splits;
  per = (per + TK - 1) / TK * TK;
  splits = static_cast<int>((P + per - 1) / per);
  dim3 grid(static_cast<unsigned>(co_tiles * ci_tiles), static_cast<unsigned>(taps), static_cast<unsigned>(splits));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_dtype == ICGAN_BF16)
    wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x),
                                                         static_cast<const __nv_bfloat16*>(dy), dwk, p, ci_tiles, per);
  else
    wgrad_simt_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), static_cast<const float*>(dy), dwk, p,
                                                 ci_tiles, per);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_channel_sum(const void* x, float* out, int64_t P, int C, int dtype, void* stream) {
  ICGAN_REQUIRE(x && out && P > 0 && C > 0, "icgan_channel_sum: bad arguments");
  if (dtype == ICGAN_BF16 && vec::ok(C)) {
    int64_t blocks = static_cast<int64_t>(num_sms()) * 8;
    if (blocks > (P + 127) / 128) blocks = (P + 127) / 128;
    const int64_t ppb = (P + blocks - 1) / blocks;
    blocks = (P + ppb - 1) / ppb;
    vec::colsum_vec_kernel<0><<<static_cast<unsigned>(blocks), vec::kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const vec::bf16*>(x), out, P, C, ppb);
    ICGAN_LAUNCH_CHECK();
    return 0;
  }
  int row_blocks = static_cast<int>((P + 63) / 64);
  if (row_blocks > 24 * num_sms()) row_blocks = 24 * num_sms();
  const int64_t rpb = (P + row_blocks - 1) / row_blocks;
  row_blocks = static_cast<int>((P + rpb - 1) / rpb);
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : (C >= 64 ? 64 : 32));
  dim3 grid(static_cast<unsigned>(row_blocks), static_cast<unsigned>((C + threads - 1) / threads));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == ICGAN_BF16)
    channel_sum_kernel<__nv_bfloat16><<<grid, threads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), out, P, C, rpb);
  else
    channel_sum_kernel<float><<<grid, threads, 0, s>>>(static_cast<const float*>(x), out, P, C, rpb);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
