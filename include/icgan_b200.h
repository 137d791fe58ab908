/*
 * icgan_b200.h — C ABI of libicgan_b200.so, the B200 (sm_100a) kernels behind IC-GAN's G/D hot path.
 *
 * Conventions:
 *   - every entry point returns int: 0 = ok, <0 = invalid argument, >0 = cudaError_t; text via icgan_last_error();
 *   - all buffers (inputs, outputs, workspaces) are owned by the caller and are DEVICE pointers unless the name
 *     ends in _host; the library never allocates or frees device memory and keeps no pointer past return;
 *   - every launch goes to the cudaStream_t passed as `stream` (void* here so that C callers need no CUDA headers);
 *     no entry point synchronises the host;
 *   - activations are NHWC ("channels_last"), dtype code ICGAN_F32 or ICGAN_BF16; statistics, weights masters and
 *     gradients of weights are float32.
 *
 * Each declaration cites the reference interface it replaces (paths relative to facebookresearch/ic_gan).
 */
#ifndef ICGAN_B200_H_
#define ICGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICGAN_F32 0
#define ICGAN_BF16 1

#define ICGAN_ACT_NONE 0
#define ICGAN_ACT_RELU 1
#define ICGAN_ACT_TANH 2

const char* icgan_last_error(void);
int icgan_version(void);

/* ------------------------------------------------------------------------------------------------
 * Convolutions (replace F.conv2d + cuDNN behind layers.SNConv2d.forward, BigGAN_PyTorch/layers.py:144-153,
 * and its autograd backward).  Weight tensors are in "kernel layout" [Cout][k][k][Cin] (see icgan_sn_prepare_weight).
 * ---------------------------------------------------------------------------------------------- */

/* Tensor-core implicit GEMM (TMA + tcgen05, bf16 operands, fp32 accumulate in TMEM).
 *   y[n,h,w,co] = act( sum_{kh,kw,ci} x[n,h+kh-p,w+kw-p,ci] * wk[co,kh,kw,ci] + bias[co] + residual[...] )
 * x: [B,H,W,Cin] bf16, wk: [Cout,k,k,Cin] bf16, k in {1,3}, stride 1, pad k/2, Cin%16==0, Cout%8==0, H,W powers of two.
 * out_dtype/res_dtype: ICGAN_F32|ICGAN_BF16. residual may be NULL; res_shift=1 reads residual[n,h/2,w/2,co] from a
 * half-resolution tensor (the nearest-upsampled shortcut of GBlock, layers.py:545-552). bias may be NULL.
 * Serves forward, and dgrad when called with the flipped/transposed weight copy. */
int icgan_conv2d_tc(const void* x, const void* wk, const float* bias, const void* residual, void* y, int B, int H,
                    int W, int Cin, int Cout, int ksize, int out_dtype, int res_dtype, int res_shift, int act,
                    void* stream);

/* Tensor-core weight gradient for 3x3/1x1 stride-1 convs (replaces cudnn_convolution_backward_weight,
 * stylegan2_ada_pytorch/torch_utils/ops/conv2d_gradfix.py:223-227, and ATen's conv backward under BigGAN):
 *   dwk[co,kh,kw,ci] += sum_{n,h,w} dy[n,h,w,co] * x[n,h+kh-p,w+kw-p,ci]
 * x:[B,H,W,Cin] and dy:[B,H,W,Cout] are the NHWC bf16 tensors themselves (MN-major UMMA operands, no transposes).
 * dwk is float32 [Cout,k,k,Cin] and is ACCUMULATED into (caller zeroes it). Cin%16==0, Cout%8==0. */
int icgan_conv2d_wgrad_tc(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin, int Cout,
                          int ksize, void* stream);

/* Generic CUDA-core path (float32 accumulate; any channel counts, e.g. Cin=3 / Cout=3; any stride/pad).
 * H, W are INPUT dims; output is [(H+2*pad-k)/stride+1, ...]. x/y dtype per in_dtype/out_dtype; wk float32
 * [Cout,k,k,Cin]. Same epilogue as icgan_conv2d_tc. */
int icgan_conv2d_simt(const void* x, const float* wk, const float* bias, const void* residual, void* y, int B, int H,
                      int W, int Cin, int Cout, int ksize, int stride, int pad, int in_dtype, int out_dtype,
                      int res_dtype, int res_shift, int act, void* stream);
/* dwk (float32, accumulated) from NHWC x [B,H,W,Cin] and dy [B,Hout,Wout,Cout] of dtype in_dtype. */
int icgan_conv2d_wgrad_simt(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin, int Cout,
                            int ksize, int stride, int pad, int in_dtype, void* stream);
/* out[c] += sum over pixels of x[p][c]  (bias gradient; NHWC column sums). */
int icgan_channel_sum(const void* x, float* out, int64_t pixels, int C, int dtype, void* stream);

/* NHWC [B,H,W,C] -> channel-major [C][B*H*W] bf16. */
int icgan_nhwc_to_cnhw(const void* x, void* xT, int64_t pixels, int C, int in_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICGAN_B200_H_ */
