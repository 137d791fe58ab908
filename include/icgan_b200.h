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
#define ICGAN_F16 2 /* StyleGAN2 ops only */

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
 * half-resolution tensor (the nearest-upsampled shortcut of GBlock, layers.py:545-552); res_shift=2 uses `residual` as a
 * gate instead of an addend, y = residual[n,h,w,co] > 0 ? y : 0 -- the backward of the F.relu that produced this conv's
 * input (layers.py:594,603 DBlock), fused into the dgrad launch. bias may be NULL.
 * alpha_dev (device scalar, may be NULL): the accumulator is scaled by it before bias -- 1/sigma of SN.W_ (layers.py:112).
 * bn_stats (may be NULL; needs act none, Cout%32==0): float32 [2*Cout], ACCUMULATES sum and sum of squares over all output
 * pixels of (y - bias), i.e. the batch statistics the following ccbn/bn needs (layers.py:412-421), from the fp32
 * accumulators -- saves the two statistics passes over the activation (see icgan_bn_stats_from_sums).
 * Serves forward, and dgrad when called with the flipped/transposed weight copy. */
int icgan_conv2d_tc(const void* x, const void* wk, const float* alpha_dev, const float* bias, const void* residual,
                    void* y, float* bn_stats, int B, int H, int W, int Cin, int Cout, int ksize, int out_dtype,
                    int res_dtype, int res_shift, int act, void* stream);

/* RGB-side input convolution on the tensor cores with the im2col fused into the kernel (the first conv of D,
 * BigGAN_PyTorch/BigGAN.py:491-495 -> layers.py DBlock.conv1; replaces F.conv2d on a 3-channel image):
 *   y[n,h,w,co] = act(alpha * sum_{kh,kw,ci} x[n,h+kh-p,w+kw-p,ci] * wcol[co, (kh*k+kw)*Cs+ci] + bias[co])
 * x:[B,H,W,Cs] bf16 with Cs<=3 and k*k*Cs<=32; wcol:[Cout,32] bf16 (columns >= k*k*Cs zero); Cout%8==0, Cout<=256. */
int icgan_conv2d_rgb_tc(const void* x, const void* wcol, const float* alpha_dev, const float* bias, void* y, int B, int H,
                        int W, int Cs, int Cout, int ksize, int out_dtype, int act, void* stream);

/* Tensor-core weight gradient for 3x3/1x1 stride-1 convs (replaces cudnn_convolution_backward_weight,
 * stylegan2_ada_pytorch/torch_utils/ops/conv2d_gradfix.py:223-227, and ATen's conv backward under BigGAN):
 *   dwk[co,kh,kw,ci] += sum_{n,h,w} dy[n,h,w,co] * x[n,h+kh-p,w+kw-p,ci]
 * x:[B,H,W,Cin] and dy:[B,H,W,Cout] are the NHWC bf16 tensors themselves (MN-major UMMA operands, no transposes).
 * dwk is float32 [Cout,k,k,Cin] and is ACCUMULATED into (caller zeroes it). Cin%16==0, Cout%8==0. */
int icgan_conv2d_wgrad_tc(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin, int Cout,
                          int ksize, void* stream);

/* Generalised tensor-core convolution: an explicit tap list instead of a k x k window, optional input stride 2 and an
 * affine output-pixel mapping -- what the strided and transposed convolutions of StyleGAN2's conv2d_resample need
 * (stylegan2_ada_pytorch/torch_utils/ops/conv2d_resample.py:152-195 -> conv2d_gradfix.conv2d(stride=2) /
 * conv_transpose2d(stride=2), i.e. cudnn_convolution / cudnn_convolution_transpose) without zero-insertion:
 *   for (h, w) in the Hd x Wd tile domain:
 *     y[n, h*osy+ooy, w*osx+oox, co] = alpha * sum_t sum_ci x[n, h*in_stride+tap_dh[t], w*in_stride+tap_dw[t], ci] * wk[co, tap_w[t], ci]
 *                                      + bias[co] + residual[n, h*osy+ooy, w*osx+oox, co]     (alpha_dev: device scalar or NULL = 1)
 * x: [B,H,W,Cin] bf16 (reads outside the tensor are zero), wk: [Cout, wtaps, Cin] bf16, y/residual: [B,OH,OW,Cout];
 * res_mask = 1: `residual` is not added but gates the result (y = residual > 0 ? y : 0: a ReLU backward fused into a dgrad).
 * tap_* are HOST int arrays of ntaps (<= 16) entries. A stride-2 convolution is one call (in_stride=2); a stride-2
 * transposed convolution is four calls, one per output parity class (osy=osx=2, oo* = parity), each with the 1, 2 or 4
 * taps that reach that class. */
int icgan_conv2d_tc_ex(const void* x, const void* wk, const float* alpha_dev, const float* bias, const void* residual,
                       void* y, int B, int H, int W, int Cin, int Cout, int wtaps, int ntaps, const int* tap_dh_host, const int* tap_dw_host,
                       const int* tap_w_host, int in_stride, int Hd, int Wd, int OH, int OW, int osy, int ooy, int osx,
                       int oox, int out_dtype, int res_dtype, int res_mask, void* stream);
/* Weight gradient of the same family (replaces cudnn_convolution_backward_weight / cudnn_convolution_transpose_backward_weight,
 * conv2d_gradfix.py:223-227):  out[ca, t, cb] += sum_{n,h,w} a[n,h,w,ca] * b[n, h*in_stride+tap_dh[t], w*in_stride+tap_dw[t], cb]
 * a: [B,Ha,Wa,Ca] bf16, b: [B,Hb,Wb,Cb] bf16 (zero outside), out: float32 [Ca, ntaps, Cb], ACCUMULATED. Cb%16==0, Ca%8==0. */
int icgan_conv2d_wgrad_tc_ex(const void* a, const void* b, float* out, int B, int Ha, int Wa, int Ca, int Hb, int Wb,
                             int Cb, int ntaps, const int* tap_dh_host, const int* tap_dw_host, int in_stride,
                             void* stream);

/* Generic CUDA-core path (float32 accumulate; any channel counts, e.g. Cin=3 / Cout=3; any stride/pad).
 * H, W are INPUT dims; output is [(H+2*pad-k)/stride+1, ...]. x/y dtype per in_dtype/out_dtype; wk float32
 * [Cout,k,k,Cin]. Same epilogue as icgan_conv2d_tc. */
int icgan_conv2d_simt(const void* x, const float* wk, const float* alpha_dev, const float* bias, const void* residual,
                      void* y, int B, int H, int W, int Cin, int Cout, int ksize, int stride, int pad, int in_dtype,
                      int out_dtype, int res_dtype, int res_shift, int act, void* stream);
/* dwk (float32, accumulated) from NHWC x [B,H,W,Cin] and dy [B,Hout,Wout,Cout] of dtype in_dtype. */
int icgan_conv2d_wgrad_simt(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin, int Cout,
                            int ksize, int stride, int pad, int in_dtype, void* stream);
/* Image-side layers where Cin<=4 or Cout<=4 (RGB): HBM-bound streaming kernels instead of GEMM tiles. wk float32
 * [Cout,k,k,Cin], k in {1,3}, stride 1, pad k/2; serves forward and (with the dgrad weight copy) dgrad. */
int icgan_conv2d_small(const void* x, const float* wk, const float* alpha_dev, const float* bias, void* y, int B, int H,
                       int W, int Cin, int Cout, int ksize, int in_dtype, int out_dtype, int act, void* stream);
int icgan_conv2d_wgrad_small(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin, int Cout,
                             int ksize, int x_dtype, int dy_dtype, void* stream);
/* Explicit im2col of a <=4-channel NHWC tensor: out[p][tap*Cs+ci] = x[p+tap][ci] (zero outside / beyond k*k*Cs),
 * out bf16 [B,H,W,KP]; lets the RGB-side layers run on icgan_conv2d_tc / icgan_conv2d_wgrad_tc as 1x1 convs. */
int icgan_im2col_small(const void* x, void* out, int B, int H, int W, int Cs, int ksize, int KP, int in_dtype,
                       void* stream);
/* out[c] += sum over pixels of x[p][c]  (bias gradient; NHWC column sums). */
int icgan_channel_sum(const void* x, float* out, int64_t pixels, int C, int dtype, void* stream);

/* NHWC [B,H,W,C] -> channel-major [C][B*H*W] bf16. */
int icgan_nhwc_to_cnhw(const void* x, void* xT, int64_t pixels, int C, int in_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral normalisation (layers.py:39-61 power_iteration, :98-112 SN.W_) — batched over all SN layers of a network.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* W;  /* [rows, cols] row-major master weight (OIHW flattened / [out,in] / [num_embeddings, dim]) */
  float* u;        /* [rows]  in/out: the `u0` buffer; overwritten with u' iff update_u (training)            */
  float* v;        /* [cols]  out: normalised right vector                                                     */
  float* u_new;    /* [rows]  out: normalised left vector u' that defines sigma (needed by the backward)       */
  float* sigma;    /* [2]     out: sigma = u'^T W v, 1/sigma                                                   */
  float* scratch;  /* [2]     workspace                                                                        */
  int rows, cols;
} IcganSnLayer;

/* One power-iteration step for n_layers layers (descriptor table in DEVICE memory). max_rows/max_cols size the
 * row-dot grid; wt_u_items = sum over layers of ceil(cols/128)*ceil(rows/64) sizes the flat W^T u grid. */
int icgan_sn_power_iteration(const IcganSnLayer* layers_dev, int n_layers, int max_rows, int max_cols,
                             int64_t wt_u_items, float eps, int update_u, void* stream);
/* W (float32 OIHW) * inv_sigma -> wk_fwd [Cout,k,k,Cin] and/or wk_dgrad [Cin,k,k,Cout] (taps flipped), out_dtype.
 * inv_sigma_dev NULL = no scaling. ksize=1 covers SNLinear/SNEmbedding ([out,in] stays [out,in]). */
int icgan_sn_prepare_weight(const float* W, const float* inv_sigma_dev, void* wk_fwd, void* wk_dgrad, int Cout,
                            int Cin, int ksize, int out_dtype, void* stream);
/* Gradient of the scaled weight (kernel layout, float32) -> gradient of the master weight (OIHW):
 *   dW = (G - <G, W/sigma> u' v^T) / sigma       (sigma NULL: plain re-layout). scratch: 1 float. */
int icgan_sn_weight_grad(const float* G_k, const float* W, const float* u_new, const float* v, const float* sigma,
                         float* scratch, float* dW, int Cout, int Cin, int ksize, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batch norm with per-sample affine (layers.ccbn, layers.py:398-437; layers.bn :485-503), NHWC.
 * ---------------------------------------------------------------------------------------------- */
/* Batch statistics over P = B*H*W pixels: mean, invstd = rsqrt(biased var + eps); running buffers (may be NULL)
 * updated with momentum and the UNBIASED variance as F.batch_norm does. ws: 2*C floats of workspace. float32 input:
 * exact two-pass (mean, then centred squares); bf16 input: ONE pass of moments shifted by shift[c] (NULL = 0; pass the
 * previous step's batch mean to keep E[d^2]-E[d]^2 well conditioned). */
int icgan_bn_train_stats(const void* x, int64_t P, int C, int dtype, float* ws, const float* shift,
                         float* running_mean, float* running_var, float* mean, float* invstd, float eps,
                         float momentum, void* stream);
/* mean/invstd (+ running-stat update) from the shifted sums a conv epilogue accumulated: sums = [sum(y-shift) | sum((y-shift)^2)],
 * shift = the conv bias (NULL = 0). Same outputs as icgan_bn_train_stats. */
int icgan_bn_stats_from_sums(const float* sums, const float* shift, int64_t P, int C, float* running_mean,
                             float* running_var, float* mean, float* invstd, float eps, float momentum, void* stream);
/* y = act(((x-mean)*invstd) * gain[n,c] + bias[n,c]); gain/bias row stride gain_stride (0 = shared [C] vectors);
 * relu: fuse ReLU; up: write the nearest-upsampled x2 tensor [B,2H,2W,C] (GBlock, layers.py:543-546). */
int icgan_bn_apply(const void* x, void* y, const float* mean, const float* invstd, const float* gain,
                   const float* bias, int gain_stride, int B, int H, int W, int C, int relu, int up, int in_dtype,
                   int out_dtype, void* stream);
/* Backward, step 1: s1[n,c] = sum_hw g, s2[n,c] = sum_hw g*xhat, g = dL/d(affine output) after undoing up/relu. */
int icgan_bn_bwd_reduce(const void* x, const void* dy, const float* mean, const float* invstd, const float* gain,
                        const float* bias, int gain_stride, float* s1, float* s2, int B, int H, int W, int C, int relu,
                        int up, int x_dtype, int dy_dtype, void* stream);
/* Backward, step 2: dx = invstd * (gain*g - m1[c] - xhat*m2[c]); m1/m2 = batch means of gain*g and gain*g*xhat
 * (zeros in eval mode). dx has dy's dtype. */
int icgan_bn_bwd_apply(const void* x, const void* dy, void* dx, const float* mean, const float* invstd,
                       const float* gain, const float* bias, int gain_stride, const float* m1, const float* m2, int B,
                       int H, int W, int C, int relu, int up, int x_dtype, int dy_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise / pooling (F.relu, torch.tanh backward, nn.AvgPool2d(2), F.max_pool2d, F.interpolate(x2),
 * D's sum pooling BigGAN.py:624, attention residual layers.py:244).
 * ---------------------------------------------------------------------------------------------- */
int icgan_relu(const void* x, void* y, int64_t n, int dtype, void* stream);
int icgan_relu_bwd(const void* dy, const void* ref, void* dx, int64_t n, int ref_dtype, int g_dtype, void* stream);
int icgan_tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, int y_dtype, int g_dtype, void* stream);
/* out = alpha*a + beta*b (b may be NULL); alpha_dev/beta_dev, when non-NULL, are device scalars used instead. */
int icgan_axpby(const void* a, const void* b, void* out, float alpha, const float* alpha_dev, float beta,
                const float* beta_dev, int64_t n, int dtype, void* stream);
int icgan_dot(const void* a, const void* b, float* out, int64_t n, int dtype, void* stream);
/* mode 0: y = scale * sum_2x2(x) (+ add) ; mode 1: y = max_2x2(x).  x:[B,2Hout,2Wout,C] -> y:[B,Hout,Wout,C] */
int icgan_pool2(const void* x, const void* add, void* y, int B, int Hout, int Wout, int C, float scale, int mode,
                int dtype, void* stream);
/* mode 0: dx = scale * nearest_up2(dy) ; mode 1: max-pool backward using the forward input xref. */
int icgan_unpool2(const void* dy, const void* xref, void* dx, int B, int Hout, int Wout, int C, float scale, int mode,
                  int ref_dtype, int g_dtype, void* stream);
int icgan_relu_sumpool(const void* x, float* out, int B, int HW, int C, int dtype, void* stream);
int icgan_relu_sumpool_bwd(const void* x, const float* dh, void* dx, int B, int HW, int C, int dtype, void* stream);
int icgan_softmax_rows(const void* s, void* p, int64_t rows, int cols, int s_dtype, int p_dtype, void* stream);
int icgan_softmax_rows_bwd(const void* p, const void* dp, void* ds, int64_t rows, int cols, int p_dtype, int dp_dtype,
                           int ds_dtype, void* stream);

/* Strided batched GEMM on CUDA cores, float32 accumulate (SNLinear layers.py:164-165; attention bmm layers.py:237-243):
 * C[b][m][n] = alpha*(alpha_dev?*alpha_dev:1) * sum_k A[b][m][k]*B[b][k][n] + bias[n] + beta*C[b][m][n]; strides in elements. */
int icgan_gemm(const void* A, const void* B, void* C, int M, int N, int K, int batch, int64_t sam, int64_t sak,
               int64_t sab, int64_t sbk, int64_t sbn, int64_t sbb, int64_t scm, int64_t scn, int64_t scb, float alpha,
               const float* alpha_dev, float beta, const float* bias, int a_dtype, int b_dtype, int c_dtype,
               void* stream);

/* Batched tensor-core GEMM (tcgen05, bf16 operands, fp32 accumulate) for the attention products torch.bmm computes in
 * layers.Attention.forward (layers.py:237-243) and their backward: C[b] = alpha * op(A[b]) op(B[b]), C [M,N] row-major
 * (ldc), dtype c_dtype. a_mn/b_mn = 0: operand stored [rows][K] (K contiguous, ld = row stride); = 1: stored [K][rows]
 * (rows contiguous, ld = stride between K indices). sab/sbb/scb: batch strides (elements). ld*, N multiples of 8. */
int icgan_gemm_tc(const void* A, const void* B, void* C, int M, int N, int K, int batch, int a_mn, int b_mn, int64_t lda,
                  int64_t sab, int64_t ldb, int64_t sbb, int64_t ldc, int64_t scb, float alpha, int c_dtype,
                  void* stream);

/* Fused non-local block core (layers.Attention.forward, BigGAN_PyTorch/layers.py:233-243: beta = softmax(theta^T phi),
 * o = g beta^T) on tcgen05, bf16 operands: theta [B,Q,d], phi [B,Kk,d], g [B,Kk,dv], o [B,Q,dv]; Q and Kk multiples of
 * 128, d a multiple of 8 in [8,64], dv a multiple of 16 in [16,192].  The logits and probabilities stay on the SM.
 * lse2 (nullable): per-row log2-sum-exp2 of the logits * log2(e), float32 [B,Q], what the backward restarts from. */
int icgan_attn_fwd(const void* theta, const void* phi, const void* g, void* o, float* lse2, int B, int Q, int Kk, int d,
                   int dv, void* stream);
/* Query side of its backward: recomputes the probabilities from lse2, forms dP = dout g^T in tensor memory,
 * ds = probs * (dP - rowsum(dout * o)) and dtheta = ds phi [B,Q,d].  ds (nullable): the bf16 [B,Q,Kk] copy of ds;
 * dsum: rowsum(dout * o), float32 [B,Q], filled by a pre-pass of this call and also the input of the key side. */
int icgan_attn_bwd_q(const void* theta, const void* phi, const void* g, const void* o, const void* dout,
                     const float* lse2, void* dtheta, void* ds, float* dsum, int B, int Q, int Kk, int d, int dv,
                     void* stream);
/* Key side: dphi = ds^T theta [B,Kk,d] and dg = probs^T dout [B,Kk,dv], with probs and ds rebuilt 128 keys x 64 queries
 * at a time in tensor / shared memory from lse2 and dsum. */
int icgan_attn_bwd_kv(const void* theta, const void* phi, const void* g, const void* dout, const float* lse2,
                      const float* dsum, void* dphi, void* dg, int B, int Q, int Kk, int d, int dv, void* stream);

/* ------------------------------------------------------------------------------------------------
 * k-NN conditioning build (replaces faiss.IndexFlatL2.add/search in ILSVRC_HDF5_feats._obtain_nns,
 * data_utils/datasets_common.py:695-745; output format of data_utils/make_hdf5_nns.py:132-172).
 * X: [N, d] float32 row-major (the float32 cast of the float64-normalised features, datasets_common.py:422-428).
 * ---------------------------------------------------------------------------------------------- */
/* split-bf16 copies (hi, lo) of X and squared norms; Xhi/Xlo: [N, d] bf16, norms: [N] float32. */
int icgan_knn_prepare(const float* X, void* Xhi, void* Xlo, float* norms, int64_t N, int d, void* stream);
/* Tensor-core distance sweep for query rows [q_begin, q_end): per row the C (<=64) smallest coarse squared distances
 * (ascending) and their database indices. passes = 3 (hi.hi+hi.lo+lo.hi, ~fp32 accuracy) or 1 (bf16). d %% 8 == 0. */
int icgan_knn_coarse(const void* Xhi, const void* Xlo, const float* norms, int64_t N, int d, int64_t q_begin,
                     int64_t q_end, int C, int passes, int* cand_idx, float* cand_d, void* stream);
/* Exact float64 re-rank of the candidates: nn_out [nq, k] int64 (self removed by value, ties -> lower index),
 * radius_out [nq] (float32 sqrt of the (k+1)-th squared distance, stored as double), flags[nq] = 1 where the row could
 * not be certified (use icgan_knn_exact_row), *max_err = max |coarse - exact| seen (float, device). */
int icgan_knn_rerank(const float* X, int64_t N, int d, int64_t q_begin, int64_t q_end, int C, int k,
                     const int* cand_idx, const float* cand_d, int64_t* nn_out, double* radius_out, int* flags,
                     float* max_err, float margin, void* stream);
/* Brute-force float64 answer for one row (scratch: N doubles). */
int icgan_knn_exact_row(const float* X, int64_t N, int d, int64_t row, int k, double* scratch, int64_t* nn_out_row,
                        double* radius_out_row, void* stream);

/* ------------------------------------------------------------------------------------------------
 * StyleGAN2-ADA native ops (the reference's pybind11 plugins).
 * ---------------------------------------------------------------------------------------------- */
/* Replaces `Tensor bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)`
 * (stylegan2_ada_pytorch/torch_utils/ops/bias_act.cpp:35; kernel bias_act.cu:26-150). Flat over n elements of `dtype`
 * (ICGAN_F32/BF16/F16); b (same dtype, may be NULL) is indexed by (i / step_b) %% size_b (step_b = prod of dims after
 * `dim`, valid for dense NCHW and channels-last alike); act = cuda_idx 1..9 of bias_act.py:26-99; grad 0/1/2;
 * clamp < 0 disables clamping; xref/yref/dy may be NULL ("empty tensor" in the reference). */
int icgan_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y, int64_t n,
                   int64_t step_b, int size_b, int grad, int act, float alpha, float gain, float clamp, int dtype,
                   void* stream);
/* Replaces `Tensor upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)`
 * (upfirdn2d.cpp:19; kernels upfirdn2d.cu:32-203). x: [N,C,inH,inW] (channels_last=0) or NHWC memory (=1), f: float32
 * [fh,fw]; y: [N,C,outH,outW] with outW = (inW*upx+padx0+padx1-fw+downx)/downx (upfirdn2d.cpp:35-36). */
int icgan_upfirdn2d(const void* x, const float* f, void* y, int N, int C, int inH, int inW, int fh, int fw, int upx,
                    int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                    int channels_last, int dtype, void* stream);

/* Channels-last (NHWC) fast paths of the two ops above plus the elementwise pieces of modulated_conv2d
 * (stylegan2_ada_pytorch/training/networks.py:77-95: x*styles -> conv -> *dcoefs + noise -> bias_act), HBM-bound, one
 * pass each.  Epilogue inputs are float32: pre_scale [N,C] (demodulation coefficients), noise [N or 1, outH, outW],
 * noise_strength (device scalar, NULL = 1), bias [C], s2 [N,C] (styles of the NEXT layer; y2 = y*s2 is its modulated
 * input). act: 0 = no epilogue, 1 = linear, 3 = lrelu (bias_act.py:26-99 ids); clamp < 0 disables clamping. */
/* Shared-memory-tiled upfirdn2d (upfirdn2d.cu:100-203 `upfirdn2d_kernel_small` re-designed for NHWC), 4x4 filter,
 * (up, down) in {(1,1), (2,1), (1,2)}:  y = epilogue(gain * upfirdn(x)); x [N,inH,inW,C], y [N,outH,outW,C].
 * fx_host / fy_host (HOST arrays of 4 floats, or both NULL): when the filter is the outer product fy (x) fx -- every
 * resampling filter setup_filter builds from 1-D taps -- pass the factors and the kernel filters separably. */
int icgan_upfirdn2d_nhwc(const void* x, const float* f4x4, void* y, int N, int C, int inH, int inW, int up, int down,
                         int padx0, int padx1, int pady0, int pady1, int flip, float gain, const float* pre_scale,
                         const float* noise, const float* noise_strength, int noise_per_sample, const float* bias,
                         int act, float alpha, float act_gain, float clamp, const float* s2, void* y2,
                         const float* fx_host, const float* fy_host, int dtype, void* stream);
/* y[n,p,c] = x[n,p,c] * s[n,c]  (networks.py:78 `x * styles`), optional float32 <-> bfloat16 cast. C % 8 == 0. */
int icgan_modulate(const void* x, const float* s, void* y, int N, int64_t hw, int C, int in_dtype, int out_dtype,
                   void* stream);
/* out[n,c] = sum_p a[n,p,c]*b[n,p,c]  (the adjoint of icgan_modulate: gradients w.r.t. styles / dcoefs). out float32 [N,C]. */
int icgan_chan_dot(const void* a, const void* b, float* out, int N, int64_t hw, int C, int a_dtype, int b_dtype,
                   void* stream);
/* bias_act for NHWC tensors, 8 elements per thread (bias_act.cu:26-150 for act 1/3, grad 0/1):
 *   grad 0: y = clamp(act(x*pre_scale[n,c] + noise[n,p]*noise_strength + bias[c]) * gain)
 *   grad 1: y = x * gain * act'(yref) * [|yref| < clamp] * pre_scale[n,c]       (x = incoming gradient) */
int icgan_bias_act_nhwc(const void* x, const void* yref, void* y, const float* bias, const float* pre_scale,
                        const float* noise, const float* noise_strength, int noise_per_sample, int N, int64_t hw, int C,
                        int grad, int act, float alpha, float gain, float clamp, int dtype, void* stream);

/* First-order backward of icgan_bias_act_nhwc's forward in one pass (replaces the reference's bias_act grad kernel,
 * fma backward and their reductions, bias_act.py:230-300, fma.py:31-52), bfloat16 NHWC:
 *   t = dy*gain*act'(y)*[|y|<clamp]; dx = t*pre_scale[n,c]; dpre[n,c] = sum_p t*x; dbias_n[n,c] = sum_p t; dnoise[n,p] = sum_c t.
 * x / pre_scale / dpre / dbias_n / dnoise may be NULL. C/8 must divide 256. */
int icgan_mod_bias_act_bwd(const void* dy, const void* y, const void* x, const float* pre_scale, void* dx, float* dpre,
                           float* dbias_n, float* dnoise, int N, int64_t hw, int C, int act, float alpha, float gain,
                           float clamp, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step fused with the EMA of the generator.
 * ---------------------------------------------------------------------------------------------- */
/* One torch.optim.Adam step (amsgrad off, weight_decay 0, maximize off -- every IC-GAN config; BigGAN_PyTorch/trainer.py:
 * 158-171, stepped at train_fns.py:115,177) over a FLAT float32 parameter buffer of n elements, `step` = 1-based step count:
 *   g' = grad*grad_scale; m = lerp(m, g', 1-beta1); v = v*beta2 + (1-beta2)*g'^2; p -= lr/(1-beta1^step) * m / (sqrt(v)/
 *   sqrt(1-beta2^step) + eps); and, when ema != NULL and ema_decay >= 0, utils.ema.update's rule for the same elements
 *   (BigGAN_PyTorch/utils.py:1062-1066): ema = ema*ema_decay + p*(1-ema_decay). All buffers 16-byte aligned. */
int icgan_adam_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                        double lr, double beta1, double beta2, double eps, int64_t step, double grad_scale,
                        double ema_decay, void* stream);
/* ema = ema*decay + src*(1-decay) over n floats: the non-parameter state entries utils.ema also averages
 * (BN running statistics, SN u0/sv0; utils.py:1060-1066). */
int icgan_ema_lerp(float* ema, const float* src, int64_t n, double decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICGAN_B200_H_ */
