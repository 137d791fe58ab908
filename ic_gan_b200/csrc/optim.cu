// Fused Adam + EMA over the flat parameter buffer of one network: replaces
// torch.optim.Adam.step (BigGAN_PyTorch/trainer.py:158-171 builds the optimisers, train_fns.py:115,177 steps them) and the
// parameter part of utils.ema.update (BigGAN_PyTorch/utils.py:1055-1067) with ONE HBM pass:
//   g'   = g * grad_scale                               (1/world of the data-parallel mean, folded in)
//   m    = lerp(m, g', 1-beta1)                         (torch: exp_avg.lerp_(grad, 1 - beta1))
//   v    = v*beta2 + (1-beta2) * g'*g'                  (exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2))
//   p   -= (lr / bc1) * m / (sqrt(v)/sqrt(bc2) + eps)   (param.addcdiv_(exp_avg, denom, value=-step_size))
//   ema  = ema*decay + p*(1-decay)                      (utils.py:1062-1066, with the freshly updated p)
// Every operation is rounded where torch's single-tensor CUDA path rounds it (explicit _rn intrinsics, the two FMAs
// nvcc contracts in ATen's lerp/addcmul/addcdiv functors written out), so the result agrees with torch.optim.Adam to
// the last bit or one ulp (tests/test_optim_gpu.py).  HBM-bound: 20 B read + 16 B written per parameter.
#include "common.cuh"

namespace icgan {
namespace {

struct AdamArgs {
  float w1, beta2, one_minus_beta2, inv_bc2_sqrt, eps, neg_step_size, grad_scale, ema_decay, ema_one_minus;
  int use_scale, use_ema;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float& e, const AdamArgs& a) {
  if (a.use_scale) g = __fmul_rn(g, a.grad_scale);
  const float diff = __fsub_rn(g, m);
  m = a.w1 < 0.5f ? __fmaf_rn(a.w1, diff, m) : __fsub_rn(g, __fmul_rn(diff, __fsub_rn(1.f, a.w1)));
  v = __fmaf_rn(a.one_minus_beta2, __fmul_rn(g, g), __fmul_rn(v, a.beta2));
  const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), a.inv_bc2_sqrt), a.eps);
  p = __fmaf_rn(a.neg_step_size, __fdiv_rn(m, denom), p);
  if (a.use_ema) e = __fadd_rn(__fmul_rn(e, a.ema_decay), __fmul_rn(p, a.ema_one_minus));
}

__global__ void __launch_bounds__(256)
adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float* __restrict__ ema, int64_t n, const AdamArgs a) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* e4 = reinterpret_cast<float4*>(ema);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float4 ee = a.use_ema ? e4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    adam_one(pp.x, gg.x, mm.x, vv.x, ee.x, a);
    adam_one(pp.y, gg.y, mm.y, vv.y, ee.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, ee.z, a);
    adam_one(pp.w, gg.w, mm.w, vv.w, ee.w, a);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (a.use_ema) e4[i] = ee;
  }
  // tail (n % 4 elements)
  const int64_t t = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) {
    float pp = p[t], mm = m[t], vv = v[t], ee = a.use_ema ? ema[t] : 0.f;
    adam_one(pp, g[t], mm, vv, ee, a);
    p[t] = pp; m[t] = mm; v[t] = vv;
    if (a.use_ema) ema[t] = ee;
  }
}

// ema = ema*decay + src*(1-decay) over a flat buffer (the non-parameter state entries: BN running statistics, SN u/sv)
__global__ void __launch_bounds__(256)
ema_lerp_kernel(float* __restrict__ ema, const float* __restrict__ src, int64_t n, float decay, float one_minus) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    ema[i] = __fadd_rn(__fmul_rn(ema[i], decay), __fmul_rn(src[i], one_minus));
}

}  // namespace
}  // namespace icgan

using namespace icgan;

extern "C" int icgan_adam_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema,
                                   int64_t n, double lr, double beta1, double beta2, double eps, int64_t step,
                                   double grad_scale, double ema_decay, void* stream) {
  ICGAN_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "icgan_adam_ema_step: bad arguments");
  ICGAN_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                  reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                  reinterpret_cast<uintptr_t>(ema)) & 15u) == 0, "icgan_adam_ema_step: buffers must be 16-byte aligned");
  // scalar arithmetic in double exactly as torch/optim/adam.py does in Python floats, then narrowed to float like ATen
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  AdamArgs a;
  a.w1 = static_cast<float>(1.0 - beta1);
  a.beta2 = static_cast<float>(beta2);
  a.one_minus_beta2 = static_cast<float>(1.0 - beta2);
  a.inv_bc2_sqrt = 1.0f / static_cast<float>(sqrt(bc2));  // ATen divides by a scalar as a multiply by its float reciprocal
  a.eps = static_cast<float>(eps);
  a.neg_step_size = static_cast<float>(-(lr / bc1));
  a.grad_scale = static_cast<float>(grad_scale);
  a.use_scale = grad_scale != 1.0;
  a.use_ema = ema != nullptr && ema_decay >= 0.0;
  a.ema_decay = static_cast<float>(ema_decay);
  a.ema_one_minus = static_cast<float>(1.0 - ema_decay);
  const int64_t work = (n + 3) / 4;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  adam_ema_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, ema, n, a);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_ema_lerp(float* ema, const float* src, int64_t n, double decay, void* stream) {
  ICGAN_REQUIRE(ema && src && n > 0, "icgan_ema_lerp: bad arguments");
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  ema_lerp_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ema, src, n, static_cast<float>(decay), static_cast<float>(1.0 - decay));
  ICGAN_LAUNCH_CHECK();
  return 0;
}
