// GPU self-test of the tcgen05 kernels against a host double-precision reference on sampled outputs.
// Built by __graft_entry__.build() into tests/cuda/tc_selftest and run by tests/test_tc_selftest.py (gpu marker).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/icgan_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t g_seed = 12345u;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct ConvCase {
  const char* name;
  int B, H, W, Cin, Cout, k, out_bf16, res_mode /*0 none,1 same,2 half*/, act, bias;
};

static int run_conv_case(const ConvCase& c, bool time_it) {
  const int taps = c.k * c.k, pad = c.k / 2;
  const size_t nx = (size_t)c.B * c.H * c.W * c.Cin, nw = (size_t)c.Cout * taps * c.Cin,
               ny = (size_t)c.B * c.H * c.W * c.Cout;
  std::vector<float> hx(nx), hw(nw), hb(c.Cout), hres;
  std::vector<__nv_bfloat16> bx(nx), bw(nw);
  for (size_t i = 0; i < nx; ++i) { hx[i] = bf16_round(frand()); bx[i] = __float2bfloat16_rn(hx[i]); }
  for (size_t i = 0; i < nw; ++i) { hw[i] = bf16_round(frand() * 0.25f); bw[i] = __float2bfloat16_rn(hw[i]); }
  for (int i = 0; i < c.Cout; ++i) hb[i] = frand();
  const int rH = c.res_mode == 2 ? c.H / 2 : c.H, rW = c.res_mode == 2 ? c.W / 2 : c.W;
  const size_t nr = c.res_mode ? (size_t)c.B * rH * rW * c.Cout : 0;
  hres.resize(nr);
  for (size_t i = 0; i < nr; ++i) hres[i] = frand();

  void *dx, *dw, *dy;
  float *db = nullptr, *dr = nullptr;
  CK(cudaMalloc(&dx, nx * 2)); CK(cudaMalloc(&dw, nw * 2)); CK(cudaMalloc(&dy, ny * 4));
  CK(cudaMalloc(&db, c.Cout * 4));
  CK(cudaMemcpy(dx, bx.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, bw.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0xFF, ny * 4));
  if (nr) { CK(cudaMalloc(&dr, nr * 4)); CK(cudaMemcpy(dr, hres.data(), nr * 4, cudaMemcpyHostToDevice)); }

  int rc = icgan_conv2d_tc(dx, dw, nullptr, c.bias ? db : nullptr, dr, dy, nullptr, c.B, c.H, c.W, c.Cin, c.Cout, c.k,
                           c.out_bf16 ? ICGAN_BF16 : ICGAN_F32, ICGAN_F32, c.res_mode == 2, c.act, nullptr);
  if (rc) { printf("[%s] launch error %d: %s\n", c.name, rc, icgan_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel error: %s\n", c.name, cudaGetErrorString(e)); exit(3); }

  std::vector<float> hy(ny);
  if (c.out_bf16) {
    std::vector<__nv_bfloat16> t(ny);
    CK(cudaMemcpy(t.data(), dy, ny * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ny; ++i) hy[i] = __bfloat162float(t[i]);
  } else {
    CK(cudaMemcpy(hy.data(), dy, ny * 4, cudaMemcpyDeviceToHost));
  }
  // sampled reference (all corners + random positions)
  const int nsamp = 6000;
  double max_err = 0, max_ref = 0;
  int bad = 0;
  for (int s = 0; s < nsamp; ++s) {
    g_seed = g_seed * 1664525u + 1013904223u;
    int n = (g_seed >> 4) % c.B;
    g_seed = g_seed * 1664525u + 1013904223u;
    int h = (g_seed >> 4) % c.H;
    g_seed = g_seed * 1664525u + 1013904223u;
    int w = (g_seed >> 4) % c.W;
    g_seed = g_seed * 1664525u + 1013904223u;
    int co = (g_seed >> 4) % c.Cout;
    if (s < 8) { h = (s & 1) ? c.H - 1 : 0; w = (s & 2) ? c.W - 1 : 0; n = (s & 4) ? c.B - 1 : 0; }
    double acc = 0;
    for (int kh = 0; kh < c.k; ++kh)
      for (int kw = 0; kw < c.k; ++kw) {
        const int ih = h + kh - pad, iw = w + kw - pad;
        if (ih < 0 || ih >= c.H || iw < 0 || iw >= c.W) continue;
        const float* xp = &hx[(((size_t)n * c.H + ih) * c.W + iw) * c.Cin];
        const float* wp = &hw[((size_t)co * taps + kh * c.k + kw) * c.Cin];
        for (int ci = 0; ci < c.Cin; ++ci) acc += (double)xp[ci] * wp[ci];
      }
    if (c.bias) acc += hb[co];
    if (c.res_mode == 1) acc += hres[(((size_t)n * c.H + h) * c.W + w) * c.Cout + co];
    if (c.res_mode == 2) acc += hres[(((size_t)n * rH + h / 2) * rW + w / 2) * c.Cout + co];
    if (c.act == ICGAN_ACT_RELU) acc = acc > 0 ? acc : 0;
    if (c.act == ICGAN_ACT_TANH) acc = tanh(acc);
    const double got = hy[(((size_t)n * c.H + h) * c.W + w) * c.Cout + co];
    const double err = fabs(got - acc);
    const double tol = (c.out_bf16 ? 1e-2 : 1e-3) * (1.0 + fabs(acc));
    if (!(err <= tol)) {
      if (bad < 5) printf("  [%s] mismatch n=%d h=%d w=%d co=%d got=%g ref=%g\n", c.name, n, h, w, co, got, acc);
      ++bad;
    }
    if (err > max_err) max_err = err;
    if (fabs(acc) > max_ref) max_ref = fabs(acc);
  }
  printf("[%s] conv B=%d %dx%d Cin=%d Cout=%d k=%d : %s  max_err=%.3g (max|ref|=%.3g) bad=%d/%d\n", c.name, c.B, c.H,
         c.W, c.Cin, c.Cout, c.k, bad ? "FAIL" : "ok", max_err, max_ref, bad, nsamp);
  if (time_it && !bad) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i)
      icgan_conv2d_tc(dx, dw, nullptr, db, nullptr, dy, nullptr, c.B, c.H, c.W, c.Cin, c.Cout, c.k, ICGAN_BF16, ICGAN_F32, 0, 0, nullptr);
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i)
      icgan_conv2d_tc(dx, dw, nullptr, db, nullptr, dy, nullptr, c.B, c.H, c.W, c.Cin, c.Cout, c.k, ICGAN_BF16, ICGAN_F32, 0, 0, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double flops = 2.0 * c.B * c.H * c.W * (double)c.Cout * c.Cin * taps;
    printf("    time %.3f ms  -> %.1f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  cudaFree(dx); cudaFree(dw); cudaFree(dy); cudaFree(db);
  if (dr) cudaFree(dr);
  return bad ? 1 : 0;
}

struct WgradCase {
  const char* name;
  int B, H, W, Cin, Cout, k;
};

static int run_wgrad_case(const WgradCase& c, bool time_it) {
  const int taps = c.k * c.k, pad = c.k / 2;
  const size_t P = (size_t)c.B * c.H * c.W;
  const size_t nx = P * c.Cin, ndy = P * c.Cout, nw = (size_t)c.Cout * taps * c.Cin;
  std::vector<float> hx(nx), hdy(ndy);
  for (size_t i = 0; i < nx; ++i) hx[i] = bf16_round(frand());
  for (size_t i = 0; i < ndy; ++i) hdy[i] = bf16_round(frand());
  float* dw;
  void *dxT, *ddyT;  // NHWC bf16 operands
  CK(cudaMalloc(&dw, nw * 4));
  CK(cudaMalloc(&dxT, nx * 2)); CK(cudaMalloc(&ddyT, ndy * 2));
  {
    std::vector<__nv_bfloat16> t(nx);
    for (size_t i = 0; i < nx; ++i) t[i] = __float2bfloat16_rn(hx[i]);
    CK(cudaMemcpy(dxT, t.data(), nx * 2, cudaMemcpyHostToDevice));
    t.resize(ndy);
    for (size_t i = 0; i < ndy; ++i) t[i] = __float2bfloat16_rn(hdy[i]);
    CK(cudaMemcpy(ddyT, t.data(), ndy * 2, cudaMemcpyHostToDevice));
  }
  CK(cudaMemset(dw, 0, nw * 4));
  int rc = icgan_conv2d_wgrad_tc(dxT, ddyT, dw, c.B, c.H, c.W, c.Cin, c.Cout, c.k, nullptr);
  if (rc) { printf("[%s] launch error %d: %s\n", c.name, rc, icgan_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel error: %s\n", c.name, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hw(nw);
  CK(cudaMemcpy(hw.data(), dw, nw * 4, cudaMemcpyDeviceToHost));
  const int nsamp = 300;
  int bad = 0;
  double max_err = 0, max_ref = 0;
  for (int s = 0; s < nsamp; ++s) {
    g_seed = g_seed * 1664525u + 1013904223u;
    const int co = (g_seed >> 4) % c.Cout;
    g_seed = g_seed * 1664525u + 1013904223u;
    const int ci = (g_seed >> 4) % c.Cin;
    g_seed = g_seed * 1664525u + 1013904223u;
    const int tap = (g_seed >> 4) % taps;
    const int kh = tap / c.k, kw = tap % c.k;
    double acc = 0;
    for (int n = 0; n < c.B; ++n)
      for (int h = 0; h < c.H; ++h) {
        const int ih = h + kh - pad;
        if (ih < 0 || ih >= c.H) continue;
        for (int w = 0; w < c.W; ++w) {
          const int iw = w + kw - pad;
          if (iw < 0 || iw >= c.W) continue;
          acc += (double)hdy[(((size_t)n * c.H + h) * c.W + w) * c.Cout + co] *
                 hx[(((size_t)n * c.H + ih) * c.W + iw) * c.Cin + ci];
        }
      }
    const double got = hw[((size_t)co * taps + tap) * c.Cin + ci];
    const double err = fabs(got - acc), tol = 2e-3 * (1.0 + fabs(acc)) + 1e-3 * sqrt((double)P);
    if (!(err <= tol)) {
      if (bad < 5) printf("  [%s] mismatch co=%d tap=%d ci=%d got=%g ref=%g\n", c.name, co, tap, ci, got, acc);
      ++bad;
    }
    if (err > max_err) max_err = err;
    if (fabs(acc) > max_ref) max_ref = fabs(acc);
  }
  printf("[%s] wgrad B=%d %dx%d Cin=%d Cout=%d k=%d : %s  max_err=%.3g (max|ref|=%.3g) bad=%d/%d\n", c.name, c.B, c.H,
         c.W, c.Cin, c.Cout, c.k, bad ? "FAIL" : "ok", max_err, max_ref, bad, nsamp);
  if (time_it && !bad) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) icgan_conv2d_wgrad_tc(dxT, ddyT, dw, c.B, c.H, c.W, c.Cin, c.Cout, c.k, nullptr);
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) icgan_conv2d_wgrad_tc(dxT, ddyT, dw, c.B, c.H, c.W, c.Cin, c.Cout, c.k, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("    time %.3f ms  -> %.1f TFLOP/s\n", ms, 2.0 * P * (double)c.Cout * c.Cin * taps / ms * 1e-9);
  }
  cudaFree(dw); cudaFree(dxT); cudaFree(ddyT);
  return bad ? 1 : 0;
}

// launches only (no host reference): target for `ncu --set full`
static void run_prof() {
  const int B = 32, H = 64, W = 64, Ci = 384, Co = 384, k = 3;
  const size_t nx = (size_t)B * H * W * Ci, ny = (size_t)B * H * W * Co, nw = (size_t)Co * 9 * Ci;
  void *x, *y, *w;
  float* dw;
  CK(cudaMalloc(&x, nx * 2)); CK(cudaMalloc(&y, ny * 2)); CK(cudaMalloc(&w, nw * 2)); CK(cudaMalloc(&dw, nw * 4));
  CK(cudaMemset(x, 0x3c, nx * 2)); CK(cudaMemset(y, 0x3c, ny * 2)); CK(cudaMemset(w, 0x3c, nw * 2));
  CK(cudaMemset(dw, 0, nw * 4));
  for (int i = 0; i < 3; ++i) {
    icgan_conv2d_tc(x, w, nullptr, nullptr, nullptr, y, nullptr, B, H, W, Ci, Co, k, ICGAN_BF16, ICGAN_F32, 0, 0, nullptr);
    icgan_conv2d_wgrad_tc(x, y, dw, B, H, W, Ci, Co, k, nullptr);
  }
  // the 96-channel 256x256 layer (HBM/overhead-bound regime)
  for (int i = 0; i < 2; ++i) {
    icgan_conv2d_tc(x, w, nullptr, nullptr, nullptr, y, nullptr, 8, 256, 256, 96, 96, k, ICGAN_BF16, ICGAN_F32, 0, 0, nullptr);
    icgan_conv2d_wgrad_tc(x, y, dw, 8, 256, 256, 96, 96, k, nullptr);
  }
  CK(cudaDeviceSynchronize());
  printf("prof launches done\n");
}

// prof2 <B> <H> <W> <Cin> <Cout>: five launches of one 3x3 conv shape (for ncu -s/-c selection)
static void run_prof2(int B, int H, int W, int Ci, int Co) {
  const size_t nx = (size_t)B * H * W * Ci, ny = (size_t)B * H * W * Co, nw = (size_t)Co * 9 * Ci;
  void *x, *y, *w;
  CK(cudaMalloc(&x, nx * 2)); CK(cudaMalloc(&y, ny * 2)); CK(cudaMalloc(&w, nw * 2));
  CK(cudaMemset(x, 0x3c, nx * 2)); CK(cudaMemset(w, 0x3c, nw * 2));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) {
    CK(cudaEventRecord(e0));
    int rc = icgan_conv2d_tc(x, w, nullptr, nullptr, nullptr, y, nullptr, B, H, W, Ci, Co, 3, ICGAN_BF16, ICGAN_F32, 0, 0, nullptr);
    if (rc) { printf("launch error %s\n", icgan_last_error()); exit(2); }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("prof2 B=%d %dx%d %d->%d : %.3f ms %.1f TFLOP/s\n", B, H, W, Ci, Co, ms, 2.0 * B * H * W * 9.0 * Ci * Co / ms * 1e-9);
  }
}

int main(int argc, char** argv) {
  if (argc > 6 && !strcmp(argv[1], "prof2")) {
    run_prof2(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
    return 0;
  }
  // usage: tc_selftest [all|conv|wgrad <idx>|perf|prof]
  const char* mode = argc > 1 ? argv[1] : "all";
  if (!strcmp(mode, "prof")) { run_prof(); return 0; }
  const bool convperf = !strcmp(mode, "convperf");  // conv correctness + conv timings only
  const bool perf = !strcmp(mode, "perf") || convperf;
  const bool do_conv = !strcmp(mode, "all") || !strcmp(mode, "conv") || perf;
  const bool do_wgrad = !strcmp(mode, "all") || !strcmp(mode, "wgrad") || (perf && !convperf);
  const int only = (argc > 2) ? atoi(argv[2]) : -1;
  int fails = 0;
  const ConvCase conv_cases[] = {
      {"c1_1x1_k64", 2, 16, 16, 64, 128, 1, 0, 0, 0, 0},
      {"c2_3x3_k128", 3, 32, 32, 128, 96, 3, 0, 0, 0, 1},
      {"c3_3x3_cw32", 2, 64, 64, 96, 192, 3, 0, 1, 0, 1},
      {"c4_oobbatch_n24", 5, 8, 8, 64, 24, 3, 0, 0, 0, 1},
      {"c5_4x4_2ntiles", 20, 4, 4, 256, 512, 3, 1, 0, ICGAN_ACT_RELU, 1},
      {"c6_cw16_1x1", 2, 32, 32, 48, 96, 1, 0, 0, 0, 0},
      {"c7_rowtile_256", 1, 256, 256, 96, 96, 3, 1, 2, ICGAN_ACT_RELU, 1},
      {"c8_gemm_like", 1, 1, 1536, 2048, 768, 1, 0, 0, 0, 1},
      {"c9_tanh_cout8", 2, 16, 16, 32, 8, 3, 0, 0, ICGAN_ACT_TANH, 1},
      {"c10_many_tiles", 4, 64, 64, 192, 384, 3, 1, 1, 0, 1},
      // 3x3 shapes with H, W multiples of 16 take the halo-reuse kernel (16x16-pixel tiles, 18-row boxes)
      {"h1_one_tile", 2, 16, 16, 64, 128, 3, 0, 0, 0, 1},
      {"h2_cw32_cout24_nonsq", 3, 32, 48, 96, 24, 3, 0, 1, 0, 1},
      {"h3_cw16_cout200_tail", 2, 64, 64, 48, 200, 3, 1, 0, ICGAN_ACT_RELU, 1},
      {"h4_res_half_3ntiles", 1, 32, 32, 256, 384, 3, 1, 2, 0, 1},
      {"h5_many_units", 1, 16, 32, 512, 96, 3, 0, 0, 0, 0},
  };
  if (do_conv)
    for (const ConvCase& c : conv_cases) fails += run_conv_case(c, false);
  const WgradCase wg_cases[] = {
      {"w0_3x3_64x64", 1, 64, 64, 64, 128, 3},
      {"w1_3x3_64", 2, 16, 16, 64, 64, 3},
      {"w2_3x3_96", 2, 32, 32, 96, 96, 3},
      {"w3_1x1_192", 3, 8, 8, 192, 384, 1},
      {"w4_3x3_big", 2, 64, 64, 192, 384, 3},
      {"w5_3x3_8x8", 6, 8, 8, 256, 256, 3},
      {"w6_1x1_64x64", 1, 64, 64, 64, 128, 1},
      {"w7_3x3_4x4", 24, 4, 4, 128, 256, 3},
      {"w8_1x1_cin96", 2, 32, 32, 96, 192, 1},
      {"w9_3x3_cout24", 2, 16, 16, 48, 24, 3},
      {"w10_halo_mb3_nonsq", 2, 32, 16, 80, 192, 3},
      {"w11_halo_h4", 3, 4, 16, 64, 320, 3},
  };
  if (do_wgrad) {
    int i = 0;
    for (const WgradCase& c : wg_cases) {
      if (only < 0 || only == i) fails += run_wgrad_case(c, false);
      ++i;
    }
  }
  if (perf && !fails) {
    const ConvCase perf_cases[] = {
        {"p1_384@64", 32, 64, 64, 384, 384, 3, 1, 0, 0, 1},
        {"p2_96@256", 8, 256, 256, 96, 96, 3, 1, 0, 0, 1},
        {"p3_1536@8", 64, 8, 8, 1536, 1536, 3, 1, 0, 0, 1},
        {"p4_768@32", 32, 32, 32, 768, 768, 3, 1, 0, 0, 1},
        {"p5_192@128", 16, 128, 128, 192, 192, 3, 1, 0, 0, 1},
        {"p6_192to96@256", 8, 256, 256, 192, 96, 3, 1, 0, 0, 1},
        {"p7_1536@16", 64, 16, 16, 1536, 1536, 3, 1, 0, 0, 1},
        {"p8_768to384@64", 32, 64, 64, 768, 384, 3, 1, 0, 0, 1},
        // D's first residual conv at BASELINE config 3's micro-batch (2 x 128 images): 1.6 G elements per tensor,
        // fp32 output = 6.4 GB (byte offsets beyond 32 bits)
        {"p9_96@256_B256", 256, 256, 256, 96, 96, 3, 0, 0, 0, 1},
    };
    for (const ConvCase& c : perf_cases) fails += run_conv_case(c, true);
    if (convperf) {
      printf("TC_SELFTEST %s (%d failing cases)\n", fails ? "FAILED" : "PASSED", fails);
      return fails ? 1 : 0;
    }
    const WgradCase wperf[] = {{"wp1_384@64", 32, 64, 64, 384, 384, 3}, {"wp2_96@256", 8, 256, 256, 96, 96, 3},
                               {"wp3_768@32", 32, 32, 32, 768, 768, 3}};
    for (const WgradCase& c : wperf) fails += run_wgrad_case(c, true);
  }
  printf("TC_SELFTEST %s (%d failing cases)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
