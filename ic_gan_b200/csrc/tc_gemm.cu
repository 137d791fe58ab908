// Batched tensor-core GEMM for the non-local block (layers.Attention.forward, BigGAN_PyTorch/layers.py:227-244) and its
// backward:  C[b] = alpha * op(A[b]) * op(B[b]),  bf16 operands, fp32 accumulation in TMEM, C in bf16 or fp32.
//
// Each operand is either "K-major" (stored [rows][K], K contiguous) or "MN-major" (stored [K][rows], rows contiguous);
// both feed tcgen05.mma directly from TMA-written SWIZZLE_128B tiles (instruction-descriptor bits 15/16 select the
// major-ness), so none of the six attention products needs a transposed copy:
//     S  = theta phi^T        (A K-major,  B K-major)        O  = P g            (A K-major,  B MN-major)
//     dP = dO g^T             (A K-major,  B K-major)        dg = P^T dO         (A MN-major, B MN-major)
//     dtheta = dS phi         (A K-major,  B MN-major)       dphi = dS^T theta   (A MN-major, B MN-major)
// Same persistent, warp-specialised pipeline as tc_conv_kernel (TMA producer / MMA issuer / 4 epilogue warps,
// double-buffered TMEM accumulators).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace icgan {
namespace {

constexpr int kStagesMax = 8;
constexpr int kThreadsG = 320;  // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant)
constexpr uint32_t kSmemBudgetG = 227u * 1024u;

struct GemmTcParams {
  int M, N, K, batch;
  int a_mn, b_mn;  // 1 = MN-major operand
  int m_tiles, n_tiles, total_tiles, k_iters, BN, b_boxes, stages;
  uint32_t a_bytes, b_tx, stage_bytes, idesc;
  int64_t ldc, scb;
  int out_bf16, c16, c8;  // bf16 (f32) output is 32-byte aligned at every multiple of 16 (8) columns
  float alpha;
  void* C;
};

constexpr int kKC = 64;  // reduction extent per pipeline stage

__device__ __forceinline__ void st256(void* ptr, const uint32_t (&w)[8]) {  // one full 32-byte sector
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

__global__ void __launch_bounds__(kThreadsG, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty = full + kStagesMax;
  uint64_t* tfull = empty + kStagesMax;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {  // whole warp runs the loops (uniform control flow); one elected lane issues the TMA / MMA instructions
      if (lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles, mt = (tile / p.n_tiles) % p.m_tiles, b = tile / (p.n_tiles * p.m_tiles);
        const int m0 = mt * 128, n0 = nt * p.BN;
        for (int it = 0; it < p.k_iters; ++it) {
          const int k0 = it * kKC;
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* sa = smem + static_cast<size_t>(stage) * p.stage_bytes;
          uint8_t* sb = sa + p.a_bytes;
          if (elect_one_sync()) {
            mbar_expect_tx(&full[stage], p.a_bytes + p.b_tx);
            if (!p.a_mn) {
              tma_load_3d(sa, &tmA, &full[stage], k0, m0, b);
            } else {
              tma_load_3d(sa, &tmA, &full[stage], m0, k0, b);
              tma_load_3d(sa + 8192, &tmA, &full[stage], m0 + 64, k0, b);
            }
            if (!p.b_mn) {
              tma_load_3d(sb, &tmB, &full[stage], k0, n0, b);
            } else {
              for (int j = 0; j < p.b_boxes; ++j) tma_load_3d(sb + j * 8192, &tmB, &full[stage], n0 + 64 * j, k0, b);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = base + static_cast<uint32_t>(stage) * p.stage_bytes;
          const uint32_t sb = sa + p.a_bytes;
          const uint64_t da = p.a_mn ? desc_mn(sa, 8192) : umma_desc_kmajor(sa, 128);
          const uint64_t db = p.b_mn ? desc_mn(sb, 8192) : umma_desc_kmajor(sb, 128);
          const uint32_t sa_step = p.a_mn ? 128u : 2u, sb_step = p.b_mn ? 128u : 2u;  // 16 K elements per UMMA
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < kKC / 16; ++k)
              umma_bf16(d_tmem, da + sa_step * k, db + sb_step * k, p.idesc, (it | k) != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (it + 1 == p.k_iters) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;  // the two warps of a quadrant take alternate 32-column blocks
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles, mt = (tile / p.n_tiles) % p.m_tiles, b = tile / (p.n_tiles * p.m_tiles);
      const int m = mt * 128 + row, n0 = nt * p.BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * 256u;
      const int64_t coff = static_cast<int64_t>(b) * p.scb + static_cast<int64_t>(m) * p.ldc;
      int blk = 0;
      for (int c = 0; c < p.BN; c += 32, ++blk) {
        if ((blk & 1) != grp) continue;
        __syncwarp();
        uint32_t r[32];
        if (c + 32 <= p.BN) {
          tmem_ld32(taddr + static_cast<uint32_t>(c), r);
        } else {  // BN is a multiple of 16: last half block
          uint32_t r16[16];
          tmem_ld16(taddr + static_cast<uint32_t>(c), r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = r16[j];
        }
        tmem_ld_wait();
        const int ncols = c + 32 <= p.BN ? 32 : 16;
#pragma unroll
        for (int gp = 0; gp < 2; ++gp) {  // pairs of 8-column groups
          const int n = n0 + c + gp * 16;
          if (m >= p.M || gp * 16 >= ncols || n >= p.N) break;
          const bool second = n + 8 < p.N;  // N is a multiple of 8
          if (!p.out_bf16) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (hh == 1 && !second) break;
              float* dst = static_cast<float*>(p.C) + coff + n + hh * 8;
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(p.alpha * __uint_as_float(r[gp * 16 + hh * 8 + j]));
              if (p.c8) {  // 8 floats = one 32-byte sector
                st256(dst, w);
              } else {
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
              }
            }
          } else {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(p.alpha * __uint_as_float(r[gp * 16 + 2 * j]),
                                                             p.alpha * __uint_as_float(r[gp * 16 + 2 * j + 1]));
              w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C) + coff + n;
            if (p.c16 && second) {  // 16 bf16 = one sector
              st256(dst, w);
            } else {
              *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
              if (second) *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace icgan

using namespace icgan;

extern "C" int icgan_gemm_tc(const void* A, const void* B, void* C, int M, int N, int K, int batch, int a_mn, int b_mn,
                             int64_t lda, int64_t sab, int64_t ldb, int64_t sbb, int64_t ldc, int64_t scb, float alpha,
                             int c_dtype, void* stream) {
  ICGAN_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "icgan_gemm_tc: bad arguments");
  ICGAN_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && N % 8 == 0 && ldc % 8 == 0,
                "icgan_gemm_tc: leading dimensions and N must be multiples of 8 (lda %lld ldb %lld ldc %lld N %d)",
                (long long)lda, (long long)ldb, (long long)ldc, N);
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.a_mn = a_mn; p.b_mn = b_mn;
  if (N <= 256) p.BN = (N + 15) / 16 * 16;
  else if (N % 256 == 0) p.BN = 256;
  else if (N % 192 == 0) p.BN = 192;
  else if (N % 128 == 0) p.BN = 128;
  else p.BN = 256;
  p.b_boxes = (p.BN + 63) / 64;
  p.m_tiles = ceil_div(M, 128);
  p.n_tiles = ceil_div(N, p.BN);
  p.total_tiles = p.m_tiles * p.n_tiles * batch;
  p.k_iters = ceil_div(K, kKC);
  p.a_bytes = 16384u;
  p.b_tx = b_mn ? static_cast<uint32_t>(p.b_boxes) * 8192u : static_cast<uint32_t>(p.BN) * 128u;
  p.stage_bytes = p.a_bytes + ((p.b_tx + 1023u) & ~1023u);
  const uint32_t tail = 1024u + 512u;
  int stages = static_cast<int>((kSmemBudgetG - tail) / p.stage_bytes);
  if (stages > kStagesMax) stages = kStagesMax;
  ICGAN_REQUIRE(stages >= 2, "icgan_gemm_tc: tile does not fit shared memory");
  p.stages = stages;
  p.idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.BN)) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
  p.ldc = ldc; p.scb = scb; p.out_bf16 = c_dtype == ICGAN_BF16; p.alpha = alpha; p.C = C;
  p.c16 = (ldc % 16 == 0) && (scb % 16 == 0) && (reinterpret_cast<uintptr_t>(C) % 32 == 0);
  p.c8 = (scb % 8 == 0) && (reinterpret_cast<uintptr_t>(C) % 32 == 0);

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_map3(&tmA, A, K, M, batch, lda, sab, kKC, 128);
  else rc = make_map3(&tmA, A, M, K, batch, lda, sab, 64, kKC);
  if (rc) return rc;
  if (!b_mn) rc = make_map3(&tmB, B, K, N, batch, ldb, sbb, kKC, p.BN);
  else rc = make_map3(&tmB, B, N, K, batch, ldb, sbb, 64, kKC);
  if (rc) return rc;
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudgetG));
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  tc_gemm_kernel<<<grid, kThreadsG, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
