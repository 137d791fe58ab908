// Fused non-local block core (layers.Attention.forward, BigGAN_PyTorch/layers.py:233-243) on tcgen05:
//     beta = softmax_k(theta phi^T),  o = beta g          theta [B,Q,d], pooled phi [B,Kk,d], pooled g [B,Kk,dv]  (bf16)
// without the [B,Q,Kk] float32 logits ever leaving the SM.
//
//   attn_fwd_kernel    one CTA per 128 query rows.  The logits of a 128-key chunk are one UMMA (accumulator in TMEM); a
//                      softmax thread owns one query row (TMEM lane = row) and 64 of the chunk's columns, so row maxima
//                      and sums need no shuffles, only one exchange between the two column halves.  Pass 1 walks the
//                      Kk/128 chunks for the row maximum; pass 2 recomputes each chunk, writes exp2(s log2e - max) as
//                      bf16 straight into a SWIZZLE_128B K-major tile in shared memory and the MMA warp accumulates
//                      O += P g  from it (g is read as an MN-major operand from the very tile TMA wrote); the row sum
//                      divides O on the way out and gives the log-sum-exp the backward restarts from.
//   attn_bwd_q_kernel  same tiling for the query-side backward:  P recomputed from the saved log-sum-exp,
//                      dP = dO g^T (second TMEM accumulator), dS = P (dP - rowsum(dO o)) written as bf16 to shared memory
//                      (operand of  dtheta += dS phi, third accumulator; phi is the tile already loaded for the logits, read
//                      MN-major).  rowsum(dO o) comes from a coalesced pre-pass (attn_rowdot_kernel) that also feeds the key side.
//   attn_bwd_kv_kernel one CTA per 128 keys, walking the queries 64 at a time:  S^T = phi theta^T and dP^T = g dO^T
//                      (double-buffered TMEM accumulators, thread = key row, the per-query log-sum-exp and rowsum arrive
//                      with the stage as two 256-byte bulk copies), P^T and dS^T as bf16 K-major tiles in shared memory,
//                      dg += P^T dO and dphi += dS^T theta accumulate in TMEM over the whole walk (dO and theta are the
//                      stage's own tiles read MN-major).  Neither P nor dS ever exists in HBM.
//
// Warp roles in all three: warp 0 = TMA producer, warp 1 = MMA issuer (and TMEM owner), the rest = softmax / epilogue
// (warp w works on TMEM lanes 32 (w mod 4) ..).  Persistent over (sample, tile); every ring is driven by running counters
// so phases carry across tiles.
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace icgan {
namespace {

constexpr uint32_t kTile = 16384u;  // 128 rows x 128 bytes (64 bf16), one SWIZZLE_128B operand tile
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, Q, Kk, d, dv;
  int q_tiles, total_tiles, n_chunks, d_steps, v_boxes, dv_steps;
  uint32_t idesc_s, idesc_o, idesc_dq;
  const __nv_bfloat16* O_in;   // backward: forward output
  const __nv_bfloat16* dO;     // backward: incoming gradient
  __nv_bfloat16* O;            // forward output [B,Q,dv]
  __nv_bfloat16* P;            // forward: probabilities [B,Q,Kk] or nullptr;  backward: dS [B,Q,Kk]
  __nv_bfloat16* dTheta;       // backward [B,Q,d]
  float* lse2;                 // [B,Q] log2-domain log-sum-exp: written by the forward (nullable), read by the backward
  float* dsum;                 // [B,Q] rowsum(dO * o): written by the query-side backward (nullable), read by the key side
  __nv_bfloat16* dPhi;         // key-side backward [B,Kk,d]
  __nv_bfloat16* dG;           // key-side backward [B,Kk,dv]
  uint32_t idesc_s64;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Thread `row` of a 128-row K-major SWIZZLE_128B operand writes 32 consecutive K elements (columns col0 .. col0+31 of a
// 128-column chunk held as two 64-column tiles).  Element k of row r lives at  tile + r*128 + ((k/8 ^ (r & 7)) << 4).
__device__ __forceinline__ void store_row_block(uint32_t chunk_addr, int row, int blk, const uint32_t (&w)[16]) {
  const uint32_t tile = chunk_addr + static_cast<uint32_t>(blk >> 1) * kTile + static_cast<uint32_t>(row) * 128u;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint32_t cidx = static_cast<uint32_t>((blk & 1) * 4 + ch);
    st_shared_v4(tile + ((cidx ^ static_cast<uint32_t>(row & 7)) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2],
                 w[4 * ch + 3]);
  }
}

// ------------------------------------------------------------------------------------------------------------- forward
// shared memory: theta[2] | phi[2] | g[2] (3 tiles each) | P[2] (2 tiles each) | barriers
constexpr uint32_t kFwdTheta = 0, kFwdPhi = 2 * kTile, kFwdG = 4 * kTile, kFwdP = 10 * kTile, kFwdBars = 14 * kTile;
constexpr uint32_t kFwdXch = kFwdBars + 256u;  // 1.5 KB: row maxima / sums exchanged between the two column halves
constexpr uint32_t kFwdSmem = kFwdXch + 1536u + 1024u;
constexpr int kAttnThreads = 320;  // all three kernels: TMA warp + MMA warp + 8 softmax warps (two per TMEM lane quadrant)
static_assert(kFwdSmem <= 227u * 1024u, "forward tile set exceeds shared memory");
enum FwdBar { F_TH_FULL = 0, F_TH_EMPTY = 2, F_PH_FULL = 4, F_PH_EMPTY = 6, F_G_FULL = 8, F_G_EMPTY = 10, F_S_FULL = 12,
              F_S_EMPTY = 14, F_P_FULL = 16, F_P_EMPTY = 18, F_O_FULL = 20, F_O_EMPTY = 21, F_NBARS = 22 };

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmTheta, const __grid_constant__ CUtensorMap tmPhi,
                const __grid_constant__ CUtensorMap tmG, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFwdBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + F_NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < F_NBARS; ++i) {
      const bool eight = (i == F_S_EMPTY || i == F_S_EMPTY + 1 || i == F_P_FULL || i == F_P_FULL + 1);
      mbar_init(&bars[i], eight ? 8u : (i == F_O_EMPTY ? 4u : 1u));
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
  const int nc = p.n_chunks;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmTheta);
      tma_prefetch_desc(&tmPhi);
      tma_prefetch_desc(&tmG);
    }
    uint32_t tl = 0, fi = 0, gi = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const int b = tile / p.q_tiles, q0 = (tile % p.q_tiles) * 128;
      const uint32_t tb = tl & 1u;
      mbar_wait(&bars[F_TH_EMPTY + tb], ((tl >> 1) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_expect_tx(&bars[F_TH_FULL + tb], kTile);
        tma_load_3d(smem + kFwdTheta + tb * kTile, &tmTheta, &bars[F_TH_FULL + tb], 0, q0, b);
      }
      __syncwarp();
      // pass 1 (row maxima) needs phi only and g's ring would sit idle: its two 3-tile slots carry the phi chunks three
      // at a time, so up to six chunk loads are in flight instead of the phi ring's two
      for (int c0 = 0; c0 < nc; c0 += 3) {
        const uint32_t gs = gi & 1u;
        const int n = nc - c0 < 3 ? nc - c0 : 3;
        mbar_wait(&bars[F_G_EMPTY + gs], ((gi >> 1) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars[F_G_FULL + gs], static_cast<uint32_t>(n) * kTile);
          for (int j = 0; j < n; ++j)
            tma_load_3d(smem + kFwdG + (gs * 3u + j) * kTile, &tmPhi, &bars[F_G_FULL + gs], 0, (c0 + j) * 128, b);
        }
        __syncwarp();
        ++gi;
      }
      for (int c = 0; c < nc; ++c) {  // pass 2: phi chunk -> phi ring, g chunk -> g ring
        const uint32_t slot = fi & 1u;
        mbar_wait(&bars[F_PH_EMPTY + slot], ((fi >> 1) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars[F_PH_FULL + slot], kTile);
          tma_load_3d(smem + kFwdPhi + slot * kTile, &tmPhi, &bars[F_PH_FULL + slot], 0, c * 128, b);
        }
        __syncwarp();
        ++fi;
        const uint32_t gs = gi & 1u;
        mbar_wait(&bars[F_G_EMPTY + gs], ((gi >> 1) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars[F_G_FULL + gs], static_cast<uint32_t>(p.v_boxes) * kTile);
          for (int j = 0; j < p.v_boxes; ++j)
            tma_load_3d(smem + kFwdG + (gs * 3u + j) * kTile, &tmG, &bars[F_G_FULL + gs], j * 64, c * 128, b);
        }
        __syncwarp();
        ++gi;
      }
    }
  } else if (warp == 1) {
    uint32_t tl = 0, fi = 0, gi = 0, si = 0, pi = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const uint32_t tb = tl & 1u;
      mbar_wait(&bars[F_TH_FULL + tb], (tl >> 1) & 1u);
      const uint64_t da = umma_desc_kmajor(base + kFwdTheta + tb * kTile, 128);
      // O += P_j g_j for chunk j of this tile
      auto issue_pv = [&](int j, bool last) {
        const uint32_t gs = gi & 1u, pb = pi & 1u;
        mbar_wait(&bars[F_G_FULL + gs], (gi >> 1) & 1u);
        mbar_wait(&bars[F_P_FULL + pb], (pi >> 1) & 1u);
        if (j == 0) mbar_wait(&bars[F_O_EMPTY], (tl & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t db = desc_mn(base + kFwdG + gs * 3u * kTile, kTile);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t dp = umma_desc_kmajor(base + kFwdP + (pb * 2u + (ks >> 2)) * kTile, 128) + 2u * (ks & 3);
            umma_bf16(tmem_base + 256u, dp, db + 128u * ks, p.idesc_o, (j | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&bars[F_G_EMPTY + gs]);
          umma_commit(&bars[F_P_EMPTY + pb]);
          if (last) umma_commit(&bars[F_O_FULL]);
        }
        __syncwarp();
        ++gi;
        ++pi;
      };
      for (int c0 = 0; c0 < nc; c0 += 3) {  // pass 1: logits from the phi chunks parked in g's ring
        const uint32_t gs = gi & 1u;
        const int n = nc - c0 < 3 ? nc - c0 : 3;
        mbar_wait(&bars[F_G_FULL + gs], (gi >> 1) & 1u);
        for (int j = 0; j < n; ++j) {
          const uint32_t sb = si & 1u;
          mbar_wait(&bars[F_S_EMPTY + sb], ((si >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t db = umma_desc_kmajor(base + kFwdG + (gs * 3u + j) * kTile, 128);
            for (int k = 0; k < p.d_steps; ++k)
              umma_bf16(tmem_base + sb * 128u, da + 2u * k, db + 2u * k, p.idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&bars[F_S_FULL + sb]);
            if (j == n - 1) umma_commit(&bars[F_G_EMPTY + gs]);
          }
          __syncwarp();
          ++si;
        }
        ++gi;
      }
      for (int c = 0; c < nc; ++c) {  // pass 2
        const uint32_t slot = fi & 1u, sb = si & 1u;
        mbar_wait(&bars[F_PH_FULL + slot], (fi >> 1) & 1u);
        mbar_wait(&bars[F_S_EMPTY + sb], ((si >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t db = umma_desc_kmajor(base + kFwdPhi + slot * kTile, 128);
          for (int k = 0; k < p.d_steps; ++k)
            umma_bf16(tmem_base + sb * 128u, da + 2u * k, db + 2u * k, p.idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&bars[F_PH_EMPTY + slot]);
          umma_commit(&bars[F_S_FULL + sb]);
          if (c == nc - 1) umma_commit(&bars[F_TH_EMPTY + tb]);
        }
        __syncwarp();
        ++fi;
        ++si;
        if (c > 0) issue_pv(c - 1, false);
      }
      issue_pv(nc - 1, true);
    }
  } else {
    // eight softmax warps: warp w owns TMEM lanes 32 (w mod 4) .. and the 64-column half `hh` of every logits chunk
    const int q = warp & 3, hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t taddr = lane_base + static_cast<uint32_t>(hh * 64);
    float* xm = reinterpret_cast<float*>(smem + kFwdXch);  // [2][128] row maxima of the two halves
    float* xl = xm + 256;                                  // [128] row sums of the upper half
    uint32_t tl = 0, si = 0, pi = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const int b = tile / p.q_tiles, q0 = (tile % p.q_tiles) * 128;
      const int64_t grow = static_cast<int64_t>(b) * p.Q + q0 + row;
      float m = -INFINITY;
      for (int c = 0; c < nc; ++c) {  // pass 1: row maximum
        const uint32_t sb = si & 1u;
        mbar_wait(&bars[F_S_FULL + sb], (si >> 1) & 1u);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + sb * 128u, r0);
        tmem_ld32(taddr + sb * 128u + 32u, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[F_S_EMPTY + sb]);  // the values are in registers
        float m0 = m, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          m0 = fmaxf(m0, __uint_as_float(r0[j]));
          m1 = fmaxf(m1, __uint_as_float(r1[j]));
        }
        m = fmaxf(m0, m1);
        ++si;
      }
      xm[hh * 128 + row] = m;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      m = fmaxf(m, xm[(hh ^ 1) * 128 + row]) * kLog2e;
      float l = 0.f;
      for (int c = 0; c < nc; ++c) {  // pass 2: exp2(s log2e - m) -> bf16 operand tile; row sum in float32
        const uint32_t sb = si & 1u, pb = pi & 1u;
        mbar_wait(&bars[F_S_FULL + sb], (si >> 1) & 1u);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + sb * 128u, r0);
        tmem_ld32(taddr + sb * 128u + 32u, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[F_S_EMPTY + sb]);
        uint32_t w0[16], w1[16];
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a0 = ex2(fmaf(__uint_as_float(r0[2 * j]), kLog2e, -m));
          const float a1 = ex2(fmaf(__uint_as_float(r0[2 * j + 1]), kLog2e, -m));
          const float b0 = ex2(fmaf(__uint_as_float(r1[2 * j]), kLog2e, -m));
          const float b1 = ex2(fmaf(__uint_as_float(r1[2 * j + 1]), kLog2e, -m));
          l0 += a0 + a1;
          l1 += b0 + b1;
          w0[j] = pack_bf16(a0, a1);
          w1[j] = pack_bf16(b0, b1);
        }
        l += l0 + l1;
        mbar_wait(&bars[F_P_EMPTY + pb], ((pi >> 1) & 1u) ^ 1u);
        store_row_block(base + kFwdP + pb * 2u * kTile, row, hh * 2, w0);
        store_row_block(base + kFwdP + pb * 2u * kTile, row, hh * 2 + 1, w1);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[F_P_FULL + pb]);
        ++si;
        ++pi;
      }
      if (hh == 1) {
        xl[row] = l;
        asm volatile("bar.arrive 2, 256;" ::: "memory");
      } else {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        l += xl[row];
        if (p.lse2) p.lse2[grow] = m + log2f(l);
        const float inv = 1.0f / l;
        mbar_wait(&bars[F_O_FULL], tl & 1u);
        tc_fence_after();
        __nv_bfloat16* orow = p.O + grow * p.dv;
        for (int c0 = 0; c0 < p.dv; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + 256u + static_cast<uint32_t>(c0), r);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            w[j] = pack_bf16(inv * __uint_as_float(r[2 * j]), inv * __uint_as_float(r[2 * j + 1]));
          uint4* dst = reinterpret_cast<uint4*>(orow + c0);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[F_O_EMPTY]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dsum[r] = sum_c a[r][c] * b[r][c]: one warp per row, 16-byte loads, rows of at most 256 bf16 (dv <= 192 here).
__global__ void __launch_bounds__(256) attn_rowdot_kernel(const __nv_bfloat16* __restrict__ a,
                                                           const __nv_bfloat16* __restrict__ b, float* __restrict__ out,
                                                           int64_t rows, int dv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float acc = 0.f;
    if (lane * 8 < dv) {
      const uint4 x = *reinterpret_cast<const uint4*>(a + r * dv + lane * 8);
      const uint4 y = *reinterpret_cast<const uint4*>(b + r * dv + lane * 8);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[e]));
        const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[e]));
        acc = fmaf(fx.x, fy.x, acc);
        acc = fmaf(fx.y, fy.y, acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = acc;
  }
}

// ------------------------------------------------------------------------------------------- backward, query side
// shared memory: theta | dO (3 tiles) | phi[2] | g[2] (3 tiles each) | dS (2 tiles) | barriers
constexpr uint32_t kBwdTheta = 0, kBwdDO = kTile, kBwdPhi = 4 * kTile, kBwdG = 6 * kTile, kBwdDS = 12 * kTile,
                   kBwdBars = 14 * kTile;
constexpr uint32_t kBwdSmem = kBwdBars + 256u + 1024u;
enum BwdBar { Q_TD_FULL = 0, Q_TD_EMPTY = 1, Q_PH_FULL = 2, Q_PH_EMPTY = 4, Q_G_FULL = 6, Q_G_EMPTY = 8, Q_S_FULL = 10,
              Q_S_EMPTY = 12, Q_DP_FULL = 14, Q_DP_EMPTY = 15, Q_DS_FULL = 16, Q_DS_EMPTY = 17, Q_DT_FULL = 18,
              Q_DT_EMPTY = 19, Q_NBARS = 20 };
// TMEM columns: logits [0,128) and [128,256), dP [256,384), dtheta [384,448)

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_q_kernel(const __grid_constant__ CUtensorMap tmTheta, const __grid_constant__ CUtensorMap tmPhi,
                  const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmDO,
                  const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Q_NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < Q_NBARS; ++i) {
      const bool eight = (i == Q_S_EMPTY || i == Q_S_EMPTY + 1 || i == Q_DP_EMPTY || i == Q_DS_FULL);
      mbar_init(&bars[i], eight ? 8u : (i == Q_DT_EMPTY ? 4u : 1u));
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
  const int nc = p.n_chunks;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmTheta);
      tma_prefetch_desc(&tmPhi);
      tma_prefetch_desc(&tmG);
      tma_prefetch_desc(&tmDO);
    }
    uint32_t tl = 0, fi = 0, gi = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const int b = tile / p.q_tiles, q0 = (tile % p.q_tiles) * 128;
      mbar_wait(&bars[Q_TD_EMPTY], (tl & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_expect_tx(&bars[Q_TD_FULL], static_cast<uint32_t>(1 + p.v_boxes) * kTile);
        tma_load_3d(smem + kBwdTheta, &tmTheta, &bars[Q_TD_FULL], 0, q0, b);
        for (int j = 0; j < p.v_boxes; ++j)
          tma_load_3d(smem + kBwdDO + j * kTile, &tmDO, &bars[Q_TD_FULL], j * 64, q0, b);
      }
      __syncwarp();
      for (int c = 0; c < nc; ++c) {
        const uint32_t slot = fi & 1u;
        mbar_wait(&bars[Q_PH_EMPTY + slot], ((fi >> 1) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars[Q_PH_FULL + slot], kTile);
          tma_load_3d(smem + kBwdPhi + slot * kTile, &tmPhi, &bars[Q_PH_FULL + slot], 0, c * 128, b);
        }
        __syncwarp();
        ++fi;
        const uint32_t gs = gi & 1u;
        mbar_wait(&bars[Q_G_EMPTY + gs], ((gi >> 1) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars[Q_G_FULL + gs], static_cast<uint32_t>(p.v_boxes) * kTile);
          for (int j = 0; j < p.v_boxes; ++j)
            tma_load_3d(smem + kBwdG + (gs * 3u + j) * kTile, &tmG, &bars[Q_G_FULL + gs], j * 64, c * 128, b);
        }
        __syncwarp();
        ++gi;
      }
    }
  } else if (warp == 1) {
    uint32_t tl = 0, fi = 0, gi = 0, si = 0, di = 0, fq = 0;  // fq: phi ring position of the next dtheta product
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      mbar_wait(&bars[Q_TD_FULL], tl & 1u);
      const uint64_t da_theta = umma_desc_kmajor(base + kBwdTheta, 128);
      // dtheta += dS_j phi_j
      auto issue_dq = [&](int j, bool last) {
        const uint32_t slot = fq & 1u;
        mbar_wait(&bars[Q_DS_FULL], di & 1u);
        if (j == 0) mbar_wait(&bars[Q_DT_EMPTY], (tl & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t db = desc_mn(base + kBwdPhi + slot * kTile, kTile);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t dsd = umma_desc_kmajor(base + kBwdDS + static_cast<uint32_t>(ks >> 2) * kTile, 128) +
                                 2u * (ks & 3);
            umma_bf16(tmem_base + 384u, dsd, db + 128u * ks, p.idesc_dq, (j | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&bars[Q_DS_EMPTY]);
          umma_commit(&bars[Q_PH_EMPTY + slot]);
          if (last) umma_commit(&bars[Q_DT_FULL]);
        }
        __syncwarp();
        ++di;
        ++fq;
      };
      for (int c = 0; c < nc; ++c) {
        const uint32_t slot = fi & 1u, sb = si & 1u;
        mbar_wait(&bars[Q_PH_FULL + slot], (fi >> 1) & 1u);
        mbar_wait(&bars[Q_S_EMPTY + sb], ((si >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t db = umma_desc_kmajor(base + kBwdPhi + slot * kTile, 128);
          for (int k = 0; k < p.d_steps; ++k)
            umma_bf16(tmem_base + sb * 128u, da_theta + 2u * k, db + 2u * k, p.idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&bars[Q_S_FULL + sb]);
        }
        __syncwarp();
        ++fi;
        ++si;
        const uint32_t gs = gi & 1u;
        mbar_wait(&bars[Q_G_FULL + gs], (gi >> 1) & 1u);
        mbar_wait(&bars[Q_DP_EMPTY], (gi & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int ks = 0; ks < p.dv_steps; ++ks) {
            const uint32_t off = static_cast<uint32_t>(ks >> 2) * kTile;
            const uint64_t da = umma_desc_kmajor(base + kBwdDO + off, 128) + 2u * (ks & 3);
            const uint64_t db = umma_desc_kmajor(base + kBwdG + gs * 3u * kTile + off, 128) + 2u * (ks & 3);
            umma_bf16(tmem_base + 256u, da, db, p.idesc_s, ks != 0 ? 1u : 0u);
          }
          umma_commit(&bars[Q_G_EMPTY + gs]);
          umma_commit(&bars[Q_DP_FULL]);
          if (c == nc - 1) umma_commit(&bars[Q_TD_EMPTY]);  // theta and dO have no reader after this product
        }
        __syncwarp();
        ++gi;
        if (c > 0) issue_dq(c - 1, false);
      }
      issue_dq(nc - 1, true);
    }
  } else {
    // eight warps: warp w owns TMEM lanes 32 (w mod 4) .. and the 64-column half `hh` of every chunk
    const int q = warp & 3, hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t taddr = lane_base + static_cast<uint32_t>(hh * 64);
    uint32_t tl = 0, si = 0, gi = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const int b = tile / p.q_tiles, q0 = (tile % p.q_tiles) * 128;
      const int64_t grow = static_cast<int64_t>(b) * p.Q + q0 + row;
      const float lse = p.lse2[grow];
      const float dsum = p.dsum[grow];  // rowsum(dO * o) = rowsum(dP * P), from attn_rowdot_kernel
      __nv_bfloat16* dsrow = p.P ? p.P + grow * p.Kk + hh * 64 : nullptr;
      for (int c = 0; c < nc; ++c) {
        const uint32_t sb = si & 1u;
        mbar_wait(&bars[Q_S_FULL + sb], (si >> 1) & 1u);
        mbar_wait(&bars[Q_DP_FULL], gi & 1u);
        tc_fence_after();
        uint32_t w[2][16];
        uint32_t g0[32], g1[32];  // dP first: it is single-buffered, and the next chunk's product waits for this read
        tmem_ld32(taddr + 256u, g0);
        tmem_ld32(taddr + 256u + 32u, g1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[Q_DP_EMPTY]);
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          uint32_t r[32];
          tmem_ld32(taddr + sb * 128u + static_cast<uint32_t>(b2 * 32), r);
          tmem_ld_wait();
          if (b2 == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[Q_S_EMPTY + sb]);
          }
          const uint32_t(&g)[32] = b2 == 0 ? g0 : g1;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ex2(fmaf(__uint_as_float(r[2 * j]), kLog2e, -lse));
            const float p1 = ex2(fmaf(__uint_as_float(r[2 * j + 1]), kLog2e, -lse));
            w[b2][j] = pack_bf16(p0 * (__uint_as_float(g[2 * j]) - dsum), p1 * (__uint_as_float(g[2 * j + 1]) - dsum));
          }
        }
        mbar_wait(&bars[Q_DS_EMPTY], (gi & 1u) ^ 1u);  // dtheta of the previous chunk has read the operand tile
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          store_row_block(base + kBwdDS, row, hh * 2 + b2, w[b2]);
          if (dsrow) {
            uint4* dst = reinterpret_cast<uint4*>(dsrow + c * 128 + b2 * 32);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              dst[ch] = make_uint4(w[b2][4 * ch], w[b2][4 * ch + 1], w[b2][4 * ch + 2], w[b2][4 * ch + 3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[Q_DS_FULL]);
        ++si;
        ++gi;
      }
      if (hh == 0) {
        mbar_wait(&bars[Q_DT_FULL], tl & 1u);
        tc_fence_after();
        __nv_bfloat16* trow = p.dTheta + grow * p.d;
        for (int c0 = 0; c0 < p.d; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + 384u + static_cast<uint32_t>(c0), r);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          uint4* dst = reinterpret_cast<uint4*>(trow + c0);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          if (c0 + 8 < p.d) dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[Q_DT_EMPTY]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------- backward, key side
// shared memory: phi_t | g_t (3 tiles) | stage[4] = {theta_h 8 KB, dO_h 3 x 8 KB} | P^T | dS^T |
//                stage rows[4] = {lse 256 B, dsum 256 B} | barriers
constexpr uint32_t kHalf = 8192u;  // 64 rows x 128 bytes
constexpr uint32_t kKvSlots = 4;   // the walk is latency-bound on these loads: each extra stage in flight pays directly
constexpr uint32_t kKvPhi = 0, kKvG = kTile, kKvStage = 4 * kTile, kKvStageBytes = 4 * kHalf,
                   kKvPT = kKvStage + kKvSlots * kKvStageBytes, kKvDST = kKvPT + kTile, kKvRows = kKvDST + kTile,
                   kKvBars = kKvRows + kKvSlots * 512u;
// four 32 KB stages leave 864 bytes for re-aligning the dynamic window to 1024; it starts right after the 1 KB the
// system reserves, i.e. aligned, and the kernel traps if a toolchain ever places it otherwise
constexpr uint32_t kKvSmem = 227u * 1024u, kKvAlignSlack = kKvSmem - (kKvBars + 160u);
static_assert(kKvPT % 1024u == 0, "operand tiles must stay 1024-byte aligned");
static_assert(kKvBars + 160u <= kKvSmem, "key-side tile set exceeds shared memory");
enum KvBar { K_KV_FULL = 0, K_KV_EMPTY = 1, K_ST_FULL = 2, K_ST_EMPTY = 6, K_SD_FULL = 10, K_SD_EMPTY = 12, K_PD_FULL = 14,
             K_PD_EMPTY = 15, K_ACC_FULL = 16, K_ACC_EMPTY = 17, K_NBARS = 18 };
// TMEM columns: S^T [0,64) [64,128), dP^T [128,192) [192,256), dg [256,448), dphi [448,512)

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_kv_kernel(const __grid_constant__ CUtensorMap tmThetaH, const __grid_constant__ CUtensorMap tmPhi,
                   const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmDOH,
                   const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  if (base - raw > kKvAlignSlack) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kKvBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + K_NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < K_NBARS; ++i) {
      const bool eight = (i == K_SD_EMPTY || i == K_SD_EMPTY + 1 || i == K_PD_FULL || i == K_ACC_EMPTY);
      mbar_init(&bars[i], eight ? 8u : 1u);
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
  const int ni = p.Q / 64;           // query half-tiles per walk
  const int k_tiles = p.n_chunks;    // key tiles per sample
  const int total = p.B * k_tiles;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmThetaH);
      tma_prefetch_desc(&tmPhi);
      tma_prefetch_desc(&tmG);
      tma_prefetch_desc(&tmDOH);
    }
    uint32_t tl = 0, slot = 0, sph = 0;  // stage ring position and phase
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tl) {
      const int b = tile / k_tiles, k0 = (tile % k_tiles) * 128;
      mbar_wait(&bars[K_KV_EMPTY], (tl & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_expect_tx(&bars[K_KV_FULL], static_cast<uint32_t>(1 + p.v_boxes) * kTile);
        tma_load_3d(smem + kKvPhi, &tmPhi, &bars[K_KV_FULL], 0, k0, b);
        for (int j = 0; j < p.v_boxes; ++j) tma_load_3d(smem + kKvG + j * kTile, &tmG, &bars[K_KV_FULL], j * 64, k0, b);
      }
      __syncwarp();
      for (int i = 0; i < ni; ++i) {
        mbar_wait(&bars[K_ST_EMPTY + slot], sph ^ 1u);
        if (elect_one_sync()) {
          uint8_t* st = smem + kKvStage + slot * kKvStageBytes;
          uint64_t* bar = &bars[K_ST_FULL + slot];
          mbar_expect_tx(bar, static_cast<uint32_t>(1 + p.v_boxes) * kHalf + 512u);
          tma_load_3d(st, &tmThetaH, bar, 0, i * 64, b);
          for (int j = 0; j < p.v_boxes; ++j) tma_load_3d(st + (1 + j) * kHalf, &tmDOH, bar, j * 64, i * 64, b);
          const int64_t qoff = static_cast<int64_t>(b) * p.Q + i * 64;
          bulk_load_1d(smem + kKvRows + slot * 512u, p.lse2 + qoff, 256u, bar);
          bulk_load_1d(smem + kKvRows + slot * 512u + 256u, p.dsum + qoff, 256u, bar);
        }
        __syncwarp();
        if (++slot == kKvSlots) {
          slot = 0;
          sph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    uint32_t tl = 0, it = 0, ia = 0;  // it: next S^T/dP^T product, ia: next accumulation (both count query half-tiles)
    uint32_t slot = 0, sph = 0, aslot = 0;  // stage ring position / phase of `it`, position of `ia`
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tl) {
      mbar_wait(&bars[K_KV_FULL], tl & 1u);
      const uint64_t da_phi = umma_desc_kmajor(base + kKvPhi, 128);
      // dg += P^T_j dO_j,  dphi += dS^T_j theta_j  for query half-tile j of this walk
      auto issue_acc = [&](int j, bool last) {
        mbar_wait(&bars[K_PD_FULL], ia & 1u);
        if (j == 0) mbar_wait(&bars[K_ACC_EMPTY], (tl & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t st = base + kKvStage + aslot * kKvStageBytes;
          const uint64_t db_do = desc_mn(st + kHalf, kHalf), db_th = desc_mn(st, kHalf);
          const uint64_t da_p = umma_desc_kmajor(base + kKvPT, 128);
          const uint64_t da_ds = umma_desc_kmajor(base + kKvDST, 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + 256u, da_p + 2u * ks, db_do + 128u * ks, p.idesc_o, (j | ks) != 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + 448u, da_ds + 2u * ks, db_th + 128u * ks, p.idesc_dq, (j | ks) != 0 ? 1u : 0u);
          umma_commit(&bars[K_PD_EMPTY]);
          umma_commit(&bars[K_ST_EMPTY + aslot]);
          if (last) umma_commit(&bars[K_ACC_FULL]);
        }
        __syncwarp();
        ++ia;
        if (++aslot == kKvSlots) aslot = 0;
      };
      for (int i = 0; i < ni; ++i, ++it) {
        const uint32_t sb = it & 1u;
        mbar_wait(&bars[K_ST_FULL + slot], sph);
        mbar_wait(&bars[K_SD_EMPTY + sb], ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t st = base + kKvStage + slot * kKvStageBytes;
          const uint64_t db_th = umma_desc_kmajor(st, 128);
          for (int k = 0; k < p.d_steps; ++k)
            umma_bf16(tmem_base + sb * 64u, da_phi + 2u * k, db_th + 2u * k, p.idesc_s64, k != 0 ? 1u : 0u);
          for (int ks = 0; ks < p.dv_steps; ++ks) {
            const uint64_t da = umma_desc_kmajor(base + kKvG + static_cast<uint32_t>(ks >> 2) * kTile, 128) + 2u * (ks & 3);
            const uint64_t db = umma_desc_kmajor(st + kHalf + static_cast<uint32_t>(ks >> 2) * kHalf, 128) + 2u * (ks & 3);
            umma_bf16(tmem_base + 128u + sb * 64u, da, db, p.idesc_s64, ks != 0 ? 1u : 0u);
          }
          umma_commit(&bars[K_SD_FULL + sb]);
          if (i == ni - 1) umma_commit(&bars[K_KV_EMPTY]);  // phi_t and g_t have no reader after these products
        }
        __syncwarp();
        if (++slot == kKvSlots) {
          slot = 0;
          sph ^= 1u;
        }
        if (i > 0) issue_acc(i - 1, false);
      }
      issue_acc(ni - 1, true);
    }
  } else {
    // eight warps: warp w owns TMEM lanes (= key rows) 32 (w mod 4) .. and the 32-query block `hh` of every half-tile
    const int q = warp & 3, hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t taddr = lane_base + static_cast<uint32_t>(hh * 32);
    uint32_t tl = 0, it = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tl) {
      const int b = tile / k_tiles, k0 = (tile % k_tiles) * 128;
      for (int i = 0; i < ni; ++i, ++it) {
        const uint32_t sb = it & 1u;
        mbar_wait(&bars[K_ST_FULL + slot], sph);   // lse / dsum of these 64 queries
        mbar_wait(&bars[K_SD_FULL + sb], (it >> 1) & 1u);
        tc_fence_after();
        const float4* lse4 = reinterpret_cast<const float4*>(smem + kKvRows + slot * 512u) + hh * 8;
        const float4* dsm4 = lse4 + 16;
        uint32_t wp[16], wd[16];
        {
          uint32_t r[32], g[32];
          tmem_ld32(taddr + sb * 64u, r);
          tmem_ld32(taddr + 128u + sb * 64u, g);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[K_SD_EMPTY + sb]);  // both accumulators are in registers now
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 ls = lse4[j4], dm = dsm4[j4];
            const float p0 = ex2(fmaf(__uint_as_float(r[4 * j4]), kLog2e, -ls.x));
            const float p1 = ex2(fmaf(__uint_as_float(r[4 * j4 + 1]), kLog2e, -ls.y));
            const float p2 = ex2(fmaf(__uint_as_float(r[4 * j4 + 2]), kLog2e, -ls.z));
            const float p3 = ex2(fmaf(__uint_as_float(r[4 * j4 + 3]), kLog2e, -ls.w));
            wp[2 * j4] = pack_bf16(p0, p1);
            wp[2 * j4 + 1] = pack_bf16(p2, p3);
            wd[2 * j4] = pack_bf16(p0 * (__uint_as_float(g[4 * j4]) - dm.x), p1 * (__uint_as_float(g[4 * j4 + 1]) - dm.y));
            wd[2 * j4 + 1] =
                pack_bf16(p2 * (__uint_as_float(g[4 * j4 + 2]) - dm.z), p3 * (__uint_as_float(g[4 * j4 + 3]) - dm.w));
          }
        }
        mbar_wait(&bars[K_PD_EMPTY], (it & 1u) ^ 1u);        // the previous accumulation has read the operand tiles
        store_row_block(base + kKvPT, row, hh, wp);
        store_row_block(base + kKvDST, row, hh, wd);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[K_PD_FULL]);
        if (++slot == kKvSlots) {
          slot = 0;
          sph ^= 1u;
        }
      }
      mbar_wait(&bars[K_ACC_FULL], tl & 1u);
      tc_fence_after();
      const int64_t krow = static_cast<int64_t>(b) * p.Kk + k0 + row;
      __nv_bfloat16* grow_ = p.dG + krow * p.dv;
      const int units = p.dv / 16, u0 = hh == 0 ? 0 : (units + 1) / 2, u1 = hh == 0 ? (units + 1) / 2 : units;
      for (int u = u0; u < u1; ++u) {
        uint32_t r[16];
        tmem_ld16(lane_base + 256u + static_cast<uint32_t>(u * 16), r);
        tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        uint4* dst = reinterpret_cast<uint4*>(grow_ + u * 16);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
      if (hh == 1) {
        __nv_bfloat16* prow = p.dPhi + krow * p.d;
        for (int c0 = 0; c0 < p.d; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + 448u + static_cast<uint32_t>(c0), r);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          uint4* dst = reinterpret_cast<uint4*>(prow + c0);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          if (c0 + 8 < p.d) dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[K_ACC_EMPTY]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int fill_params(AttnParams* p, int B, int Q, int Kk, int d, int dv, const char* who) {
  ICGAN_REQUIRE(B > 0 && Q > 0 && Kk > 0 && Q % 128 == 0 && Kk % 128 == 0,
                "%s: Q (%d) and Kk (%d) must be multiples of 128", who, Q, Kk);
  ICGAN_REQUIRE(d >= 8 && d <= 64 && d % 8 == 0, "%s: d = %d must be a multiple of 8 in [8, 64]", who, d);
  ICGAN_REQUIRE(dv >= 16 && dv <= 192 && dv % 16 == 0, "%s: dv = %d must be a multiple of 16 in [16, 192]", who, dv);
  p->B = B; p->Q = Q; p->Kk = Kk; p->d = d; p->dv = dv;
  p->q_tiles = Q / 128;
  p->total_tiles = B * p->q_tiles;
  p->n_chunks = Kk / 128;
  p->d_steps = (d + 15) / 16;
  p->dv_steps = dv / 16;
  p->v_boxes = (dv + 63) / 64;
  p->idesc_s = umma_idesc_bf16(128, 128);
  p->idesc_o = umma_idesc_bf16(128, static_cast<uint32_t>(dv)) | (1u << 16);
  p->idesc_dq = umma_idesc_bf16(128, static_cast<uint32_t>(p->d_steps * 16)) | (1u << 16);
  p->idesc_s64 = umma_idesc_bf16(128, 64);
  return 0;
}

}  // namespace
}  // namespace icgan

using namespace icgan;

extern "C" int icgan_attn_fwd(const void* theta, const void* phi, const void* g, void* o, float* lse2, int B, int Q,
                              int Kk, int d, int dv, void* stream) {
  ICGAN_REQUIRE(theta && phi && g && o, "icgan_attn_fwd: null pointer");
  AttnParams p{};
  if (int rc = fill_params(&p, B, Q, Kk, d, dv, "icgan_attn_fwd")) return rc;
  p.O = static_cast<__nv_bfloat16*>(o);
  p.lse2 = lse2;
  CUtensorMap tmT, tmP, tmG;
  if (int rc = make_map3(&tmT, theta, d, Q, B, d, static_cast<uint64_t>(Q) * d, 64, 128)) return rc;
  if (int rc = make_map3(&tmP, phi, d, Kk, B, d, static_cast<uint64_t>(Kk) * d, 64, 128)) return rc;
  if (int rc = make_map3(&tmG, g, dv, Kk, B, dv, static_cast<uint64_t>(Kk) * dv, 64, 128)) return rc;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured))
    ICGAN_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  attn_fwd_kernel<<<grid, kAttnThreads, kFwdSmem, static_cast<cudaStream_t>(stream)>>>(tmT, tmP, tmG, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_attn_bwd_q(const void* theta, const void* phi, const void* g, const void* o, const void* dout,
                                const float* lse2, void* dtheta, void* ds, float* dsum, int B, int Q, int Kk, int d,
                                int dv, void* stream) {
  ICGAN_REQUIRE(theta && phi && g && o && dout && lse2 && dtheta && dsum, "icgan_attn_bwd_q: null pointer");
  AttnParams p{};
  if (int rc = fill_params(&p, B, Q, Kk, d, dv, "icgan_attn_bwd_q")) return rc;
  p.O_in = static_cast<const __nv_bfloat16*>(o);
  p.dO = static_cast<const __nv_bfloat16*>(dout);
  p.P = static_cast<__nv_bfloat16*>(ds);
  p.dTheta = static_cast<__nv_bfloat16*>(dtheta);
  p.lse2 = const_cast<float*>(lse2);
  p.dsum = dsum;
  CUtensorMap tmT, tmP, tmG, tmD;
  if (int rc = make_map3(&tmT, theta, d, Q, B, d, static_cast<uint64_t>(Q) * d, 64, 128)) return rc;
  if (int rc = make_map3(&tmP, phi, d, Kk, B, d, static_cast<uint64_t>(Kk) * d, 64, 128)) return rc;
  if (int rc = make_map3(&tmG, g, dv, Kk, B, dv, static_cast<uint64_t>(Kk) * dv, 64, 128)) return rc;
  if (int rc = make_map3(&tmD, dout, dv, Q, B, dv, static_cast<uint64_t>(Q) * dv, 64, 128)) return rc;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured))
    ICGAN_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  const int64_t rows = static_cast<int64_t>(B) * Q;
  const int dot_blocks = static_cast<int>(rows / 8 < 8 * num_sms() ? (rows + 7) / 8 : 8 * num_sms());
  attn_rowdot_kernel<<<dot_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p.dO, p.O_in, dsum, rows, dv);
  attn_bwd_q_kernel<<<grid, kAttnThreads, kBwdSmem, static_cast<cudaStream_t>(stream)>>>(tmT, tmP, tmG, tmD, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_attn_bwd_kv(const void* theta, const void* phi, const void* g, const void* dout, const float* lse2,
                                 const float* dsum, void* dphi, void* dg, int B, int Q, int Kk, int d, int dv,
                                 void* stream) {
  ICGAN_REQUIRE(theta && phi && g && dout && lse2 && dsum && dphi && dg, "icgan_attn_bwd_kv: null pointer");
  AttnParams p{};
  if (int rc = fill_params(&p, B, Q, Kk, d, dv, "icgan_attn_bwd_kv")) return rc;
  p.lse2 = const_cast<float*>(lse2);
  p.dsum = const_cast<float*>(dsum);
  p.dPhi = static_cast<__nv_bfloat16*>(dphi);
  p.dG = static_cast<__nv_bfloat16*>(dg);
  CUtensorMap tmT, tmP, tmG, tmD;
  if (int rc = make_map3(&tmT, theta, d, Q, B, d, static_cast<uint64_t>(Q) * d, 64, 64)) return rc;
  if (int rc = make_map3(&tmP, phi, d, Kk, B, d, static_cast<uint64_t>(Kk) * d, 64, 128)) return rc;
  if (int rc = make_map3(&tmG, g, dv, Kk, B, dv, static_cast<uint64_t>(Kk) * dv, 64, 128)) return rc;
  if (int rc = make_map3(&tmD, dout, dv, Q, B, dv, static_cast<uint64_t>(Q) * dv, 64, 64)) return rc;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured))
    ICGAN_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKvSmem));
  const int total = B * p.n_chunks;
  const int grid = total < num_sms() ? total : num_sms();
  attn_bwd_kv_kernel<<<grid, kAttnThreads, kKvSmem, static_cast<cudaStream_t>(stream)>>>(tmT, tmP, tmG, tmD, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
