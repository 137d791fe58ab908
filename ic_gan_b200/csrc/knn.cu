// Exact k-nearest-neighbour build over the stored instance features (replaces the Faiss IndexFlatL2 add/search in
// ILSVRC_HDF5_feats._obtain_nns, data_utils/datasets_common.py:695-745, driven by data_utils/make_hdf5_nns.py:97-172).
//
// Stage 1 (tensor cores, never materialises the N x N matrix): d~2(i,j) = |x_i|^2 + |x_j|^2 - 2 x_i.x_j with the dot
//   product from tcgen05 UMMAs over a split-bf16 representation x = hi + lo (three products hi.hi + hi.lo + lo.hi =>
//   ~2^-16 relative accuracy instead of bf16's 2^-8).  A persistent CTA owns a tile of 128 query rows, sweeps ALL
//   database columns, and its four epilogue warps (one thread per query row) keep a sorted list of the C smallest
//   distances in shared memory while the next column tile is being multiplied (double-buffered TMEM accumulators).
// Stage 2 (exact): every candidate distance is recomputed in float64 from the float32 features,
//   sum_k (double(x_ik) - double(x_jk))^2, candidates are sorted by (distance, index) -- ties to the lower index -- the
//   row's own index is dropped by VALUE (datasets_common.py:739-743) and the first k survive.  The radius is the
//   (k+1)-th distance of the un-pruned list (make_hdf5_nns.py:133), float32 square root like Faiss' float32 output.
// Stage 3 (proof of exactness): a row is certified iff its exact (k+1)-th distance lies below the smallest coarse
//   distance that was NOT kept, minus twice the largest coarse-vs-exact deviation observed; uncertified rows are
//   recomputed by brute force in float64 (icgan_knn_exact_row).  No CPU fallback anywhere.
#include <cuda.h>
#include <float.h>
#include <limits.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace icgan {
namespace {

constexpr int kKnnThreads = 192;
constexpr int kKnnBN = 256;   // database columns per tile
constexpr int kKnnKC = 64;    // reduction elements per pipeline stage
constexpr int kKnnMaxC = 64;  // candidates kept per query row
constexpr uint32_t kKnnSmem = 227u * 1024u;

int make_map2(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  ICGAN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t str[1] = {cols * 2};
  cuuint32_t box[2] = {kKnnKC, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, str, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ICGAN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}

// x -> bf16 hi, bf16 lo (= x - hi), |x|^2
__global__ void knn_prepare_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo, float* __restrict__ norms, int64_t N, int d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int k = lane; k < d; k += 32) {
    const float v = X[row * d + k];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[row * d + k] = h;
    lo[row * d + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    s += static_cast<double>(v) * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norms[row] = static_cast<float>(s);
}

struct KnnParams {
  int64_t N, q_begin, q_end;
  int d, C, passes, k_iters, col_tiles, row_tiles, stages;
  uint32_t stage_bytes;
  const float* norms;
  int* cand_idx;
  float* cand_d;
};

__global__ void __launch_bounds__(kKnnThreads, 1)
knn_coarse_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                  const KnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* tailp = smem + static_cast<size_t>(p.stages) * p.stage_bytes;
  float* cd = reinterpret_cast<float*>(tailp);                 // [C][128] sorted candidate distances
  int* ci = reinterpret_cast<int*>(tailp + kKnnMaxC * 128 * 4);  // [C][128] candidate indices
  uint64_t* full = reinterpret_cast<uint64_t*>(tailp + 2 * kKnnMaxC * 128 * 4);
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;
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
      mbar_init(&tempty[i], 4);
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
  const uint32_t a_bytes = 128u * 128u, b_bytes = kKnnBN * 128u;

  if (warp == 0) {
    {  // whole warp runs the loops (uniform control flow); one elected lane issues the TMA / MMA instructions
      int stage = 0;
      uint32_t phase = 0;
      for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
        const int m0 = static_cast<int>(p.q_begin) + rt * 128;
        for (int ct = 0; ct < p.col_tiles; ++ct) {
          const int n0 = ct * kKnnBN;
          for (int kc = 0; kc < p.k_iters; ++kc) {
            for (int ps = 0; ps < p.passes; ++ps) {  // hi.hi, hi.lo, lo.hi
              mbar_wait(&empty[stage], phase ^ 1u);
              uint8_t* sa = smem + static_cast<size_t>(stage) * p.stage_bytes;
              if (elect_one_sync()) {
                mbar_expect_tx(&full[stage], a_bytes + b_bytes);
                tma_load_2d(sa, ps == 2 ? &tmAlo : &tmAhi, &full[stage], kc * kKnnKC, m0);
                tma_load_2d(sa + a_bytes, ps == 1 ? &tmBlo : &tmBhi, &full[stage], kc * kKnnKC, n0);
              }
              __syncwarp();
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint64_t desc0 = umma_desc_kmajor(0, 128);
      const uint32_t idesc = umma_idesc_bf16(128, kKnnBN);
      const int iters = p.k_iters * p.passes;
      for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
        for (int ct = 0; ct < p.col_tiles; ++ct) {
          mbar_wait(&tempty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t da = desc0 + (((base + static_cast<uint32_t>(stage) * p.stage_bytes) & 0x3FFFFu) >> 4);
            const uint64_t db = da + (a_bytes >> 4);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < kKnnKC / 16; ++k) umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, (it | k) ? 1u : 0u);
              umma_commit(&empty[stage]);
              if (it + 1 == iters) umma_commit(&tfull[acc]);
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
    }
  } else {
    // ---- selection: one thread per query row keeps the C smallest (distance, index) pairs, sorted, in smem [pos][row]
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int C = p.C;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
      const int64_t qi = p.q_begin + static_cast<int64_t>(rt) * 128 + row;
      const bool qvalid = qi < p.q_end;
      const float nq = qvalid ? p.norms[qi] : 0.f;
      int cnt = 0;
      float tau = FLT_MAX;
      for (int ct = 0; ct < p.col_tiles; ++ct) {
        const int n0 = ct * kKnnBN;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * 256u;
        for (int c = 0; c < kKnnBN; c += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + static_cast<uint32_t>(c), r);
          tmem_ld_wait();
          if (qvalid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = n0 + c + j;
              if (col < p.N) {
                const float dist = nq + p.norms[col] - 2.f * __uint_as_float(r[j]);
                if (dist < tau || cnt < C) {
                  int pos = cnt < C ? cnt++ : C - 1;
                  while (pos > 0) {
                    const float pd = cd[(pos - 1) * 128 + row];
                    if (pd < dist || (pd == dist && ci[(pos - 1) * 128 + row] < col)) break;
                    cd[pos * 128 + row] = pd;
                    ci[pos * 128 + row] = ci[(pos - 1) * 128 + row];
                    --pos;
                  }
                  cd[pos * 128 + row] = dist;
                  ci[pos * 128 + row] = col;
                  if (cnt == C) tau = cd[(C - 1) * 128 + row];
                }
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
      if (qvalid) {
        const int64_t o = (qi - p.q_begin) * C;
        for (int j = 0; j < C; ++j) {
          p.cand_idx[o + j] = j < cnt ? ci[j * 128 + row] : -1;
          p.cand_d[o + j] = j < cnt ? cd[j * 128 + row] : FLT_MAX;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one block per query row: exact float64 distances to its candidates, (distance, index) sort, prune self, certify
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const float* __restrict__ X, int64_t N, int d, int64_t q_begin, int64_t nq, int C, int k,
                  const int* __restrict__ cand_idx, const float* __restrict__ cand_d, int64_t* __restrict__ nn_out,
                  double* __restrict__ radius_out, int* __restrict__ flags, unsigned int* __restrict__ max_err_bits,
                  float margin) {
  extern __shared__ float xq[];  // [d]
  __shared__ double ex[kKnnMaxC];
  __shared__ int idx[kKnnMaxC];
  __shared__ double sd[kKnnMaxC];
  __shared__ int si[kKnnMaxC];
  const int64_t r = blockIdx.x;
  if (r >= nq) return;
  const int64_t qi = q_begin + r;
  for (int t = threadIdx.x; t < d; t += blockDim.x) xq[t] = X[qi * d + t];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < C; c += (blockDim.x >> 5)) {
    const int j = cand_idx[r * C + c];
    double s = 0.0;
    if (j >= 0) {
      const float* y = X + static_cast<int64_t>(j) * d;
      for (int t = lane; t < d; t += 32) {
        const double df = static_cast<double>(xq[t]) - static_cast<double>(y[t]);
        s = fma(df, df, s);
      }
      s = warp_sum_d(s);
    }
    if (lane == 0) {
      ex[c] = j >= 0 ? s : DBL_MAX;
      idx[c] = j >= 0 ? j : 0x7fffffff;
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {  // rank sort by (distance, index)
    const double mv = ex[threadIdx.x];
    const int mi = idx[threadIdx.x];
    int rank = 0;
    for (int t = 0; t < C; ++t) rank += (ex[t] < mv || (ex[t] == mv && idx[t] < mi)) ? 1 : 0;
    sd[rank] = mv;
    si[rank] = mi;
    if (cand_idx[r * C + threadIdx.x] >= 0) {
      const float dev = fabsf(static_cast<float>(mv) - cand_d[r * C + threadIdx.x]);
      atomicMax(max_err_bits, __float_as_uint(dev));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int w = 0;
    for (int t = 0; t <= k && t < C && w < k; ++t)  // drop the query's own index by value, keep the first k
      if (si[t] != static_cast<int>(qi)) nn_out[r * k + w++] = si[t];
    for (; w < k; ++w) nn_out[r * k + w] = -1;
    const int kk = k < C ? k : C - 1;
    radius_out[r] = static_cast<double>(sqrtf(static_cast<float>(sd[kk])));
    // certified iff nothing outside the candidate list can beat the exact (k+1)-th distance
    const float kept_max = cand_d[r * C + C - 1];  // FLT_MAX when fewer than C database rows exist
    const bool ok = (N <= C) || (sd[kk] + static_cast<double>(margin) < static_cast<double>(kept_max));
    flags[r] = ok ? 0 : 1;
  }
}

// brute-force float64 distances from one row to every database row (uncertified rows only)
__global__ void knn_exact_dist_kernel(const float* __restrict__ X, int64_t N, int d, int64_t row,
                                      double* __restrict__ out) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= N) return;
  const int lane = threadIdx.x & 31;
  const float* a = X + row * d;
  const float* b = X + j * d;
  double s = 0.0;
  for (int t = lane; t < d; t += 32) {
    const double df = static_cast<double>(a[t]) - static_cast<double>(b[t]);
    s = fma(df, df, s);
  }
  s = warp_sum_d(s);
  if (lane == 0) out[j] = s;
}
// k+1 successive lexicographic minima of (distance, index) over N entries; single block
__global__ void __launch_bounds__(1024)
knn_exact_select_kernel(const double* __restrict__ dist, int64_t N, int64_t row, int k, int64_t* __restrict__ nn_out,
                        double* __restrict__ radius_out) {
  __shared__ double bd[32];
  __shared__ long long bi[32];
  __shared__ double last_d;
  __shared__ long long last_i;
  if (threadIdx.x == 0) { last_d = -1.0; last_i = -1; }
  __syncthreads();
  int w = 0;
  for (int it = 0; it <= k; ++it) {
    double best = DBL_MAX;
    long long besti = LLONG_MAX;
    const double ld = last_d;
    const long long li = last_i;
    for (int64_t j = threadIdx.x; j < N; j += blockDim.x) {
      const double v = dist[j];
      const bool after = v > ld || (v == ld && j > li);
      if (after && (v < best || (v == best && j < besti))) { best = v; besti = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bd[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 1; t < (blockDim.x >> 5); ++t)
        if (bd[t] < bd[0] || (bd[t] == bd[0] && bi[t] < bi[0])) { bd[0] = bd[t]; bi[0] = bi[t]; }
      last_d = bd[0];
      last_i = bi[0];
      if (bi[0] != row && w < k && bi[0] != LLONG_MAX) nn_out[w++] = bi[0];
      if (it == k) radius_out[0] = static_cast<double>(sqrtf(static_cast<float>(bd[0])));
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace icgan

using namespace icgan;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int icgan_knn_prepare(const float* X, void* Xhi, void* Xlo, float* norms, int64_t N, int d, void* stream) {
  ICGAN_REQUIRE(X && Xhi && Xlo && norms && N > 0 && d > 0, "icgan_knn_prepare: bad arguments");
  knn_prepare_kernel<<<static_cast<unsigned>((N + 7) / 8), 256, 0, STREAM>>>(
      X, static_cast<__nv_bfloat16*>(Xhi), static_cast<__nv_bfloat16*>(Xlo), norms, N, d);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_knn_coarse(const void* Xhi, const void* Xlo, const float* norms, int64_t N, int d, int64_t q_begin,
                                int64_t q_end, int C, int passes, int* cand_idx, float* cand_d, void* stream) {
  ICGAN_REQUIRE(Xhi && Xlo && norms && cand_idx && cand_d, "icgan_knn_coarse: null pointer");
  ICGAN_REQUIRE(d % 8 == 0 && C >= 1 && C <= kKnnMaxC && (passes == 1 || passes == 3), "icgan_knn_coarse: bad d/C/passes");
  ICGAN_REQUIRE(0 <= q_begin && q_begin < q_end && q_end <= N && N < (1ll << 31), "icgan_knn_coarse: bad row range");
  KnnParams p{};
  p.N = N; p.q_begin = q_begin; p.q_end = q_end; p.d = d; p.C = C; p.passes = passes;
  p.k_iters = ceil_div(d, kKnnKC);
  p.col_tiles = ceil_div(N, kKnnBN);
  p.row_tiles = ceil_div(q_end - q_begin, 128);
  p.stage_bytes = 128u * 128u + kKnnBN * 128u;
  const uint32_t tail = 2u * kKnnMaxC * 128u * 4u + 512u;
  p.stages = static_cast<int>((kKnnSmem - 1024u - tail) / p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  ICGAN_REQUIRE(p.stages >= 2, "icgan_knn_coarse: pipeline does not fit shared memory");
  p.norms = norms; p.cand_idx = cand_idx; p.cand_d = cand_d;
  CUtensorMap ahi, alo, bhi, blo;
  int rc = make_map2(&ahi, Xhi, d, N, 128);
  if (!rc) rc = make_map2(&alo, Xlo, d, N, 128);
  if (!rc) rc = make_map2(&bhi, Xhi, d, N, kKnnBN);
  if (!rc) rc = make_map2(&blo, Xlo, d, N, kKnnBN);
  if (rc) return rc;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(knn_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnSmem));
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail + 1024u;
  const int grid = p.row_tiles < num_sms() ? p.row_tiles : num_sms();
  knn_coarse_kernel<<<grid, kKnnThreads, smem_bytes, STREAM>>>(ahi, alo, bhi, blo, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_knn_rerank(const float* X, int64_t N, int d, int64_t q_begin, int64_t q_end, int C, int k,
                                const int* cand_idx, const float* cand_d, int64_t* nn_out, double* radius_out,
                                int* flags, float* max_err, float margin, void* stream) {
  ICGAN_REQUIRE(X && cand_idx && cand_d && nn_out && radius_out && flags && max_err, "icgan_knn_rerank: null pointer");
  ICGAN_REQUIRE(C >= 1 && C <= kKnnMaxC && k >= 1 && k < C, "icgan_knn_rerank: need 1 <= k < C <= %d", kKnnMaxC);
  ICGAN_REQUIRE(d * sizeof(float) <= 40 * 1024, "icgan_knn_rerank: feature dimension too large");
  const int64_t nq = q_end - q_begin;
  knn_rerank_kernel<<<static_cast<unsigned>(nq), 256, d * sizeof(float), STREAM>>>(
      X, N, d, q_begin, nq, C, k, cand_idx, cand_d, nn_out, radius_out, flags,
      reinterpret_cast<unsigned int*>(max_err), margin);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_knn_exact_row(const float* X, int64_t N, int d, int64_t row, int k, double* scratch,
                                   int64_t* nn_out_row, double* radius_out_row, void* stream) {
  ICGAN_REQUIRE(X && scratch && nn_out_row && radius_out_row && row >= 0 && row < N, "icgan_knn_exact_row: bad arguments");
  knn_exact_dist_kernel<<<static_cast<unsigned>((N + 7) / 8), 256, 0, STREAM>>>(X, N, d, row, scratch);
  knn_exact_select_kernel<<<1, 1024, 0, STREAM>>>(scratch, N, row, k, nn_out_row, radius_out_row);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
