// Tensor-core convolution kernels for sm_100a: TMA-staged NHWC tiles -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM).
// All kernels are persistent (one CTA per SM) and warp-specialised: one warp issues TMA, one issues the UMMAs and owns
// TMEM (both from warp-uniform code under elect.sync), the rest run the epilogue; two TMEM accumulator buffers let the
// epilogue of tile i overlap the main loop of tile i+1.
//
//   tc_conv_halo_kernel  : 3x3 / pad 1 forward and dgrad for H, W multiples of 16 (layers.SNConv2d.forward,
//                          BigGAN_PyTorch/layers.py:144-153; dgrad = same kernel on the flipped/transposed weight copy).
//                          16 x 16-pixel tiles, one 18-row TMA box per horizontal tap offset, the three vertical taps are
//                          row offsets into it (see the comment at the kernel).
//   tc_conv_kernel<4|8>  : implicit GEMM with one 128-pixel TMA box per filter tap: 1x1 convs, 3x3 at 4x4 / 8x8, plain
//                          GEMMs (H = 1); <4> carries the optional batch-norm statistics epilogue.
//   tc_conv_rgb_kernel   : first conv of D (3 -> ch): im2col built in shared memory by four builder warps.
//   tc_wgrad_kernel      : dWk[co, tap, ci] += sum_pixels dY[pixel, co] * X[pixel + tap, ci], MN-major UMMA operands read
//                          straight from the NHWC tensors, taps as column groups, split-K with float atomics.
//   tc_wgrad_halo_kernel : the same for layers with >= 192 channels: one dx per CTA, the three dy taps are one X box viewed
//                          at row offsets (N = 192 UMMAs), up to 256 output channels per CTA.
// Shared epilogue (epilogue_block): y = act(alpha * acc + bias [+ residual | gated by residual > 0]), all global loads
// of a 32-column block issued before the TMEM load is waited for, 32-byte stores.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace icgan {

static constexpr int kMaxStages = 12;
static constexpr int kThreads = 192;      // 2 role warps + 4 epilogue warps (wgrad, conv with fused BN statistics)
static constexpr int kConvThreads = 320;  // 2 role warps + 8 epilogue warps
static constexpr uint32_t kSmemBudget = 227u * 1024u;

// ------------------------------------------------------------------------------------------------ tensor maps
// bf16 tensor, dims[0] fastest. strides_bytes[i] is the stride of dims[i+1].
// elem_strides (optional): TMA traversal stride per dimension -- with stride s a box spanning box[i] tensor elements
// delivers ceil(box[i] / s) of them (every s-th), which is how the stride-2 convolutions read their input.
static int make_tmap_bf16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, uint32_t swizzle_bytes, const uint32_t* elem_strides = nullptr) {
  EncodeTiledFn fn = encode_fn();
  ICGAN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), gdim,
                  gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ICGAN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu)",
                static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0));
  return 0;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// ------------------------------------------------------------------------------------------------ epilogue helpers
__device__ __forceinline__ void load8(const void* base, int64_t elem_off, int is_bf16, float (&v)[8]) {
  if (is_bf16) {
    const uint4 raw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem_off);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
// 32-byte (one full L2 sector) store; `p` must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void store8(void* base, int64_t elem_off, int is_bf16, const float (&v)[8]) {
  if (is_bf16) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem_off) = raw;
  } else {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[i]);
    st_global_256(static_cast<float*>(base) + elem_off, w);  // 8 floats, 32-byte aligned (offsets are multiples of 8)
  }
}

// One 32-lane x NC-column accumulator block: y[pix, co0..co0+NC) = act(alpha * acc + bias + residual).
// Each epilogue warp has a scheduler to itself, so nothing hides a dependent load: every global load of the block
// (bias, residual) is issued before the TMEM load is waited for, and only then is anything consumed.
struct EpiArgs {
  void* y;
  const float* bias;
  const void* res;
  int Cout, out_bf16, res_bf16, act;
  float alpha;
  int res_mask;  // 1: `res` is not added but gates the result, y = res > 0 ? y : 0 (ReLU backward fused into dgrad)
};

template <int NC>
__device__ __forceinline__ void epilogue_block(const EpiArgs& e, uint32_t taddr, int co0, int64_t pix, int64_t rpix,
                                               bool valid) {
  constexpr int NG = NC / 8;
  uint32_t r[NC];
  if constexpr (NC == 32) tmem_ld32(taddr, r);
  else tmem_ld16(taddr, r);
  float4 bv[2 * NG];
  uint4 rv[2 * NG];
  bool on[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) on[g] = valid && (co0 + 8 * g < e.Cout);
  if (e.bias) {
#pragma unroll
    for (int g = 0; g < NG; ++g)
      if (on[g]) {
        bv[2 * g] = *reinterpret_cast<const float4*>(e.bias + co0 + 8 * g);
        bv[2 * g + 1] = *reinterpret_cast<const float4*>(e.bias + co0 + 8 * g + 4);
      }
  }
  if (e.res) {
    const int64_t ro = rpix * e.Cout + co0;
    if (e.res_bf16) {
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if (on[g]) rv[2 * g] = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(e.res) + ro + 8 * g);
    } else {
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if (on[g]) {
          rv[2 * g] = *reinterpret_cast<const uint4*>(static_cast<const float*>(e.res) + ro + 8 * g);
          rv[2 * g + 1] = *reinterpret_cast<const uint4*>(static_cast<const float*>(e.res) + ro + 8 * g + 4);
        }
    }
  }
  tmem_ld_wait();
  const bool wide = e.out_bf16 && (e.Cout % 16 == 0);  // 16 bf16 channels = one full 32-byte sector per store
#pragma unroll
  for (int gp = 0; gp < NG / 2; ++gp) {
    float v[2][8];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int g = 2 * gp + hh;
      if (!on[g]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[hh][j] = e.alpha * __uint_as_float(r[8 * g + j]);
      if (e.bias) {
        v[hh][0] += bv[2 * g].x; v[hh][1] += bv[2 * g].y; v[hh][2] += bv[2 * g].z; v[hh][3] += bv[2 * g].w;
        v[hh][4] += bv[2 * g + 1].x; v[hh][5] += bv[2 * g + 1].y; v[hh][6] += bv[2 * g + 1].z; v[hh][7] += bv[2 * g + 1].w;
      }
      if (e.res) {
        float rr[8];
        if (e.res_bf16) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv[2 * g]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            rr[2 * i] = f.x;
            rr[2 * i + 1] = f.y;
          }
        } else {
          const float* f0 = reinterpret_cast<const float*>(&rv[2 * g]);
          const float* f1 = reinterpret_cast<const float*>(&rv[2 * g + 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            rr[i] = f0[i];
            rr[4 + i] = f1[i];
          }
        }
        if (e.res_mask) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[hh][j] = rr[j] > 0.f ? v[hh][j] : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[hh][j] += rr[j];
        }
      }
      if (e.act == ICGAN_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[hh][j] = fmaxf(v[hh][j], 0.f);
      } else if (e.act == ICGAN_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[hh][j] = tanhf(v[hh][j]);
      }
    }
    const int64_t off = pix * e.Cout + co0 + 16 * gp;
    if (wide) {
      if (on[2 * gp]) {  // Cout % 16 == 0: both halves are in range together
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0][2 * i], v[0][2 * i + 1]);
          const __nv_bfloat162 hi = __floats2bfloat162_rn(v[1][2 * i], v[1][2 * i + 1]);
          w[i] = *reinterpret_cast<const uint32_t*>(&lo);
          w[4 + i] = *reinterpret_cast<const uint32_t*>(&hi);
        }
        st_global_256(static_cast<__nv_bfloat16*>(e.y) + off, w);
      }
    } else {
      if (on[2 * gp]) store8(e.y, off, e.out_bf16, v[0]);
      if (on[2 * gp + 1]) store8(e.y, off + 8, e.out_bf16, v[1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward / dgrad
struct TcConvParams {
  int B, H, W, Cout;
  int ksz, pad, taps;
  int TW, TH, TN;
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int BN, cw, chunks, k_iters, stages;
  int G, n_stage_iters;            // k-chunks per pipeline stage, stage iterations per tile
  uint32_t a_bytes, b_tx, b_chunk, stage_bytes, swz, idesc;
  int out_bf16, res_bf16, res_shift, act;
  void* y;
  const float* bias;
  const void* res;
  const float* alpha;  // device scalar multiplied into the accumulator (1/sigma of spectral norm), may be null
  float* stats;        // optional [2*Cout]: += sum_p (y - bias), += sum_p (y - bias)^2 of the (pre-rounding) outputs
  // Generalised tap list (icgan_conv2d_tc_ex): tile-domain pixel (h, w) reads input pixel (h*in_stride + tdh[t],
  // w*in_stride + tdw[t]) against weight slice twi[t], and is written to output pixel (h*osy + ooy, w*osx + oox) of an
  // [B, OH, OW, Cout] tensor.  Plain convolutions: tdh/tdw = -pad..pad, twi[t] = t, in_stride = os* = 1, oo* = 0.
  int8_t tdh[16], tdw[16], twi[16];
  int in_stride, OH, OW, osy, ooy, osx, oox;
};

struct TileCoord {
  int w0, h0, n0, co0;
};
__device__ __forceinline__ TileCoord decode_tile(const TcConvParams& p, int tile) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int tb = mt / p.tiles_h;
  t.w0 = tw * p.TW;
  t.h0 = th * p.TH;
  t.n0 = tb * p.TN;
  t.co0 = nt * p.BN;
  return t;
}

// kEpi = 4 (192 threads; the only variant with the BN-statistics epilogue) or 8 epilogue warps (320 threads: two
// warps per TMEM lane quadrant, alternating 32-column blocks).
template <int kEpi>
__global__ void __launch_bounds__(64 + 32 * kEpi, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzled tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* xpose = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][32][33] scratch for the BN-statistics epilogue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpi);  // one arrival per epilogue warp
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
    // ===================== TMA producer =====================
    // One pipeline stage carries up to G k-chunks (G A-boxes + G B-boxes, one barrier round trip), so that the MMA
    // thread always has >= ~4 UMMAs per wait. Tap/chunk counters advance incrementally: no integer division here.
    // The whole warp runs the loops (uniform control flow); one elected lane issues the TMA instructions.
    {
      if (lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t chunk_tx = p.a_bytes + p.b_tx;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        int tap = 0, cc = 0, left = p.k_iters;
        const int wbase = t.w0 * p.in_stride, hbase = t.h0 * p.in_stride;
        for (int it = 0; it < p.n_stage_iters; ++it) {
          const int n = left < p.G ? left : p.G;
          left -= n;
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* sa = smem + static_cast<size_t>(stage) * p.stage_bytes;
          uint8_t* sb = sa + static_cast<size_t>(p.G) * p.a_bytes;
          const bool leader = elect_one_sync();
          if (leader) mbar_expect_tx(&full[stage], static_cast<uint32_t>(n) * chunk_tx);
          for (int g = 0; g < n; ++g) {
            if (leader) {
              tma_load_4d(sa, &tmA, &full[stage], cc * p.cw, wbase + p.tdw[tap], hbase + p.tdh[tap], t.n0);
              tma_load_3d(sb, &tmB, &full[stage], cc * p.cw, p.twi[tap], t.co0);
            }
            sa += p.a_bytes;
            sb += p.b_chunk;
            if (++cc == p.chunks) {
              cc = 0;
              ++tap;
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
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int ksteps = p.cw / 16;
      const uint64_t desc0 = umma_desc_kmajor(0, p.swz);  // everything but the start address
      const uint32_t a_step = p.a_bytes >> 4, b_step = p.b_chunk >> 4;
      const uint32_t b_off = (static_cast<uint32_t>(p.G) * p.a_bytes) >> 4;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
        int left = p.k_iters;
        uint32_t first = 0;  // 0 for the very first UMMA of the tile (overwrite), 1 afterwards (accumulate)
        for (int it = 0; it < p.n_stage_iters; ++it) {
          const int n = left < p.G ? left : p.G;
          left -= n;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da0 = desc0 + (((base + static_cast<uint32_t>(stage) * p.stage_bytes) & 0x3FFFFu) >> 4);
          if (elect_one_sync()) {
            uint64_t da = da0, db = da0 + b_off;
            for (int g = 0; g < n; ++g) {
              for (int k = 0; k < ksteps; ++k) {  // +32 bytes (16 bf16) along K inside the swizzle atom
                umma_bf16(d_tmem, da + 2u * k, db + 2u * k, p.idesc, first);
                first = 1u;
              }
              da += a_step;
              db += b_step;
            }
            umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
            if (it + 1 == p.n_stage_iters) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
          }
          first = 1u;
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
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int grp = (warp - 2) >> 2;  // kEpi == 8: which of the two warps of this quadrant
    const int row = q * 32 + lane;
    const int wl = row % p.TW, hl = (row / p.TW) % p.TH, nl = row / (p.TW * p.TH);
    const float alpha = p.alpha ? *p.alpha : 1.f;
    const EpiArgs ea{p.y, p.bias, p.res, p.Cout, p.out_bf16, p.res_bf16, p.act, alpha, p.res_shift == 2};
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int n = t.n0 + nl, h = t.h0 + hl, w = t.w0 + wl;
      const int oy = h * p.osy + p.ooy, ox = w * p.osx + p.oox;
      const bool valid = (n < p.B) && (h < p.H) && (w < p.W) && (oy < p.OH) && (ox < p.OW);
      const int64_t pix = (static_cast<int64_t>(n) * p.OH + oy) * p.OW + ox;
      int64_t rpix = pix;
      if (p.res_shift == 1) rpix = (static_cast<int64_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * 256u;
      if (kEpi == 8 || !p.stats) {
        int c = 0, blk = 0;
        for (; c + 32 <= p.BN; c += 32, ++blk)
          if (kEpi == 4 || (blk & 1) == grp) epilogue_block<32>(ea, taddr + static_cast<uint32_t>(c), t.co0 + c, pix, rpix, valid);
        if (c < p.BN && (kEpi == 4 || (blk & 1) == grp))
          epilogue_block<16>(ea, taddr + static_cast<uint32_t>(c), t.co0 + c, pix, rpix, valid);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        continue;
      }
      auto emit8 = [&](const uint32_t* rr, int co) {
        if (valid && co < p.Cout) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = alpha * __uint_as_float(rr[j]);
          if (p.bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(p.bias + co);
            const float4 b1 = *reinterpret_cast<const float4*>(p.bias + co + 4);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (p.res) {
            float rv[8];
            load8(p.res, rpix * p.Cout + co, p.res_bf16, rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rv[j];
          }
          if (p.act == ICGAN_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (p.act == ICGAN_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j]);
          }
          store8(p.y, pix * p.Cout + co, p.out_bf16, v);
        }
      };
      int c = 0;
      for (; c + 32 <= p.BN; c += 32) {  // 32 accumulator columns per TMEM round trip
        uint32_t r[32];
        tmem_ld32(taddr + static_cast<uint32_t>(c), r);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) emit8(r + g8 * 8, t.co0 + c + g8 * 8);
        if (p.stats) {
          // batch-norm statistics of this 32-pixel x 32-channel block: transpose through smem, one channel per lane.
          // Values are centred on the bias (y - bias = alpha*acc [+ residual]) to keep E[x^2]-E[x]^2 well conditioned.
          float* xp = xpose + (warp - 2) * (32 * 33);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = 0.f;
            if (valid && t.co0 + c + j < p.Cout) {
              v = alpha * __uint_as_float(r[j]);
              if (p.res) v += p.res_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.res)[rpix * p.Cout + t.co0 + c + j])
                                         : static_cast<const float*>(p.res)[rpix * p.Cout + t.co0 + c + j];
            }
            xp[lane * 33 + j] = v;
          }
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const float v = xp[rr * 33 + lane];
            s1 += v;
            s2 = fmaf(v, v, s2);
          }
          const int co = t.co0 + c + lane;
          if (co < p.Cout) {
            atomicAdd(p.stats + co, s1);
            atomicAdd(p.stats + p.Cout + co, s2);
          }
          __syncwarp();
        }
      }
      for (; c < p.BN; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c), r);
        tmem_ld_wait();
        emit8(r, t.co0 + c);
        emit8(r + 8, t.co0 + c + 8);
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

// ------------------------------------------------------------------------------------------------ 3x3, halo reuse
// tc_conv_halo_kernel: 3x3 / pad 1 / stride 1 convolution for H, W multiples of 16.  tc_conv_kernel above stages one
// 128-pixel box per filter tap, i.e. every input pixel crosses L2 -> shared memory 9 times, and with <= 256 output
// channels per tile that traffic (not the tensor pipe) bounds the kernel: the whole chip moves ~6300 B/clk out of L2
// (~43 B/clk/SM) against 8192 flop/clk/SM, so a tile needs ~190 flop per staged byte; 128 x N tiles give 55 (N=96)
// to 85 (N=256).  Here a tile is a 16 x 16 pixel square (M = 2 x 128) and one TMA box carries the 18 x 16 pixel
// block (rows h0-1 .. h0+16, columns shifted by dx-1, 64/32/16 channels) for ONE horizontal tap offset dx.  The
// three vertical taps of that dx read the same block at smem row offsets dy*16 pixels -- whole multiples of the
// swizzle atom, so only the descriptor start address moves -- and both 128-pixel halves of the tile share the
// weight boxes.  Staged bytes per output pixel drop 2.3x (A: 9 x 128 -> 3 x 144 pixel rows per 128 outputs, B halved).
struct TcHaloParams {
  int B, H, W, Cout;
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int BN, cw, chunks, G, n_units, n_stage_iters, stages;
  uint32_t a_bytes, b_tx, b_slot, unit_bytes, stage_bytes, swz, idesc;
  int out_bf16, res_bf16, res_shift, act;
  void* y;
  const float* bias;
  const void* res;
  const float* alpha;
};

static constexpr int kHaloT = 16;  // tile edge in pixels

struct HaloTile {
  int w0, h0, n, co0;
};
__device__ __forceinline__ HaloTile decode_halo_tile(const TcHaloParams& p, int tile) {
  HaloTile t;
  const int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  t.w0 = tw * kHaloT;
  t.h0 = (mt % p.tiles_h) * kHaloT;
  t.n = mt / p.tiles_h;
  t.co0 = nt * p.BN;
  return t;
}

__global__ void __launch_bounds__(kConvThreads, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const TcHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
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
    // ===================== TMA producer: unit = (channel chunk, dx): one 18x16-pixel A box + three weight boxes
    {
      if (lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t unit_tx = p.a_bytes + 3u * p.b_tx;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const HaloTile t = decode_halo_tile(p, tile);
        int cc = 0, dx = 0, left = p.n_units;
        for (int it = 0; it < p.n_stage_iters; ++it) {
          const int n = left < p.G ? left : p.G;
          left -= n;
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* su = smem + static_cast<size_t>(stage) * p.stage_bytes;
          const bool leader = elect_one_sync();
          if (leader) mbar_expect_tx(&full[stage], static_cast<uint32_t>(n) * unit_tx);
          for (int g = 0; g < n; ++g) {
            if (leader) {
              tma_load_4d(su, &tmA, &full[stage], cc * p.cw, t.w0 + dx - 1, t.h0 - 1, t.n);
              uint8_t* sb = su + p.a_bytes;
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
                tma_load_3d(sb + dy * p.b_slot, &tmB, &full[stage], cc * p.cw, dy * 3 + dx, t.co0);
            }
            su += p.unit_bytes;
            if (++dx == 3) {
              dx = 0;
              ++cc;
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
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int ksteps = p.cw / 16;
      const uint64_t desc0 = umma_desc_kmajor(0, p.swz);
      const uint32_t row16 = (static_cast<uint32_t>(kHaloT) * p.swz) >> 4;  // one image row of the block, in 16-byte units
      const uint32_t b_off = p.a_bytes >> 4, b_step = p.b_slot >> 4, u_step = p.unit_bytes >> 4;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + static_cast<uint32_t>(acc) * 256u;
        const uint32_t d1 = d0 + static_cast<uint32_t>(p.BN);
        int left = p.n_units;
        uint32_t first = 0;
        for (int it = 0; it < p.n_stage_iters; ++it) {
          const int n = left < p.G ? left : p.G;
          left -= n;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t du0 = desc0 + (((base + static_cast<uint32_t>(stage) * p.stage_bytes) & 0x3FFFFu) >> 4);
          if (elect_one_sync()) {
            uint64_t du = du0;
            for (int g = 0; g < n; ++g) {
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                const uint64_t da0 = du + static_cast<uint32_t>(dy) * row16;  // output rows 0..7  of the tile
                const uint64_t da1 = da0 + 8u * row16;                        // output rows 8..15
                const uint64_t db = du + b_off + static_cast<uint32_t>(dy) * b_step;
                for (int k = 0; k < ksteps; ++k) {
                  umma_bf16(d0, da0 + 2u * k, db + 2u * k, p.idesc, first);
                  umma_bf16(d1, da1 + 2u * k, db + 2u * k, p.idesc, first);
                  first = 1u;
                }
              }
              du += u_step;
            }
            umma_commit(&empty[stage]);
            if (it + 1 == p.n_stage_iters) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
          }
          first = 1u;
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
    // ===================== epilogue: 8 warps; warp (q, half) owns TMEM lanes [32q, 32q+32) of one 128-pixel half
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = half * 128 + q * 32 + lane;
    const float alpha = p.alpha ? *p.alpha : 1.f;
    const EpiArgs ea{p.y, p.bias, p.res, p.Cout, p.out_bf16, p.res_bf16, p.act, alpha, p.res_shift == 2};
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const HaloTile t = decode_halo_tile(p, tile);
      const int h = t.h0 + (m >> 4), w = t.w0 + (m & 15);
      const int64_t pix = (static_cast<int64_t>(t.n) * p.H + h) * p.W + w;
      int64_t rpix = pix;
      if (p.res_shift == 1) rpix = (static_cast<int64_t>(t.n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * 256u +
                             static_cast<uint32_t>(half * p.BN);
      int c = 0;
      for (; c + 32 <= p.BN; c += 32) epilogue_block<32>(ea, taddr + static_cast<uint32_t>(c), t.co0 + c, pix, rpix, true);
      if (c < p.BN) epilogue_block<16>(ea, taddr + static_cast<uint32_t>(c), t.co0 + c, pix, rpix, true);
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

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Returns kHaloIneligible when the shape does not fit (caller falls back to tc_conv_kernel), else 0 / an error code.
static constexpr int kHaloIneligible = -2;
static int launch_conv_halo(const void* x, const void* wk, const float* alpha_dev, const float* bias,
                            const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int out_dtype,
                            int res_dtype, int res_shift, int act, cudaStream_t stream) {
  static const int enabled = env_int("ICGAN_TC_HALO", 1);
  if (!enabled || (H % kHaloT) || (W % kHaloT)) return kHaloIneligible;
  TcHaloParams p{};
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.tiles_w = W / kHaloT;
  p.tiles_h = H / kHaloT;
  static const int cw_override = env_int("ICGAN_TC_HALO_CW", 0);
  p.cw = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  if (cw_override && cw_override <= p.cw && Cin % cw_override == 0) p.cw = cw_override;
  // Cin = 96, 160, ...: 64-channel chunks with the last one part out of bounds (TMA zero-fills both operands).
  // Measured twice, a loss both times: Cout = 96: 932 -> 866 TFLOP/s (the zero K columns cost MMA time); the 96 -> 3
  // image conv (Cout padded to 8, pure A traffic): 16.6 -> 22.4 ms per step -- boxes that are half out of bounds move
  // slower than the narrower 64-byte-row boxes they replace.  Kept as an experiment switch only.
  static const int pad64 = env_int("ICGAN_TC_HALO_PAD64", 0);
  if (pad64 && p.cw < 64 && Cin > 64) p.cw = 64;
  p.swz = static_cast<uint32_t>(p.cw * 2);
  p.chunks = ceil_div(Cin, p.cw);
  p.n_units = 3 * p.chunks;
  if (Cout <= 128) p.BN = (Cout + 15) / 16 * 16;
  else if (Cout % 128 == 0) p.BN = 128;
  else if (Cout % 96 == 0) p.BN = 96;
  else if (Cout % 112 == 0) p.BN = 112;
  else if (Cout % 80 == 0) p.BN = 80;
  else if (Cout % 64 == 0) p.BN = 64;
  else p.BN = 128;
  p.n_tiles = ceil_div(Cout, p.BN);
  p.total_tiles = p.tiles_w * p.tiles_h * B * p.n_tiles;
  p.a_bytes = static_cast<uint32_t>((kHaloT + 2) * kHaloT) * p.swz;
  p.b_tx = static_cast<uint32_t>(p.BN) * p.swz;
  p.b_slot = (p.b_tx + 1023u) & ~1023u;
  p.unit_bytes = p.a_bytes + 3u * p.b_slot;
  static const int stage_target = env_int("ICGAN_TC_HALO_STAGE_KB", 48) * 1024;
  int G = stage_target / static_cast<int>(p.unit_bytes);
  if (G < 1) G = 1;
  if (G > p.n_units) G = p.n_units;
  p.G = G;
  p.n_stage_iters = ceil_div(p.n_units, G);
  p.stage_bytes = static_cast<uint32_t>(G) * p.unit_bytes;
  const uint32_t tail = 1024u + 512u;
  int stages = static_cast<int>((kSmemBudget - tail) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return kHaloIneligible;
  p.stages = stages;
  p.idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.BN));
  p.out_bf16 = out_dtype == ICGAN_BF16;
  p.res_bf16 = res_dtype == ICGAN_BF16;
  p.res_shift = res_shift;
  p.act = act;
  p.y = y; p.bias = bias; p.res = residual; p.alpha = alpha_dev;

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {(uint32_t)p.cw, (uint32_t)kHaloT, (uint32_t)(kHaloT + 2), 1u};
    int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, p.swz);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, 9ull, (uint64_t)Cout};
    const uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)9 * Cin * 2};
    const uint32_t box[3] = {(uint32_t)p.cw, 1u, (uint32_t)p.BN};
    int rc = make_tmap_bf16(&tmB, wk, 3, dims, str, box, p.swz);
    if (rc) return rc;
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  tc_conv_halo_kernel<<<grid, kConvThreads, smem_bytes, stream>>>(tmA, tmB, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

// Split-K factor of the weight-gradient kernels: `items` output tiles x `splits` pixel ranges = CTAs (one per SM at a
// time, `rounds` waves of them), each walking ceil(k_chunks / splits) stages plus a fixed prologue/epilogue worth about
// 16 stages.  The old rule (2 * SMs / items) often produced e.g. 2.2 waves, i.e. a third wave at 20 % occupancy.
static int choose_splits(int items, int k_chunks) {
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 160 && s <= k_chunks; ++s) {
    const int rounds = ceil_div(items * s, num_sms());
    const double cost = static_cast<double>(rounds) * (ceil_div(k_chunks, s) + 16.0);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

// ------------------------------------------------------------------------------------------------ RGB-side input
// tc_conv_rgb_kernel: k x k convolution of a <= 3-channel image (k*k*Cs <= 32, the first layer of D,
// BigGAN.py:491 with DBlock conv1).  The reduction is only 27 long, so the layer is bound by writing its output; an
// explicit im2col buffer (32 bf16 per pixel) cost more traffic than the result.  Here four builder warps gather each
// pixel's 27 inputs and write the 128 x 32 bf16 A tile straight into shared memory in the UMMA K-major/64-byte-swizzle
// layout (16-byte chunk c of row r lands at chunk c ^ ((r >> 1) & 3)), the weights sit in shared memory for the whole
// kernel, one elected lane issues the two K=16 UMMAs of a tile, and eight epilogue warps stream the result out.
struct TcRgbParams {
  int B, H, W, Cs, ksz, Cout, BN;
  int64_t P;
  int total_tiles;
  uint32_t idesc;
  const __nv_bfloat16* x;
  const __nv_bfloat16* wcol;  // [Cout][32], column j = tap * Cs + ci
  int out_bf16, act;
  void* y;
  const float* bias;
  const float* alpha;
};
static constexpr int kRgbThreads = 13 * 32;  // 4 builder warps, 1 MMA warp, 8 epilogue warps
static constexpr int kRgbBufs = 4;

// KSZ / CS > 0: compile-time filter size and channel count (the gather unrolls into registers); 0 = run-time values.
template <int KSZ, int CS>
__global__ void __launch_bounds__(kRgbThreads, 1)
tc_conv_rgb_kernel(const TcRgbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* sA = smem;                                  // kRgbBufs x 128 rows x 64 bytes
  uint8_t* sB = smem + kRgbBufs * 8192;                // 256 rows x 64 bytes
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + 256 * 64);
  uint64_t* a_empty = a_full + kRgbBufs;
  uint64_t* tfull = a_empty + kRgbBufs;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRgbBufs; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  // weights -> shared memory (swizzled K-major rows of 64 bytes; rows >= Cout are zero)
  for (int idx = threadIdx.x; idx < p.BN * 4; idx += blockDim.x) {
    const int r = idx >> 2, c = idx & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < p.Cout) v = *reinterpret_cast<const uint4*>(p.wcol + static_cast<int64_t>(r) * 32 + c * 8);
    *reinterpret_cast<uint4*>(sB + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;
  }
  fence_proxy_async();
  if (warp == 4) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== builders: thread t owns row t of the A tile =====================
    // The gathers of the next two tiles are in flight while a tile is packed and published (nothing else would hide
    // their L2/HBM latency: a builder warp has its scheduler to itself).
    const int r = threadIdx.x;
    const int ksz = KSZ > 0 ? KSZ : p.ksz, cs = CS > 0 ? CS : p.Cs;
    const int pad = ksz >> 1;
    constexpr int NV = (KSZ > 0 && CS > 0) ? KSZ * KSZ * CS : 32;
    auto gather = [&](int tile, unsigned short (&rv)[NV]) {
#pragma unroll
      for (int j = 0; j < NV; ++j) rv[j] = 0;
      const int64_t pix = static_cast<int64_t>(tile) * 128 + r;
      if (tile >= p.total_tiles || pix >= p.P) return;
      const int w = static_cast<int>(pix % p.W);
      const int h = static_cast<int>((pix / p.W) % p.H);
      const int64_t n = pix / (static_cast<int64_t>(p.H) * p.W);
#pragma unroll
      for (int kh = 0; kh < (KSZ > 0 ? KSZ : 3); ++kh) {
        if (kh >= ksz) break;
        const int ih = h + kh - pad;
#pragma unroll
        for (int kw = 0; kw < (KSZ > 0 ? KSZ : 3); ++kw) {
          if (kw >= ksz) break;
          const int iw = w + kw - pad;
          const bool in = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
          const unsigned short* src = reinterpret_cast<const unsigned short*>(p.x) + ((n * p.H + ih) * p.W + iw) * cs;
#pragma unroll
          for (int ci = 0; ci < (CS > 0 ? CS : 3); ++ci) {
            if (ci >= cs) break;
            if (in) rv[(kh * ksz + kw) * cs + ci] = __ldg(src + ci);
          }
        }
      }
    };
    int buf = 0;
    uint32_t phase = 0;
    auto publish = [&](const unsigned short (&rv)[NV]) {
      uint32_t wv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t lo = 2 * i < NV ? rv[2 * i] : 0u, hi = 2 * i + 1 < NV ? rv[2 * i + 1] : 0u;
        wv[i] = lo | (hi << 16);
      }
      mbar_wait(&a_empty[buf], phase ^ 1u);
      uint8_t* row = sA + buf * 8192 + r * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(row + ((c ^ ((r >> 1) & 3)) << 4)) =
            make_uint4(wv[4 * c], wv[4 * c + 1], wv[4 * c + 2], wv[4 * c + 3]);
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[buf]);
      if (++buf == kRgbBufs) {
        buf = 0;
        phase ^= 1u;
      }
    };
    unsigned short ra[NV], rb[NV];
    const int step = static_cast<int>(gridDim.x);
    int tile = blockIdx.x;
    gather(tile, ra);
    gather(tile + step, rb);
    while (tile < p.total_tiles) {
      publish(ra);
      gather(tile + 2 * step, ra);
      if (tile + step >= p.total_tiles) break;
      publish(rb);
      gather(tile + 3 * step, rb);
      tile += 2 * step;
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    int buf = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint64_t desc0 = umma_desc_kmajor(0, 64);
    const uint64_t db = desc0 + (((base + kRgbBufs * 8192u) & 0x3FFFFu) >> 4);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      mbar_wait(&a_full[buf], phase);
      tc_fence_after();
      const uint64_t da = desc0 + (((base + static_cast<uint32_t>(buf) * 8192u) & 0x3FFFFu) >> 4);
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
      if (elect_one_sync()) {
        umma_bf16(d_tmem, da, db, p.idesc, 0u);
        umma_bf16(d_tmem, da + 2u, db + 2u, p.idesc, 1u);
        umma_commit(&a_empty[buf]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++buf == kRgbBufs) {
        buf = 0;
        phase ^= 1u;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue: 8 warps, two per TMEM lane quadrant =====================
    const int q = warp & 3;
    const int grp = (warp - 5) >> 2;
    const float alpha = p.alpha ? *p.alpha : 1.f;
    const EpiArgs ea{p.y, p.bias, nullptr, p.Cout, p.out_bf16, 0, p.act, alpha, 0};
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int64_t pix = static_cast<int64_t>(tile) * 128 + q * 32 + lane;
      const bool valid = pix < p.P;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * 256u;
      int c = 0, blk = 0;
      for (; c + 32 <= p.BN; c += 32, ++blk)
        if ((blk & 1) == grp) epilogue_block<32>(ea, taddr + static_cast<uint32_t>(c), c, pix, pix, valid);
      if (c < p.BN && (blk & 1) == grp) epilogue_block<16>(ea, taddr + static_cast<uint32_t>(c), c, pix, pix, valid);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ wgrad
// dWk[co, tap, ci] += sum_pixels dY[pixel, co] * X[pixel + tap, ci], operands read straight from the NHWC tensors:
// the reduction index (pixels) is the ROW of each TMA box and channels are contiguous, i.e. both UMMA operands are
// "MN-major".  A stage holds KP=64 pixels: A = two 64-channel boxes of dY (M = 128 output channels); B = one
// 64-channel box of X per filter tap at the tap-shifted pixel coordinates (halo zero-filled by TMA).  The tap boxes sit
// back to back in shared memory, so they are simply consecutive 64-wide COLUMN GROUPS of one MN-major B operand
// (leading-dimension byte offset = box size): a single UMMA with N=256 multiplies dY against 4 taps at once.
// A CTA owns a group of <= 5 taps (5*64 = 320 of the 512 TMEM columns); 3x3 filters use two groups (5 + 4 taps).
// For 1x1 filters the column groups are further input channels instead of taps.
struct TcWgradParams {
  int Cin, Cout, ksz, pad, taps;
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_b, k_chunks;
  int co_tiles, ci_tiles, groups, splits;
  int ci_per_tile, stages;
  uint32_t a_bytes, box_bytes, stage_bytes;
  float* dwk;
  // generalised taps (icgan_conv2d_wgrad_tc_ex): tap t pairs dY pixel (h, w) with X pixel (h*in_stride + tdh[t], ...)
  int8_t tdh[16], tdw[16];
  int in_stride;
};

static constexpr int kWgradKP = 64;      // pixels (reduction rows) per pipeline stage
static constexpr int kWgradMaxBoxes = 5;  // 64-channel column groups per CTA

// MN-major SW128 operand: 64-channel column groups `lbo_bytes` apart, 8-pixel row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                const TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
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

  // work item: (co tile, tap/channel group, ci tile, pixel-chunk range)
  int wi = blockIdx.x;
  const int split = wi % p.splits;
  wi /= p.splits;
  const int ci_t = wi % p.ci_tiles;
  wi /= p.ci_tiles;
  const int grp = wi % p.groups;
  const int co_t = wi / p.groups;
  const int per = (p.k_chunks + p.splits - 1) / p.splits;
  const int kc_begin = split * per;
  const int kc_end = min(p.k_chunks, kc_begin + per);
  const int n_chunks = max(0, kc_end - kc_begin);
  const int co0 = co_t * 128, ci0 = ci_t * p.ci_per_tile;
  // column groups (boxes) of this CTA: taps [tap0, tap0+nb) for 3x3, further 64-channel slices for 1x1
  int tap0 = 0, nb;
  if (p.taps > 1) {
    tap0 = grp * kWgradMaxBoxes;
    nb = min(kWgradMaxBoxes, p.taps - tap0);
  } else {
    nb = min(kWgradMaxBoxes, (min(p.ci_per_tile, p.Cin - ci0) + 63) / 64);
  }
  const uint32_t n1 = static_cast<uint32_t>(min(nb, 4) * 64), n2 = static_cast<uint32_t>((nb - min(nb, 4)) * 64);
  const uint32_t idesc1 = umma_idesc_bf16(128, n1) | (1u << 15) | (1u << 16);  // A and B MN-major
  const uint32_t idesc2 = umma_idesc_bf16(128, n2 ? n2 : 64) | (1u << 15) | (1u << 16);

  if (warp == 0) {
    // whole warp runs the loop (uniform control flow), one elected lane issues the TMA instructions
    if (n_chunks > 0) {
      if (lane == 0) {
        tma_prefetch_desc(&tmDy);
        tma_prefetch_desc(&tmX);
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = p.a_bytes + static_cast<uint32_t>(nb) * p.box_bytes;
      int tw = kc_begin % p.tiles_w, th = (kc_begin / p.tiles_w) % p.tiles_h, tb = kc_begin / (p.tiles_w * p.tiles_h);
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tb * p.TN;
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            ++tb;
          }
        }
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* sa = smem + static_cast<size_t>(stage) * p.stage_bytes;
        if (elect_one_sync()) {
          mbar_expect_tx(&full[stage], tx);
          tma_load_4d(sa, &tmDy, &full[stage], co0, w0, h0, n0);
          tma_load_4d(sa + p.box_bytes, &tmDy, &full[stage], co0 + 64, w0, h0, n0);
          uint8_t* sb = sa + p.a_bytes;
          if (p.taps > 1) {
            for (int j = 0; j < nb; ++j) {
              tma_load_4d(sb, &tmX, &full[stage], ci0, w0 * p.in_stride + p.tdw[tap0 + j],
                          h0 * p.in_stride + p.tdh[tap0 + j], n0);
              sb += p.box_bytes;
            }
          } else {
            for (int j = 0; j < nb; ++j) {
              tma_load_4d(sb, &tmX, &full[stage], ci0 + 64 * j, w0 * p.in_stride + p.tdw[0],
                          h0 * p.in_stride + p.tdh[0], n0);
              sb += p.box_bytes;
            }
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (n_chunks > 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t desc0 = umma_desc_mnmajor(0, p.box_bytes);
      const uint32_t b_off = p.a_bytes >> 4, b2_off = (p.a_bytes + 4u * p.box_bytes) >> 4;
      for (int it = 0; it < n_chunks; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t da = desc0 + (((base + static_cast<uint32_t>(stage) * p.stage_bytes) & 0x3FFFFu) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kWgradKP / 16; ++k) {  // 16 pixel rows = 2 swizzle atoms = 2048 bytes per UMMA
            const uint32_t accf = (it | k) != 0 ? 1u : 0u;
            umma_bf16(tmem_base, da + 128u * k, da + b_off + 128u * k, idesc1, accf);
            if (n2) umma_bf16(tmem_base + 256u, da + 128u * k, da + b2_off + 128u * k, idesc2, accf);
          }
          umma_commit(&empty[stage]);
          if (it + 1 == n_chunks) umma_commit(&tfull[0]);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (n_chunks > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int j = 0; j < nb; ++j) {
      const int tap = p.taps > 1 ? tap0 + j : 0;
      const int cib = p.taps > 1 ? ci0 : ci0 + 64 * j;
      for (int c = 0; c < 64; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + static_cast<uint32_t>(j * 64 + c), r);
        tmem_ld_wait();
        if (co < p.Cout) {
          float* dst = p.dwk + (static_cast<int64_t>(co) * p.taps + tap) * p.Cin + cib + c;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cib + c + i < p.Cin) atomicAdd(dst + i, __uint_as_float(r[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ wgrad, halo reuse
// tc_wgrad_halo_kernel: 3x3 weight gradient for W % 16 == 0, H % 4 == 0.  Same MN-major formulation as tc_wgrad_kernel,
// but a CTA owns ONE horizontal tap offset dx and up to 256 output channels:
//   * a stage holds 64 pixels (16 wide x 4 rows): up to four 64-channel dY boxes (A, M = 2 x 128) and one X box of
//     (4+2) x 16 pixels loaded at column offset dx-1 (B);
//   * the three vertical taps are the same X box viewed at row offsets dy*16 pixels (2048 bytes), i.e. three
//     "column groups" of one MN-major operand whose leading-dimension stride is 2048 bytes (the groups overlap in
//     shared memory, which only matters to the address generator): ONE UMMA with N = 192 covers them.
// Staged rows per flop drop 1.5x against tc_wgrad_kernel (which reloads X for every tap and dY for both tap groups).
struct TcWgradHaloParams {
  int Cin, Cout;
  int tiles_w, tiles_h, k_chunks;
  int co_tiles, ci_tiles, splits, stages;
  uint32_t stage_bytes;
  float* dwk;
};
static constexpr uint32_t kWhBox = 64u * 128u;       // one dY box: 64 pixels x 64 channels
static constexpr uint32_t kWhXBox = 96u * 128u;      // X box: (4+2) x 16 pixels x 64 channels
static constexpr uint32_t kWhRow = 16u * 128u;       // one image row of a box (16 pixels)

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                     const TcWgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
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

  // work item: (co tile, ci tile, dx, pixel-chunk range)
  int wi = blockIdx.x;
  const int split = wi % p.splits;
  wi /= p.splits;
  const int dx = wi % 3;
  wi /= 3;
  const int ci_t = wi % p.ci_tiles;
  const int co_t = wi / p.ci_tiles;
  const int per = (p.k_chunks + p.splits - 1) / p.splits;
  const int kc_begin = split * per;
  const int kc_end = min(p.k_chunks, kc_begin + per);
  const int n_chunks = max(0, kc_end - kc_begin);
  const int co0 = co_t * 256, ci0 = ci_t * 64;
  const int mb = min(4, (p.Cout - co0 + 63) / 64);  // 64-channel dY boxes of this tile
  const int halves = mb > 2 ? 2 : 1;
  const uint32_t idesc = umma_idesc_bf16(128, 192) | (1u << 15) | (1u << 16);  // A and B MN-major

  if (warp == 0) {
    if (n_chunks > 0) {
      if (lane == 0) {
        tma_prefetch_desc(&tmDy);
        tma_prefetch_desc(&tmX);
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = static_cast<uint32_t>(mb) * kWhBox + kWhXBox;
      const int per_img = p.tiles_w * p.tiles_h;
      int tw = kc_begin % p.tiles_w, th = (kc_begin / p.tiles_w) % p.tiles_h, tb = kc_begin / per_img;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        const int w0 = tw * 16, h0 = th * 4, n0 = tb;
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            ++tb;
          }
        }
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* sa = smem + static_cast<size_t>(stage) * p.stage_bytes;
        if (elect_one_sync()) {
          mbar_expect_tx(&full[stage], tx);
          for (int j = 0; j < mb; ++j) tma_load_4d(sa + j * kWhBox, &tmDy, &full[stage], co0 + 64 * j, w0, h0, n0);
          tma_load_4d(sa + 4u * kWhBox, &tmX, &full[stage], ci0, w0 + dx - 1, h0 - 1, n0);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (n_chunks > 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t desc_a0 = umma_desc_mnmajor(0, kWhBox);  // two 64-channel dY boxes = M 128
      const uint64_t desc_b0 = umma_desc_mnmajor(0, kWhRow);  // three dy views of the X box, 16 pixel rows apart = N 192
      for (int it = 0; it < n_chunks; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sbase = ((base + static_cast<uint32_t>(stage) * p.stage_bytes) & 0x3FFFFu) >> 4;
        const uint64_t da = desc_a0 + sbase;
        const uint64_t db = desc_b0 + sbase + ((4u * kWhBox) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 pixels (one image row of the tile) per UMMA
            const uint32_t accf = (it | k) != 0 ? 1u : 0u;
            umma_bf16(tmem_base, da + 128u * k, db + 128u * k, idesc, accf);
            if (halves == 2) umma_bf16(tmem_base + 256u, da + ((2u * kWhBox) >> 4) + 128u * k, db + 128u * k, idesc, accf);
          }
          umma_commit(&empty[stage]);
          if (it + 1 == n_chunks) umma_commit(&tfull[0]);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (n_chunks > 0) {
    const int q = warp & 3;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    for (int hf = 0; hf < halves; ++hf) {
      const int co = co0 + hf * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(hf) * 256u;
      for (int dy = 0; dy < 3; ++dy) {
        for (int c = 0; c < 64; c += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + static_cast<uint32_t>(dy * 64 + c), r);
          tmem_ld_wait();
          if (co < p.Cout) {
            float* dst = p.dwk + (static_cast<int64_t>(co) * 9 + dy * 3 + dx) * p.Cin + ci0 + c;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ci0 + c + i < p.Cin) atomicAdd(dst + i, __uint_as_float(r[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static constexpr int kWgradHaloIneligible = -2;
static int launch_wgrad_halo(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin, int Cout,
                             cudaStream_t stream) {
  // Measured (r01, bench step): 1.12-1.29x over tc_wgrad_kernel for Cin, Cout >= 192, 0.6-0.9x for 96-channel layers
  // (one M half, so only N = 192 per staged dY box instead of 320) -> mode 1 (default) takes the wide layers only.
  static const int mode = env_int("ICGAN_TC_WGRAD_HALO", 1);  // 0 off, 1 wide layers, 2 every eligible shape
  if (!mode || (W % 16) || (H % 4)) return kWgradHaloIneligible;
  if (mode == 1 && (Cin < 192 || Cout < 192)) return kWgradHaloIneligible;
  TcWgradHaloParams p{};
  p.Cin = Cin; p.Cout = Cout;
  p.tiles_w = W / 16;
  p.tiles_h = H / 4;
  p.k_chunks = p.tiles_w * p.tiles_h * B;
  p.co_tiles = ceil_div(Cout, 256);
  p.ci_tiles = ceil_div(Cin, 64);
  p.stage_bytes = 4u * kWhBox + kWhXBox;  // 44 KB (the dY slots of a narrow tile stay unused)
  const uint32_t tail = 1024u + 512u;
  int stages = static_cast<int>((kSmemBudget - tail) / p.stage_bytes);
  if (stages > 8) stages = 8;
  p.stages = stages;
  const int out_tiles = p.co_tiles * p.ci_tiles * 3;
  p.splits = choose_splits(out_tiles, p.k_chunks);
  p.dwk = dwk;

  CUtensorMap tmDy, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    const uint32_t box[4] = {64u, 16u, 4u, 1u};
    int rc = make_tmap_bf16(&tmDy, dy, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64u, 16u, 6u, 1u};
    int rc = make_tmap_bf16(&tmX, x, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const int grid = out_tiles * p.splits;
  tc_wgrad_halo_kernel<<<grid, kThreads, smem_bytes, stream>>>(tmDy, tmX, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

// NHWC [P][C] (f32 or bf16) -> [C][P] bf16 through a padded shared-memory tile.
template <typename T>
__global__ void nhwc_to_cnhw_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ xT, int64_t P, int C) {
  __shared__ float tile[32][33];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t pp = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (pp < P && c < C) ? ld_as_float(x, pp * C + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t pp = p0 + threadIdx.x;
    if (c < C && pp < P) xT[static_cast<int64_t>(c) * P + pp] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

}  // namespace icgan

using namespace icgan;

extern "C" int icgan_conv2d_tc(const void* x, const void* wk, const float* alpha_dev, const float* bias,
                               const void* residual, void* y, float* bn_stats, int B, int H, int W, int Cin, int Cout,
                               int ksize, int out_dtype, int res_dtype, int res_shift, int act, void* stream) {
  ICGAN_REQUIRE(x && wk && y, "icgan_conv2d_tc: null pointer");
  ICGAN_REQUIRE(ksize == 1 || ksize == 3, "icgan_conv2d_tc: ksize must be 1 or 3 (got %d)", ksize);
  ICGAN_REQUIRE(B > 0 && H > 0 && W > 0, "icgan_conv2d_tc: bad shape B=%d H=%d W=%d", B, H, W);
  ICGAN_REQUIRE(Cin % 16 == 0 && Cin >= 16, "icgan_conv2d_tc: Cin must be a multiple of 16 (got %d)", Cin);
  ICGAN_REQUIRE(Cout % 8 == 0 && Cout >= 8, "icgan_conv2d_tc: Cout must be a multiple of 8 (got %d)", Cout);
  ICGAN_REQUIRE(res_shift >= 0 && res_shift <= 2, "icgan_conv2d_tc: res_shift must be 0, 1 or 2 (got %d)", res_shift);
  ICGAN_REQUIRE(!(res_shift == 1 && ((H | W) & 1)), "icgan_conv2d_tc: res_shift=1 needs even H, W");
  ICGAN_REQUIRE(!(res_shift == 2 && (!residual || bn_stats)), "icgan_conv2d_tc: res_shift=2 (mask) needs a mask tensor");

  if (ksize == 3 && !bn_stats) {
    const int rc = launch_conv_halo(x, wk, alpha_dev, bias, residual, y, B, H, W, Cin, Cout, out_dtype, res_dtype,
                                    res_shift, act, static_cast<cudaStream_t>(stream));
    if (rc != kHaloIneligible) return rc;
  }

  TcConvParams p{};
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.ksz = ksize; p.pad = ksize / 2; p.taps = ksize * ksize;
  p.TW = pow2_floor(W < 128 ? W : 128);
  p.TH = pow2_floor(H < 128 / p.TW ? H : 128 / p.TW);
  p.TN = 128 / (p.TW * p.TH);
  p.tiles_w = ceil_div(W, p.TW);
  p.tiles_h = ceil_div(H, p.TH);
  const int tiles_b = ceil_div(B, p.TN);
  // K-chunk width = largest of 64/32/16 dividing Cin. (Measured: padding Cin=96 to 2x64 with TMA zero-fill is SLOWER,
  // 481 -> 392 TFLOP/s: the kernel is bound by bytes staged L2 -> smem, and padding stages 33% more.)
  p.cw = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.swz = static_cast<uint32_t>(p.cw * 2);
  p.chunks = Cin / p.cw;
  p.k_iters = p.taps * p.chunks;
  if (Cout <= 256) p.BN = (Cout + 15) / 16 * 16;
  else if (Cout % 256 == 0) p.BN = 256;
  else if (Cout % 192 == 0) p.BN = 192;
  else if (Cout % 128 == 0) p.BN = 128;
  else p.BN = 256;
  p.n_tiles = ceil_div(Cout, p.BN);
  p.total_tiles = p.tiles_w * p.tiles_h * tiles_b * p.n_tiles;
  p.a_bytes = 128u * p.swz;
  p.b_tx = static_cast<uint32_t>(p.BN) * p.swz;
  p.b_chunk = (p.b_tx + 1023u) & ~1023u;
  const uint32_t chunk_bytes = p.a_bytes + p.b_chunk;
  int G = static_cast<int>(49152u / chunk_bytes);  // aim at ~48 KB and >= 4 UMMAs per barrier round trip
  if (G < 1) G = 1;
  if (G > 8) G = 8;
  if (G > p.k_iters) G = p.k_iters;
  p.G = G;
  p.n_stage_iters = (p.k_iters + G - 1) / G;
  p.stage_bytes = static_cast<uint32_t>(G) * chunk_bytes;
  const uint32_t tail = 1024u /*align slack*/ + 512u /*barriers*/ + 4u * 32u * 33u * 4u /*stats transpose*/;
  int stages = static_cast<int>((kSmemBudget - tail) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  ICGAN_REQUIRE(stages >= 2, "icgan_conv2d_tc: tile does not fit shared memory");
  p.stages = stages;
  p.idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.BN));
  p.out_bf16 = out_dtype == ICGAN_BF16;
  p.res_bf16 = res_dtype == ICGAN_BF16;
  p.res_shift = res_shift;
  p.act = act;
  p.y = y; p.bias = bias; p.res = residual; p.alpha = alpha_dev; p.stats = bn_stats;
  for (int t = 0; t < p.taps; ++t) {
    p.tdh[t] = static_cast<int8_t>(t / ksize - p.pad);
    p.tdw[t] = static_cast<int8_t>(t % ksize - p.pad);
    p.twi[t] = static_cast<int8_t>(t);
  }
  p.in_stride = 1; p.OH = H; p.OW = W; p.osy = p.osx = 1; p.ooy = p.oox = 0;
  ICGAN_REQUIRE(!bn_stats || (act == ICGAN_ACT_NONE && Cout % 32 == 0),
                "icgan_conv2d_tc: bn_stats needs act=none and Cout a multiple of 32 (got %d)", Cout);

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {(uint32_t)p.cw, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, p.swz);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)p.taps, (uint64_t)Cout};
    const uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)p.taps * Cin * 2};
    const uint32_t box[3] = {(uint32_t)p.cw, 1u, (uint32_t)p.BN};
    int rc = make_tmap_bf16(&tmB, wk, 3, dims, str, box, p.swz);
    if (rc) return rc;
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    ICGAN_CUDA(cudaFuncSetAttribute(tc_conv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (bn_stats) tc_conv_kernel<4><<<grid, kThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  else tc_conv_kernel<8><<<grid, kConvThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

// Tile shape (TW x TH x TN = `pixels`, powers of two) that wastes the fewest padded pixels over a Wd x Hd x B domain.
static void choose_tile(int Wd, int Hd, int B, int pixels, int max_w, int* TW, int* TH, int* TN) {
  double best = 1e30;
  for (int tw = 1; tw <= pixels && tw <= max_w; tw *= 2)
    for (int th = 1; tw * th <= pixels; th *= 2) {
      const int tn = pixels / (tw * th);
      if (tw > 2 * Wd && tw > 8) continue;
      const double vol = static_cast<double>(ceil_div(Wd, tw)) * tw * ceil_div(Hd, th) * th * ceil_div(B, tn) * tn;
      // prefer wide rows at equal volume (longer contiguous runs per TMA box row)
      const double cost = vol * (1.0 + 0.001 * th + 0.002 * tn);
      if (cost < best) {
        best = cost;
        *TW = tw; *TH = th; *TN = tn;
      }
    }
}

extern "C" int icgan_conv2d_tc_ex(const void* x, const void* wk, const float* alpha_dev, const float* bias,
                                  const void* residual, void* y, int B,
                                  int H, int W, int Cin, int Cout, int wtaps, int ntaps, const int* tap_dh,
                                  const int* tap_dw, const int* tap_w, int in_stride, int Hd, int Wd, int OH, int OW,
                                  int osy, int ooy, int osx, int oox, int out_dtype, int res_dtype, int res_mask,
                                  void* stream) {
  ICGAN_REQUIRE(x && wk && y && tap_dh && tap_dw && tap_w, "icgan_conv2d_tc_ex: null pointer");
  ICGAN_REQUIRE(!res_mask || residual, "icgan_conv2d_tc_ex: res_mask needs a mask tensor");
  ICGAN_REQUIRE(ntaps >= 1 && ntaps <= 16 && wtaps >= 1 && wtaps <= 127, "icgan_conv2d_tc_ex: 1..16 taps (got %d)", ntaps);
  ICGAN_REQUIRE(in_stride == 1 || in_stride == 2, "icgan_conv2d_tc_ex: input stride 1 or 2 (got %d)", in_stride);
  ICGAN_REQUIRE(B > 0 && H > 0 && W > 0 && Hd > 0 && Wd > 0 && OH > 0 && OW > 0, "icgan_conv2d_tc_ex: bad shape");
  ICGAN_REQUIRE(Cin % 16 == 0 && Cin >= 16, "icgan_conv2d_tc_ex: Cin must be a multiple of 16 (got %d)", Cin);
  ICGAN_REQUIRE(Cout % 8 == 0 && Cout >= 8, "icgan_conv2d_tc_ex: Cout must be a multiple of 8 (got %d)", Cout);
  ICGAN_REQUIRE(osy >= 1 && osx >= 1 && ooy >= 0 && oox >= 0, "icgan_conv2d_tc_ex: bad output mapping");
  TcConvParams p{};
  p.B = B; p.H = Hd; p.W = Wd; p.Cout = Cout;
  p.ksz = 0; p.pad = 0; p.taps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    ICGAN_REQUIRE(tap_dh[t] >= -64 && tap_dh[t] <= 64 && tap_dw[t] >= -64 && tap_dw[t] <= 64 && tap_w[t] >= 0 &&
                  tap_w[t] < wtaps, "icgan_conv2d_tc_ex: tap %d out of range", t);
    p.tdh[t] = static_cast<int8_t>(tap_dh[t]);
    p.tdw[t] = static_cast<int8_t>(tap_dw[t]);
    p.twi[t] = static_cast<int8_t>(tap_w[t]);
  }
  p.in_stride = in_stride; p.OH = OH; p.OW = OW; p.osy = osy; p.ooy = ooy; p.osx = osx; p.oox = oox;
  choose_tile(Wd, Hd, B, 128, 128, &p.TW, &p.TH, &p.TN);
  p.tiles_w = ceil_div(Wd, p.TW);
  p.tiles_h = ceil_div(Hd, p.TH);
  const int tiles_b = ceil_div(B, p.TN);
  p.cw = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.swz = static_cast<uint32_t>(p.cw * 2);
  p.chunks = Cin / p.cw;
  p.k_iters = p.taps * p.chunks;
  if (Cout <= 256) p.BN = (Cout + 15) / 16 * 16;
  else if (Cout % 256 == 0) p.BN = 256;
  else if (Cout % 192 == 0) p.BN = 192;
  else if (Cout % 128 == 0) p.BN = 128;
  else p.BN = 256;
  p.n_tiles = ceil_div(Cout, p.BN);
  p.total_tiles = p.tiles_w * p.tiles_h * tiles_b * p.n_tiles;
  p.a_bytes = 128u * p.swz;
  p.b_tx = static_cast<uint32_t>(p.BN) * p.swz;
  p.b_chunk = (p.b_tx + 1023u) & ~1023u;
  const uint32_t chunk_bytes = p.a_bytes + p.b_chunk;
  int G = static_cast<int>(49152u / chunk_bytes);
  if (G < 1) G = 1;
  if (G > 8) G = 8;
  if (G > p.k_iters) G = p.k_iters;
  p.G = G;
  p.n_stage_iters = (p.k_iters + G - 1) / G;
  p.stage_bytes = static_cast<uint32_t>(G) * chunk_bytes;
  const uint32_t tail = 1024u + 512u + 4u * 32u * 33u * 4u;
  int stages = static_cast<int>((kSmemBudget - tail) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  ICGAN_REQUIRE(stages >= 2, "icgan_conv2d_tc_ex: tile does not fit shared memory");
  p.stages = stages;
  p.idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.BN));
  p.out_bf16 = out_dtype == ICGAN_BF16;
  p.res_bf16 = res_dtype == ICGAN_BF16;
  p.res_shift = res_mask ? 2 : 0;  // 2: `residual` gates the result (y = residual > 0 ? y : 0) instead of being added
  p.act = ICGAN_ACT_NONE;
  p.y = y; p.bias = bias; p.res = residual; p.alpha = alpha_dev; p.stats = nullptr;

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    // stride 2: a box spanning 2*T-1 tensor elements delivers T of them (every second one)
    const uint32_t bw = in_stride == 1 ? (uint32_t)p.TW : (uint32_t)(2 * p.TW - 1);
    const uint32_t bh = in_stride == 1 ? (uint32_t)p.TH : (uint32_t)(2 * p.TH - 1);
    const uint32_t box[4] = {(uint32_t)p.cw, bw, bh, (uint32_t)p.TN};
    const uint32_t es[4] = {1u, (uint32_t)in_stride, (uint32_t)in_stride, 1u};
    int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, p.swz, es);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)wtaps, (uint64_t)Cout};
    const uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)wtaps * Cin * 2};
    const uint32_t box[3] = {(uint32_t)p.cw, 1u, (uint32_t)p.BN};
    int rc = make_tmap_bf16(&tmB, wk, 3, dims, str, box, p.swz);
    if (rc) return rc;
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_conv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  tc_conv_kernel<8><<<grid, kConvThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_conv2d_wgrad_tc_ex(const void* a, const void* b, float* out, int B, int Ha, int Wa, int Ca, int Hb,
                                        int Wb, int Cb, int ntaps, const int* tap_dh, const int* tap_dw, int in_stride,
                                        void* stream) {
  ICGAN_REQUIRE(a && b && out && tap_dh && tap_dw, "icgan_conv2d_wgrad_tc_ex: null pointer");
  ICGAN_REQUIRE(ntaps >= 1 && ntaps <= 16, "icgan_conv2d_wgrad_tc_ex: 1..16 taps (got %d)", ntaps);
  ICGAN_REQUIRE(in_stride == 1 || in_stride == 2, "icgan_conv2d_wgrad_tc_ex: stride 1 or 2 (got %d)", in_stride);
  ICGAN_REQUIRE(Cb % 16 == 0 && Ca % 8 == 0, "icgan_conv2d_wgrad_tc_ex: need Cb%%16==0, Ca%%8==0 (got %d, %d)", Cb, Ca);
  TcWgradParams p{};
  p.Cin = Cb; p.Cout = Ca; p.ksz = 0; p.pad = 0; p.taps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    p.tdh[t] = static_cast<int8_t>(tap_dh[t]);
    p.tdw[t] = static_cast<int8_t>(tap_dw[t]);
  }
  p.in_stride = in_stride;
  choose_tile(Wa, Ha, B, kWgradKP, kWgradKP, &p.TW, &p.TH, &p.TN);
  p.tiles_w = ceil_div(Wa, p.TW);
  p.tiles_h = ceil_div(Ha, p.TH);
  p.tiles_b = ceil_div(B, p.TN);
  p.k_chunks = p.tiles_w * p.tiles_h * p.tiles_b;
  if (p.taps > 1) {
    p.ci_per_tile = 64;
    p.groups = ceil_div(p.taps, kWgradMaxBoxes);
  } else {
    p.ci_per_tile = 64 * kWgradMaxBoxes;
    p.groups = 1;
  }
  p.co_tiles = ceil_div(Ca, 128);
  p.ci_tiles = ceil_div(Cb, p.ci_per_tile);
  p.box_bytes = static_cast<uint32_t>(kWgradKP) * 128u;
  p.a_bytes = 2u * p.box_bytes;
  p.stage_bytes = p.a_bytes + static_cast<uint32_t>(kWgradMaxBoxes) * p.box_bytes;
  const uint32_t tail = 1024u + 512u;
  int stages = static_cast<int>((kSmemBudget - tail) / p.stage_bytes);
  if (stages > 8) stages = 8;
  p.stages = stages;
  const int out_tiles = p.co_tiles * p.ci_tiles * p.groups;
  p.splits = choose_splits(out_tiles, p.k_chunks);
  p.dwk = out;
  CUtensorMap tmA, tmB;
  {
    const uint32_t box[4] = {64u, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    const uint64_t dims[4] = {(uint64_t)Ca, (uint64_t)Wa, (uint64_t)Ha, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Ca * 2, (uint64_t)Wa * Ca * 2, (uint64_t)Ha * Wa * Ca * 2};
    int rc = make_tmap_bf16(&tmA, a, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    const uint32_t bw = in_stride == 1 ? (uint32_t)p.TW : (uint32_t)(2 * p.TW - 1);
    const uint32_t bh = in_stride == 1 ? (uint32_t)p.TH : (uint32_t)(2 * p.TH - 1);
    const uint32_t box[4] = {64u, bw, bh, (uint32_t)p.TN};
    const uint32_t es[4] = {1u, (uint32_t)in_stride, (uint32_t)in_stride, 1u};
    const uint64_t dims[4] = {(uint64_t)Cb, (uint64_t)Wb, (uint64_t)Hb, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cb * 2, (uint64_t)Wb * Cb * 2, (uint64_t)Hb * Wb * Cb * 2};
    int rc = make_tmap_bf16(&tmB, b, 4, dims, str, box, 128, es);
    if (rc) return rc;
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const int grid = out_tiles * p.splits;
  tc_wgrad_kernel<<<grid, kThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_conv2d_rgb_tc(const void* x, const void* wcol, const float* alpha_dev, const float* bias, void* y,
                                   int B, int H, int W, int Cs, int Cout, int ksize, int out_dtype, int act,
                                   void* stream) {
  ICGAN_REQUIRE(x && wcol && y, "icgan_conv2d_rgb_tc: null pointer");
  ICGAN_REQUIRE(B > 0 && H > 0 && W > 0 && Cs > 0, "icgan_conv2d_rgb_tc: bad shape");
  ICGAN_REQUIRE((ksize == 1 || ksize == 3) && ksize * ksize * Cs <= 32,
                "icgan_conv2d_rgb_tc: needs k*k*Cs <= 32 (got k=%d Cs=%d)", ksize, Cs);
  ICGAN_REQUIRE(Cout % 8 == 0 && Cout >= 8 && Cout <= 256, "icgan_conv2d_rgb_tc: Cout must be a multiple of 8 in [8, 256]");
  TcRgbParams p{};
  p.B = B; p.H = H; p.W = W; p.Cs = Cs; p.ksz = ksize; p.Cout = Cout;
  p.BN = (Cout + 15) / 16 * 16;
  p.P = static_cast<int64_t>(B) * H * W;
  p.total_tiles = static_cast<int>((p.P + 127) / 128);
  p.idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.BN));
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wcol = static_cast<const __nv_bfloat16*>(wcol);
  p.out_bf16 = out_dtype == ICGAN_BF16;
  p.act = act;
  p.y = y; p.bias = bias; p.alpha = alpha_dev;
  const uint32_t smem_bytes = 1024u + kRgbBufs * 8192u + 256u * 64u + 512u;
  ICGAN_REQUIRE(Cs <= 3, "icgan_conv2d_rgb_tc: at most 3 input channels (got %d)", Cs);
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_conv_rgb_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    ICGAN_CUDA(cudaFuncSetAttribute(tc_conv_rgb_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (ksize == 3 && Cs == 3)
    tc_conv_rgb_kernel<3, 3><<<grid, kRgbThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
  else
    tc_conv_rgb_kernel<0, 0><<<grid, kRgbThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_conv2d_wgrad_tc(const void* x, const void* dy, float* dwk, int B, int H, int W, int Cin,
                                     int Cout, int ksize, void* stream) {
  ICGAN_REQUIRE(x && dy && dwk, "icgan_conv2d_wgrad_tc: null pointer");
  ICGAN_REQUIRE(ksize == 1 || ksize == 3, "icgan_conv2d_wgrad_tc: ksize must be 1 or 3");
  ICGAN_REQUIRE(Cin % 16 == 0 && Cout % 8 == 0, "icgan_conv2d_wgrad_tc: need Cin%%16==0, Cout%%8==0 (got %d, %d)", Cin,
                Cout);

  if (ksize == 3) {
    const int rc = launch_wgrad_halo(x, dy, dwk, B, H, W, Cin, Cout, static_cast<cudaStream_t>(stream));
    if (rc != kWgradHaloIneligible) return rc;
  }

  TcWgradParams p{};
  p.Cin = Cin; p.Cout = Cout; p.ksz = ksize; p.pad = ksize / 2; p.taps = ksize * ksize;
  p.TW = pow2_floor(W < kWgradKP ? W : kWgradKP);
  p.TH = pow2_floor(H < kWgradKP / p.TW ? H : kWgradKP / p.TW);
  p.TN = kWgradKP / (p.TW * p.TH);
  p.tiles_w = ceil_div(W, p.TW);
  p.tiles_h = ceil_div(H, p.TH);
  p.tiles_b = ceil_div(B, p.TN);
  p.k_chunks = p.tiles_w * p.tiles_h * p.tiles_b;
  if (p.taps > 1) {
    p.ci_per_tile = 64;
    p.groups = ceil_div(p.taps, kWgradMaxBoxes);
  } else {
    p.ci_per_tile = 64 * kWgradMaxBoxes;
    p.groups = 1;
  }
  p.co_tiles = ceil_div(Cout, 128);
  p.ci_tiles = ceil_div(Cin, p.ci_per_tile);
  p.box_bytes = static_cast<uint32_t>(kWgradKP) * 128u;
  p.a_bytes = 2u * p.box_bytes;
  p.stage_bytes = p.a_bytes + static_cast<uint32_t>(kWgradMaxBoxes) * p.box_bytes;
  const uint32_t tail = 1024u + 512u;
  int stages = static_cast<int>((kSmemBudget - tail) / p.stage_bytes);
  if (stages > 8) stages = 8;
  ICGAN_REQUIRE(stages >= 2, "icgan_conv2d_wgrad_tc: tile does not fit shared memory");
  p.stages = stages;
  const int out_tiles = p.co_tiles * p.ci_tiles * p.groups;
  static const int old_splits = env_int("ICGAN_TC_WGRAD_OLD_SPLITS", 0);
  if (old_splits) {
    int splits = ceil_div(2 * num_sms(), out_tiles);
    if (splits > p.k_chunks) splits = p.k_chunks;
    p.splits = splits < 1 ? 1 : splits;
  } else {
    p.splits = choose_splits(out_tiles, p.k_chunks);
  }
  p.dwk = dwk;
  for (int t = 0; t < p.taps; ++t) {
    p.tdh[t] = static_cast<int8_t>(t / ksize - p.pad);
    p.tdw[t] = static_cast<int8_t>(t % ksize - p.pad);
  }
  p.in_stride = 1;

  CUtensorMap tmDy, tmX;
  const uint32_t box[4] = {64u, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    int rc = make_tmap_bf16(&tmDy, dy, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    int rc = make_tmap_bf16(&tmX, x, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  const uint32_t smem_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes + tail;
  static unsigned long long configured = 0ull;
  if (first_use_on_this_device(&configured)) {
    ICGAN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const int grid = out_tiles * p.splits;
  tc_wgrad_kernel<<<grid, kThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmDy, tmX, p);
  ICGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int icgan_nhwc_to_cnhw(const void* x, void* xT, int64_t pixels, int C, int in_dtype, void* stream) {
  ICGAN_REQUIRE(x && xT && pixels > 0 && C > 0, "icgan_nhwc_to_cnhw: bad arguments");
  dim3 grid(static_cast<unsigned>((pixels + 31) / 32), static_cast<unsigned>((C + 31) / 32));
  dim3 block(32, 8);
  if (in_dtype == ICGAN_BF16)
    nhwc_to_cnhw_kernel<__nv_bfloat16><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(xT), pixels, C);
  else
    nhwc_to_cnhw_kernel<float><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(x), static_cast<__nv_bfloat16*>(xT), pixels, C);
  ICGAN_LAUNCH_CHECK();
  return 0;
}
