// gemm_tc.cu — persistent, warp-specialised tcgen05 implicit-GEMM convolution / linear kernel for sm_100a.
//
//   warp 0 (1 thread) : TMA producer  — cp.async.bulk.tensor 4-D (activations, shifted per tap, OOB zero fill
//                                         = conv padding) + 2-D (weights) into a SWIZZLE_128B smem ring
//   warp 1 (1 thread) : MMA issuer    — tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32 in TMEM), 128 x BN x 16 per
//                                         instruction, tcgen05.commit releases smem slots / publishes accumulators
//   warp 2            : TMEM allocator (alloc / dealloc)
//   warps 4..7        : epilogue      — tcgen05.ld 32x32b, fused bias / per-image vector / residual / scale /
//                                         activation (SiLU, leaky-ReLU, GEGLU) / bf16 hi-lo split, direct stores
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// See include/tango_b200.h (tng_conv_gemm) for the operator contract and the reference call sites it replaces.
#include "tng_ptx.cuh"
#include "tng_internal.h"

namespace tng {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;

struct KGroupDev {
  int view, a_c0, dw, dh, b_k0, nkb;
};

struct GemmKernelParams {
  // output pixel grid and M tiling
  int W, H, NB;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int m_tiles, n_tiles;
  int Ncols;
  int n_groups, total_kiters;
  KGroupDev g[TNG_MAX_KGROUPS];
  // epilogue
  const float* bias;
  const float* rowvec;
  long long rowvec_ld;
  const void* res;
  int res_bf16;
  long long ldr;
  float alpha;
  int accumulate;
  float* out_f32;
  long long ld_f32;
  __nv_bfloat16* out_bf16;
  long long ld_bf16;
  int act;
  float act_param;
  int split_off;
  int vec_ok;  // all row strides / bases allow 16-byte vector access
};

template <int BN>
struct GemmCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int ACC_STRIDE = TMEM_COLS / 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float x, int act, float p) {
  if (act == TNG_ACT_SILU) return silu_f(x);
  if (act == TNG_ACT_LRELU) return x > 0.f ? x : x * p;
  return x;
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap amap0, const __grid_constant__ CUtensorMap amap1,
               const __grid_constant__ CUtensorMap amap2, const __grid_constant__ CUtensorMap amap3,
               const __grid_constant__ CUtensorMap bmap, const __grid_constant__ GemmKernelParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap0);
    tma_prefetch_desc(&bmap);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    const CUtensorMap* amaps[4] = {&amap0, &amap1, &amap2, &amap3};
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tm = tile / p.n_tiles, tn = tile % p.n_tiles;
      const int tw = tm % p.tiles_w;
      const int th = (tm / p.tiles_w) % p.tiles_h;
      const int tb = tm / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tb * p.bn;
      for (int gi = 0; gi < p.n_groups; ++gi) {
        const KGroupDev g = p.g[gi];
        const CUtensorMap* am = amaps[g.view];
        for (int kb = 0; kb < g.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_4d(sA + stage * A_TILE_BYTES, am, &full_bar[stage], g.a_c0 + kb * BK, w0 + g.dw, h0 + g.dh, n0);
          tma_load_2d(sB + stage * Cfg::B_TILE_BYTES, &bmap, &full_bar[stage], g.b_k0 + kb * BK, tn * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
      for (int ki = 0; ki < p.total_kiters; ++ki) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * A_TILE_BYTES), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzled row: +2 in the (>>4) start-address field
          umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull_bar[as]);
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int ew = warp & 3;           // TMEM lane quarter this warp may access
    const int r = ew * 32 + lane;      // row inside the 128-row tile
    const bool geglu = (p.act == TNG_ACT_GEGLU);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / p.n_tiles, tn = tile % p.n_tiles;
      const int tw = tm % p.tiles_w;
      const int th = (tm / p.tiles_w) % p.tiles_h;
      const int tb = tm / (p.tiles_w * p.tiles_h);
      const int w = tw * p.bw + (r % p.bw);
      const int h = th * p.bh + (r / p.bw) % p.bh;
      const int img = tb * p.bn + r / (p.bw * p.bh);
      const bool row_ok = (w < p.W) && (h < p.H) && (img < p.NB);
      const long long row = (static_cast<long long>(img) * p.H + h) * p.W + w;

      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * Cfg::ACC_STRIDE + (static_cast<uint32_t>(ew * 32) << 16);

      if (!geglu) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          const int col0 = tn * BN + c;
          if (!row_ok || col0 >= p.Ncols) continue;
          const int ncol = min(32, p.Ncols - col0);
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          const bool vec = p.vec_ok && (ncol == 32);
          if (p.bias) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                x[j] += b4.x; x[j + 1] += b4.y; x[j + 2] += b4.z; x[j + 3] += b4.w;
              }
            } else {
              for (int j = 0; j < ncol; ++j) x[j] += __ldg(p.bias + col0 + j);
            }
          }
          if (p.rowvec) {
            const float* rv = p.rowvec + static_cast<long long>(img) * p.rowvec_ld + col0;
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + j));
                x[j] += b4.x; x[j + 1] += b4.y; x[j + 2] += b4.z; x[j + 3] += b4.w;
              }
            } else {
              for (int j = 0; j < ncol; ++j) x[j] += __ldg(rv + j);
            }
          }
          if (p.res) {
            if (p.res_bf16) {
              const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + row * p.ldr + col0;
              if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const uint4 u = *reinterpret_cast<const uint4*>(rp + j);
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float2 f = __bfloat1622float2(h2[q]);
                    x[j + 2 * q] += f.x; x[j + 2 * q + 1] += f.y;
                  }
                }
              } else {
                for (int j = 0; j < ncol; ++j) x[j] += __bfloat162float(rp[j]);
              }
            } else {
              const float* rp = reinterpret_cast<const float*>(p.res) + row * p.ldr + col0;
              if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(rp + j);
                  x[j] += b4.x; x[j + 1] += b4.y; x[j + 2] += b4.z; x[j + 3] += b4.w;
                }
              } else {
                for (int j = 0; j < ncol; ++j) x[j] += rp[j];
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] *= p.alpha;
          if (p.out_f32) {
            float* op = p.out_f32 + row * p.ld_f32 + col0;
            if (vec) {
              if (p.accumulate) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 o4 = *reinterpret_cast<const float4*>(op + j);
                  x[j] += o4.x; x[j + 1] += o4.y; x[j + 2] += o4.z; x[j + 3] += o4.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
            } else {
              for (int j = 0; j < ncol; ++j) {
                if (p.accumulate) x[j] += op[j];
                op[j] = x[j];
              }
            }
          }
          if (p.out_bf16) {
            __nv_bfloat16* op = p.out_bf16 + row * p.ld_bf16 + col0;
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = apply_act(x[j], p.act, p.act_param);
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16(y[j], y[j + 1]); u.y = pack_bf16(y[j + 2], y[j + 3]);
                u.z = pack_bf16(y[j + 4], y[j + 5]); u.w = pack_bf16(y[j + 6], y[j + 7]);
                *reinterpret_cast<uint4*>(op + j) = u;
              }
              if (p.split_off > 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  float l[8];
#pragma unroll
                  for (int q = 0; q < 8; ++q) l[q] = y[j + q] - __bfloat162float(__float2bfloat16_rn(y[j + q]));
                  uint4 u;
                  u.x = pack_bf16(l[0], l[1]); u.y = pack_bf16(l[2], l[3]);
                  u.z = pack_bf16(l[4], l[5]); u.w = pack_bf16(l[6], l[7]);
                  *reinterpret_cast<uint4*>(op + p.split_off + j) = u;
                }
              }
            } else {
              for (int j = 0; j < ncol; ++j) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(y[j]);
                op[j] = hi;
                if (p.split_off > 0) op[p.split_off + j] = __float2bfloat16_rn(y[j] - __bfloat162float(hi));
              }
            }
          }
        }
      } else {
        // GEGLU: columns [0, BN/2) of the tile are "hidden", [BN/2, BN) the matching "gate" (weights interleaved
        // on the host). out[:, tn*BN/2 + j] = (hid + b) * gelu_erf(gate + b')
        constexpr int HALF = BN / 2;
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
          uint32_t vh[32], vg[32];
          __syncwarp();
          tmem_ld32(taddr + c, vh);
          tmem_ld32(taddr + HALF + c, vg);
          tmem_ld_wait();
          if (!row_ok) continue;
          const int gcol = tn * BN + c;  // GEMM column of hidden; gate at gcol + HALF
          const int ocol = tn * HALF + c;
          __nv_bfloat16* op = p.out_bf16 + row * p.ld_bf16 + ocol;
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float hv = __uint_as_float(vh[j]);
            float gv = __uint_as_float(vg[j]);
            if (p.bias) {
              hv += __ldg(p.bias + gcol + j);
              gv += __ldg(p.bias + gcol + HALF + j);
            }
            y[j] = hv * gelu_erf_f(gv);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16(y[j], y[j + 1]); u.y = pack_bf16(y[j + 2], y[j + 3]);
            u.z = pack_bf16(y[j + 4], y[j + 5]); u.w = pack_bf16(y[j + 6], y[j + 7]);
            *reinterpret_cast<uint4*>(op + j) = u;
          }
          if (p.split_off > 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float l[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) l[q] = y[j + q] - __bfloat162float(__float2bfloat16_rn(y[j + q]));
              uint4 u;
              u.x = pack_bf16(l[0], l[1]); u.y = pack_bf16(l[2], l[3]);
              u.z = pack_bf16(l[4], l[5]); u.w = pack_bf16(l[6], l[7]);
              *reinterpret_cast<uint4*>(op + p.split_off + j) = u;
            }
          }
        }
      }
      // all tcgen05.ld of this warp are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <int BN>
static int launch_gemm(const CUtensorMap* am, const CUtensorMap& bm, const GemmKernelParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(gemm_tc<%d>): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_tc_kernel<BN><<<grid, 256, Cfg::SMEM_BYTES, st>>>(am[0], am[1], am[2], am[3], bm, p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TNG_ECUDA, "gemm_tc<%d> launch: %s", BN, cudaGetErrorString(e));
  return TNG_OK;
}

static bool is_pow2(long long x) { return x > 0 && (x & (x - 1)) == 0; }

}  // namespace tng

using namespace tng;

extern "C" int tng_conv_gemm(const tng_gemm_desc* d, void* stream) {
  if (!d) return set_error(TNG_EINVAL, "null desc");
  if (d->n_aviews < 1 || d->n_aviews > TNG_MAX_AVIEWS) return set_error(TNG_EINVAL, "n_aviews=%d", d->n_aviews);
  if (d->n_groups < 1 || d->n_groups > TNG_MAX_KGROUPS) return set_error(TNG_EINVAL, "n_groups=%d", d->n_groups);
  if (d->W <= 0 || d->H <= 0 || d->NB <= 0 || d->Ncols <= 0) return set_error(TNG_EINVAL, "bad output grid");
  if ((d->ldb > 0 ? d->ldb : d->Ktot) % 8 != 0)
    return set_error(TNG_EINVAL, "B row stride must be a multiple of 8 elements (Ktot=%lld ldb=%lld)", (long long)d->Ktot, (long long)d->ldb);

  GemmKernelParams p;
  memset(&p, 0, sizeof(p));
  p.W = d->W; p.H = d->H; p.NB = d->NB;
  // M tile = bw x bh x bn output pixels (product 128)
  if (d->W >= BM || d->H == 1) {
    p.bw = BM; p.bh = 1; p.bn = 1;
  } else {
    if (!is_pow2(d->W)) return set_error(TNG_EINVAL, "W=%d < 128 must be a power of two", d->W);
    p.bw = d->W;
    const int rem = BM / p.bw;
    if (d->H >= rem) {
      p.bh = rem; p.bn = 1;
    } else {
      if (!is_pow2(d->H)) return set_error(TNG_EINVAL, "H=%d (W=%d) must be a power of two when W*H < 128", d->H, d->W);
      p.bh = d->H; p.bn = rem / p.bh;
    }
  }
  p.tiles_w = (d->W + p.bw - 1) / p.bw;
  p.tiles_h = (d->H + p.bh - 1) / p.bh;
  p.tiles_n = (d->NB + p.bn - 1) / p.bn;
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.Ncols = (int)d->Ncols;

  int bn_tile = d->block_n;
  if (d->act == TNG_ACT_GEGLU) {
    if (bn_tile == 0) bn_tile = (d->Ncols % 256 == 0) ? 256 : 128;
    if ((bn_tile != 128 && bn_tile != 256) || d->Ncols % bn_tile != 0 || !d->out_bf16 || d->out_f32 || d->res ||
        d->rowvec)
      return set_error(TNG_EINVAL, "GEGLU epilogue needs block_n 128/256 dividing Ncols, bf16 output only");
  }
  if (bn_tile == 0) {
    const long long N = d->Ncols;
    if (N <= 32) bn_tile = 32;
    else if (N <= 64) bn_tile = 64;
    else if (N % 256 == 0 && (long long)p.m_tiles * (N / 256) >= 2 * num_sms()) bn_tile = 256;
    else if (N % 160 == 0) bn_tile = 160;
    else if (N % 128 == 0) bn_tile = 128;
    else if (N % 64 == 0 && N < 256) bn_tile = 64;
    else bn_tile = 128;
  }
  p.n_tiles = (int)((d->Ncols + bn_tile - 1) / bn_tile);

  p.n_groups = d->n_groups;
  p.total_kiters = 0;
  for (int i = 0; i < d->n_groups; ++i) {
    const tng_kgroup& g = d->g[i];
    if (g.view < 0 || g.view >= d->n_aviews || g.nkb <= 0) return set_error(TNG_EINVAL, "k-group %d invalid", i);
    // The last K block may run past Ktot / the view's channel count: TMA zero-fills the out-of-range part of A,
    // so the (possibly non-zero) B columns read there contribute nothing.
    if (g.b_k0 < 0 || g.b_k0 + (long long)(g.nkb - 1) * BK >= d->Ktot)
      return set_error(TNG_EINVAL, "k-group %d: K block outside B", i);
    if (g.a_c0 < 0 || g.a_c0 + (long long)(g.nkb - 1) * BK >= d->a[g.view].C)
      return set_error(TNG_EINVAL, "k-group %d: K block outside view channels", i);
    p.g[i] = KGroupDev{g.view, g.a_c0, g.dw, g.dh, g.b_k0, g.nkb};
    p.total_kiters += g.nkb;
  }
  p.bias = d->bias; p.rowvec = d->rowvec; p.rowvec_ld = d->rowvec_ld > 0 ? d->rowvec_ld : d->Ncols; p.res = d->res; p.res_bf16 = (d->res_dtype == TNG_DT_BF16);
  p.ldr = d->ldr; p.alpha = d->alpha; p.accumulate = d->accumulate;
  p.out_f32 = d->out_f32; p.ld_f32 = d->ld_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d->out_bf16); p.ld_bf16 = d->ld_bf16;
  p.act = d->act; p.act_param = d->act_param; p.split_off = d->split_off;
  if (!d->out_f32 && !d->out_bf16) return set_error(TNG_EINVAL, "no output");
  if (d->accumulate && !d->out_f32) return set_error(TNG_EINVAL, "accumulate needs out_f32");

  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  bool vec = true;
  if (d->bias && !al16(d->bias)) vec = false;
  if (d->rowvec && (!al16(d->rowvec) || p.rowvec_ld % 4)) vec = false;
  if (d->res) {
    if (!al16(d->res)) vec = false;
    if (p.res_bf16 ? (d->ldr % 8) : (d->ldr % 4)) vec = false;
  }
  if (d->out_f32 && (!al16(d->out_f32) || d->ld_f32 % 4)) vec = false;
  if (d->out_bf16 && (!al16(d->out_bf16) || d->ld_bf16 % 8 || d->split_off % 8)) vec = false;
  p.vec_ok = vec ? 1 : 0;
  if (d->act == TNG_ACT_GEGLU && !vec) return set_error(TNG_EINVAL, "GEGLU epilogue needs 16-byte aligned output");

  // tensor maps
  CUtensorMap am[4];
  for (int i = 0; i < 4; ++i) {
    const tng_aview& v = d->a[i < d->n_aviews ? i : 0];
    if (v.C % 8 != 0) return set_error(TNG_EINVAL, "view %d: C=%lld must be a multiple of 8", i, (long long)v.C);
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.NB};
    uint64_t strides[3] = {(uint64_t)v.s_w * 2, (uint64_t)v.s_h * 2, (uint64_t)v.s_n * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap_bf16(&am[i], v.ptr, 4, dims, strides, box);
    if (rc) return rc;
  }
  CUtensorMap bm;
  {
    uint64_t dims[2] = {(uint64_t)d->Ktot, (uint64_t)d->Ncols};
    uint64_t strides[1] = {(uint64_t)(d->ldb > 0 ? d->ldb : d->Ktot) * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)bn_tile};
    int rc = encode_tmap_bf16(&bm, d->b, 2, dims, strides, box);
    if (rc) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn_tile) {
    case 32: return launch_gemm<32>(am, bm, p, st);
    case 64: return launch_gemm<64>(am, bm, p, st);
    case 128: return launch_gemm<128>(am, bm, p, st);
    case 160: return launch_gemm<160>(am, bm, p, st);
    case 256: return launch_gemm<256>(am, bm, p, st);
    default: return set_error(TNG_EINVAL, "block_n=%d unsupported", bn_tile);
  }
}
