// gemm_tc.cu — persistent, warp-specialised tcgen05 implicit-GEMM convolution / linear kernel for sm_100a.
//
//   warp 0 (1 thread) : TMA producer  — cp.async.bulk.tensor 4-D (activations, shifted per tap, OOB zero fill
//                                         = conv padding) + 2-D (weights) into a SWIZZLE_128B smem ring
//   warp 1 (1 thread) : MMA issuer    — tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32 in TMEM), 128 x BN x 16 per
//                                         instruction, tcgen05.commit releases smem slots / publishes accumulators
//   warp 2            : TMEM allocator (alloc / dealloc)
//   warps 4..11       : epilogue (8 warps, two per TMEM lane quarter, alternating 32-column chunks) —
//                       tcgen05.ld 32x32b -> XOR-swizzled smem transpose -> fully coalesced global traffic (8 lanes own
//                       one 128-byte row segment) with fused bias / per-image vector / residual / scale / accumulate /
//                       activation (SiLU, leaky-ReLU, GEGLU) / bf16 hi-lo split. Eight warps (two per scheduler) and
//                       loads-before-use give the epilogue the memory- and instruction-level parallelism it needs to
//                       stay hidden behind the MMA main loop of the next tile (two TMEM accumulator stages).
//
// See include/tango_b200.h (tng_conv_gemm) for the operator contract and the reference call sites it replaces.
#include "tng_ptx.cuh"
#include "tng_internal.h"
#include <stdlib.h>

namespace tng {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARPS = 8;

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
  int ksplit;    // 1, or 2: two CTAs share an output tile, each reduces half of the K iterations and red.adds fp32 partials
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
  int vec_ok;    // all row strides / bases allow 16-byte vector access
  int fast_epi;  // vec_ok && Ncols % 4 == 0
  // GroupNorm statistics of the fp32 output, emitted from the epilogue (full-tile launches only, see the host side):
  // col_stats[(img * Ncols + col) * 2 + {0, 1}] += sum / sum of squares over the rows of image img = row / stats_hw
  double* col_stats;
  long long stats_hw;
};

// PAIR: cta_group::2 — each CTA of the pair stages its own 128 A rows and only HALF of the weight tile.
// NH = 2 (pair mode only): the pair owns a 256 x (2 BN) output tile — two BN-wide accumulators fed from the SAME A
// stage — which halves the operand bytes each SM pulls from L2 per FLOP once more (the measured main-loop limiter).
template <int BN, bool PAIR = false, int NH = 1>
struct GemmCfg {
  static constexpr int B_HALF_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // one BN-wide weight tile (this CTA's share)
  static constexpr int B_TILE_BYTES = NH * B_HALF_BYTES;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;  // per epilogue warp: 32 x 32 fp32 swizzled transpose tile
  static constexpr int STAGES_RAW = (227 * 1024 - EPI_BYTES - 256) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (4 * BN <= 128) ? 128 : (4 * BN <= 256) ? 256 : 512;  // one CTA per SM: take what helps
  // accumulator slots (BN columns each) in TMEM: as many as fit (2 for N tile 256, 3 for 160, 4 for <= 128): extra
  // slots absorb the wake-up latency of the epilogue warps when the main loop of a tile is short (K = 320 linears).
  // With NH = 2 a tile takes two consecutive slots of the ring (3 slots at BN = 160: the next tile's main loop starts
  // as soon as the epilogue has drained the first half of the current one).
  static constexpr int NACC = (TMEM_COLS / BN) > 4 ? 4 : (TMEM_COLS / BN);
  static_assert(NH == 1 || NACC >= 3, "two-accumulator tiles need a ring of at least three slots");
  static constexpr int ACC_STRIDE = BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float x, int act, float p) {
  if (act == TNG_ACT_SILU) return silu_f(x);
  if (act == TNG_ACT_LRELU) return x > 0.f ? x : x * p;
  return x;
}

struct EpiRows {
  long long row[8];
  int img[8];
  uint32_t valid;
};

// One 32-row x 32-column chunk, already staged (swizzled) in `st`: lanes (rsub = lane >> 3, cg = lane & 7) own the
// 4 columns [4cg, 4cg+4) of rows 4i + rsub, i = 0..7.
template <bool RES, bool F32, bool BF16, bool VEC>
__device__ __forceinline__ void epi_chunk(const GemmKernelParams& p, const float* st, const EpiRows& R, int rsub, int cg,
                                          int col) {
  const bool col_ok = col < p.Ncols;
  const bool has_res = RES && (p.res != nullptr);
  const bool has_acc = F32 && (p.accumulate != 0);
  const bool has_f32 = F32 && (p.out_f32 != nullptr);
  const bool has_bf = BF16 && (p.out_bf16 != nullptr);
  // ---- all global loads of the chunk first (independent -> in flight together)
  float4 rres[8], rold[8];
  if (has_res) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok && ((R.valid >> i) & 1)) {
        if (VEC) {
          if (p.res_bf16) {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + R.row[i] * p.ldr + col);
            const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            r4 = make_float4(f0.x, f0.y, f1.x, f1.y);
          } else {
            r4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + R.row[i] * p.ldr + col);
          }
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j)
            if (col + j < p.Ncols)
              t[j] = p.res_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res)[R.row[i] * p.ldr + col + j])
                                : reinterpret_cast<const float*>(p.res)[R.row[i] * p.ldr + col + j];
          r4 = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      rres[i] = r4;
    }
  }
  if (has_acc) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok && ((R.valid >> i) & 1)) {
        if (VEC) {
          r4 = *reinterpret_cast<const float4*>(p.out_f32 + R.row[i] * p.ld_f32 + col);
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j)
            if (col + j < p.Ncols) t[j] = p.out_f32[R.row[i] * p.ld_f32 + col + j];
          r4 = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      rold[i] = r4;
    }
  }
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias && col_ok) {
    if (VEC) {
      b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    } else {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < 4; ++j)
        if (col + j < p.Ncols) t[j] = __ldg(p.bias + col + j);
      b4 = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  const bool has_rv = (p.rowvec != nullptr);
  const float alpha = p.alpha;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + rsub;
    float4 a = *reinterpret_cast<const float4*>(st + r * 32 + ((cg ^ (r & 7)) << 2));
    if (!col_ok || !((R.valid >> i) & 1)) continue;
    a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
    if (has_rv) {
      const float* rv = p.rowvec + static_cast<long long>(R.img[i]) * p.rowvec_ld + col;
      if (VEC) {
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(rv));
        a.x += r4.x; a.y += r4.y; a.z += r4.z; a.w += r4.w;
      } else {
        if (col + 0 < p.Ncols) a.x += __ldg(rv + 0);
        if (col + 1 < p.Ncols) a.y += __ldg(rv + 1);
        if (col + 2 < p.Ncols) a.z += __ldg(rv + 2);
        if (col + 3 < p.Ncols) a.w += __ldg(rv + 3);
      }
    }
    if (has_res) { a.x += rres[i].x; a.y += rres[i].y; a.z += rres[i].z; a.w += rres[i].w; }
    a.x *= alpha; a.y *= alpha; a.z *= alpha; a.w *= alpha;
    if (has_f32) {
      if (has_acc) { a.x += rold[i].x; a.y += rold[i].y; a.z += rold[i].z; a.w += rold[i].w; }
      float* op = p.out_f32 + R.row[i] * p.ld_f32 + col;
      if (VEC) {
        *reinterpret_cast<float4*>(op) = a;
      } else {
        const float t[4] = {a.x, a.y, a.z, a.w};
        for (int j = 0; j < 4; ++j)
          if (col + j < p.Ncols) op[j] = t[j];
      }
    }
    if (has_bf) {
      __nv_bfloat16* op = p.out_bf16 + R.row[i] * p.ld_bf16 + col;
      const float y0 = apply_act(a.x, p.act, p.act_param), y1 = apply_act(a.y, p.act, p.act_param);
      const float y2 = apply_act(a.z, p.act, p.act_param), y3 = apply_act(a.w, p.act, p.act_param);
      if (VEC) {
        uint2 u;
        u.x = pack_bf16(y0, y1); u.y = pack_bf16(y2, y3);
        *reinterpret_cast<uint2*>(op) = u;
        if (p.split_off > 0) {
          uint2 l;
          l.x = pack_bf16(y0 - __bfloat162float(__float2bfloat16_rn(y0)), y1 - __bfloat162float(__float2bfloat16_rn(y1)));
          l.y = pack_bf16(y2 - __bfloat162float(__float2bfloat16_rn(y2)), y3 - __bfloat162float(__float2bfloat16_rn(y3)));
          *reinterpret_cast<uint2*>(op + p.split_off) = l;
        }
      } else {
        const float t[4] = {y0, y1, y2, y3};
        for (int j = 0; j < 4; ++j) {
          if (col + j < p.Ncols) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(t[j]);
            op[j] = hi;
            if (p.split_off > 0) op[p.split_off + j] = __float2bfloat16_rn(t[j] - __bfloat162float(hi));
          }
        }
      }
    }
  }
}

// Lean path for FULL tiles (all 128 rows valid, all 32 columns of the chunk < Ncols, 16-byte aligned): the rows of a
// tile are consecutive output rows (the host tiling guarantees it), so slot i of a lane is row r0 + 4i and every
// pointer advances by a constant stride — a few instructions per 16-byte access, no per-element predicates.
template <bool RES, bool F32, bool BF16>
__device__ __forceinline__ void epi_chunk_full(const GemmKernelParams& p, const float* st, long long r0, int img0,
                                               int rsub, int cg, int col, bool rv_uniform, long long stats_img) {
  float4 add4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) add4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  if (p.rowvec && rv_uniform) {
    const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.rowvec + static_cast<long long>(img0) * p.rowvec_ld + col));
    add4.x += r4.x; add4.y += r4.y; add4.z += r4.z; add4.w += r4.w;
  }
  float4 rres[8];
  if (RES) {
    if (p.res_bf16) {
      const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + r0 * p.ldr + col;
      const long long rs = 4 * p.ldr;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint2 u = *reinterpret_cast<const uint2*>(rp + i * rs);
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        rres[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
      }
    } else {
      const float* rp = reinterpret_cast<const float*>(p.res) + r0 * p.ldr + col;
      const long long rs = 4 * p.ldr;
#pragma unroll
      for (int i = 0; i < 8; ++i) rres[i] = *reinterpret_cast<const float4*>(rp + i * rs);
    }
  }
  float4 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + rsub;
    a[i] = *reinterpret_cast<const float4*>(st + r * 32 + ((cg ^ (r & 7)) << 2));
    a[i].x += add4.x; a[i].y += add4.y; a[i].z += add4.z; a[i].w += add4.w;
  }
  if (p.rowvec && !rv_uniform) {
    const int rpi = p.bw * p.bh;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int im = img0 + (4 * i) / rpi;  // img0 is the image of slot 0; rows advance by 4 per slot
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.rowvec + static_cast<long long>(im) * p.rowvec_ld + col));
      a[i].x += r4.x; a[i].y += r4.y; a[i].z += r4.z; a[i].w += r4.w;
    }
  }
  if (RES) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i].x += rres[i].x; a[i].y += rres[i].y; a[i].z += rres[i].z; a[i].w += rres[i].w; }
  }
  if (p.alpha != 1.0f) {
    const float al = p.alpha;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i].x *= al; a[i].y *= al; a[i].z *= al; a[i].w *= al; }
  }
  if (F32) {
    float* op = p.out_f32 + r0 * p.ld_f32 + col;
    const long long os = 4 * p.ld_f32;
    if (p.accumulate) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 o4 = *reinterpret_cast<const float4*>(op + i * os);
        a[i].x += o4.x; a[i].y += o4.y; a[i].z += o4.z; a[i].w += o4.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(op + i * os) = a[i];
  }
  if (p.col_stats) {
    // Column sums of this warp's 32 rows x 32 columns (all rows belong to image stats_img): 8 rows in registers,
    // then across the four row-lanes (lane bits 3 and 4); lanes 0..7 hold the totals of their 4 columns and add
    // them to the fp64 per-(image, channel) accumulators — fp32 partials over 32 values, fp64 across tiles.
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0 += a[i].x; s1 += a[i].y; s2 += a[i].z; s3 += a[i].w;
      q0 = fmaf(a[i].x, a[i].x, q0); q1 = fmaf(a[i].y, a[i].y, q1);
      q2 = fmaf(a[i].z, a[i].z, q2); q3 = fmaf(a[i].w, a[i].w, q3);
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
      q2 += __shfl_xor_sync(0xffffffffu, q2, o); q3 += __shfl_xor_sync(0xffffffffu, q3, o);
    }
    if (rsub == 0) {
      double* sp = p.col_stats + (stats_img * p.Ncols + col) * 2;
      atomicAdd(sp + 0, static_cast<double>(s0)); atomicAdd(sp + 1, static_cast<double>(q0));
      atomicAdd(sp + 2, static_cast<double>(s1)); atomicAdd(sp + 3, static_cast<double>(q1));
      atomicAdd(sp + 4, static_cast<double>(s2)); atomicAdd(sp + 5, static_cast<double>(q2));
      atomicAdd(sp + 6, static_cast<double>(s3)); atomicAdd(sp + 7, static_cast<double>(q3));
    }
  }
  if (BF16) {
    if (p.act == TNG_ACT_SILU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i].x = silu_f(a[i].x); a[i].y = silu_f(a[i].y); a[i].z = silu_f(a[i].z); a[i].w = silu_f(a[i].w); }
    } else if (p.act == TNG_ACT_LRELU) {
      const float sl = p.act_param;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i].x = a[i].x > 0.f ? a[i].x : a[i].x * sl; a[i].y = a[i].y > 0.f ? a[i].y : a[i].y * sl;
        a[i].z = a[i].z > 0.f ? a[i].z : a[i].z * sl; a[i].w = a[i].w > 0.f ? a[i].w : a[i].w * sl;
      }
    }
    __nv_bfloat16* op = p.out_bf16 + r0 * p.ld_bf16 + col;
    const long long os = 4 * p.ld_bf16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint2 u;
      u.x = pack_bf16(a[i].x, a[i].y); u.y = pack_bf16(a[i].z, a[i].w);
      *reinterpret_cast<uint2*>(op + i * os) = u;
    }
    if (p.split_off > 0) {
      op += p.split_off;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint2 l;
        l.x = pack_bf16(a[i].x - __bfloat162float(__float2bfloat16_rn(a[i].x)), a[i].y - __bfloat162float(__float2bfloat16_rn(a[i].y)));
        l.y = pack_bf16(a[i].z - __bfloat162float(__float2bfloat16_rn(a[i].z)), a[i].w - __bfloat162float(__float2bfloat16_rn(a[i].w)));
        *reinterpret_cast<uint2*>(op + i * os) = l;
      }
    }
  }
}

__device__ __forceinline__ void epi_stage(float* st, int lane, const uint32_t* v) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<uint4*>(st + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

template <int BN, bool RES, bool F32, bool BF16>
__device__ __forceinline__ void epi_tile_full(const GemmKernelParams& p, float* st, long long r0, int img0, bool rv_uniform,
                                              uint32_t taddr, int tn, int lane, int hf, long long stats_img) {
  const int cg = lane & 7, rsub = lane >> 3;
#pragma unroll 1
  for (int c = hf * 32; c < BN; c += 64) {
    uint32_t v[32];
    tmem_ld32(taddr + c, v);
    tmem_ld_wait();
    __syncwarp();  // previous chunk's smem reads are complete
    epi_stage(st, lane, v);
    __syncwarp();
    epi_chunk_full<RES, F32, BF16>(p, st, r0, img0, rsub, cg, tn * BN + c + 4 * cg, rv_uniform, stats_img);
  }
}

template <int BN, bool VEC>
__device__ __noinline__ void epi_tile_generic(const GemmKernelParams& p, float* st, const EpiRows& R, uint32_t taddr, int tn,
                                              int lane, int hf) {
  const int cg = lane & 7, rsub = lane >> 3;
#pragma unroll 1
  for (int c = hf * 32; c < BN; c += 64) {
    uint32_t v[32];
    tmem_ld32(taddr + c, v);
    tmem_ld_wait();
    __syncwarp();
    epi_stage(st, lane, v);
    __syncwarp();
    epi_chunk<true, true, true, VEC>(p, st, R, rsub, cg, tn * BN + c + 4 * cg);
  }
}

// Split-K epilogue: this CTA holds the partial sum over its half of K. out (+)= alpha * (partial [+ bias + rowvec + res
// for the first half only]) with fp32 red.adds into an output the host zeroed beforehand. With exactly two partials
// per element the result does not depend on their order (0 + a + b, fp32 addition commutes), so runs stay
// reproducible. Used for under-filled launches with a long reduction (the 32x2 level of the UNet): the epilogue is
// small next to the main loop, so this is the simple row-slot form.
template <int BN>
__device__ __noinline__ void epi_tile_splitk(const GemmKernelParams& p, float* st, long long row_base, int n0, int rpi,
                                             int nvalid, int sp, uint32_t taddr, int tn, int lane, int ew, int hf) {
  const int cg = lane & 7, rsub = lane >> 3;
#pragma unroll 1
  for (int c = hf * 32; c < BN; c += 64) {
    uint32_t v[32];
    tmem_ld32(taddr + c, v);
    tmem_ld_wait();
    __syncwarp();
    epi_stage(st, lane, v);
    __syncwarp();
    const int col = tn * BN + c + 4 * cg;
    const bool col_ok = col < p.Ncols;   // Ncols % 4 == 0 (checked on the host)
    float4 add4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && sp == 0 && p.bias) add4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int rr = ew * 32 + 4 * i + rsub;
      if (!col_ok || rr >= nvalid) continue;
      const long long row = row_base + rr;
      float4 a = *reinterpret_cast<const float4*>(st + (4 * i + rsub) * 32 + ((cg ^ ((4 * i + rsub) & 7)) << 2));
      if (sp == 0) {
        a.x += add4.x; a.y += add4.y; a.z += add4.z; a.w += add4.w;
        if (p.rowvec) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.rowvec + static_cast<long long>(n0 + rr / rpi) * p.rowvec_ld + col));
          a.x += r4.x; a.y += r4.y; a.z += r4.z; a.w += r4.w;
        }
        if (p.res) {
          if (p.res_bf16) {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + row * p.ldr + col);
            const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
          } else {
            const float4 r4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + row * p.ldr + col);
            a.x += r4.x; a.y += r4.y; a.z += r4.z; a.w += r4.w;
          }
        }
      }
      a.x *= p.alpha; a.y *= p.alpha; a.z *= p.alpha; a.w *= p.alpha;
      float* op = p.out_f32 + row * p.ld_f32 + col;
      atomicAdd(op, a.x); atomicAdd(op + 1, a.y); atomicAdd(op + 2, a.z); atomicAdd(op + 3, a.w);
    }
  }
}

// GEGLU: columns [0, BN/2) of the tile are "hidden", [BN/2, BN) the matching "gate" (weights interleaved on the
// host). out[:, tn*BN/2 + j] = (hid + b) * gelu_erf(gate + b'). Rows r0 + 4i (consecutive-row tiles); rows >= nvalid
// are skipped.
template <int BN, bool FULL, bool SPLIT, bool TANH>
__device__ __forceinline__ void epi_tile_geglu(const GemmKernelParams& p, float* st, long long r0, int nleft,
                                               uint32_t taddr, int tn, int lane, int hf) {
  constexpr int HALF = BN / 2;
  const int cg = lane & 7, rsub = lane >> 3;
#pragma unroll 1
  for (int c = hf * 32; c < HALF; c += 64) {
    uint32_t v[32];
    float4 hid[8], g[8];
    tmem_ld32(taddr + c, v);
    tmem_ld_wait();
    __syncwarp();
    epi_stage(st, lane, v);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + rsub;
      hid[i] = *reinterpret_cast<const float4*>(st + r * 32 + ((cg ^ (r & 7)) << 2));
    }
    tmem_ld32(taddr + HALF + c, v);
    tmem_ld_wait();
    __syncwarp();
    epi_stage(st, lane, v);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + rsub;
      g[i] = *reinterpret_cast<const float4*>(st + r * 32 + ((cg ^ (r & 7)) << 2));
    }
    const int gcol = tn * BN + c + 4 * cg;  // GEMM column of the hidden half; gate at + HALF
    const int ocol = tn * HALF + c + 4 * cg;
    float4 bh = make_float4(0.f, 0.f, 0.f, 0.f), bg = bh;
    if (p.bias) {
      bh = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
      bg = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + HALF));
    }
    // branch-free arithmetic over all 32 values of this lane (32 independent MUFU chains to interleave)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hid[i].x = (hid[i].x + bh.x) * (TANH ? gelu_tanh_f(g[i].x + bg.x) : gelu_erf_f(g[i].x + bg.x));
      hid[i].y = (hid[i].y + bh.y) * (TANH ? gelu_tanh_f(g[i].y + bg.y) : gelu_erf_f(g[i].y + bg.y));
      hid[i].z = (hid[i].z + bh.z) * (TANH ? gelu_tanh_f(g[i].z + bg.z) : gelu_erf_f(g[i].z + bg.z));
      hid[i].w = (hid[i].w + bh.w) * (TANH ? gelu_tanh_f(g[i].w + bg.w) : gelu_erf_f(g[i].w + bg.w));
    }
    __nv_bfloat16* op = p.out_bf16 + r0 * p.ld_bf16 + ocol;
    const long long os = 4 * p.ld_bf16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (!FULL && 4 * i >= nleft) continue;
      uint2 u;
      u.x = pack_bf16(hid[i].x, hid[i].y); u.y = pack_bf16(hid[i].z, hid[i].w);
      *reinterpret_cast<uint2*>(op + i * os) = u;
      if (SPLIT) {
        uint2 l;
        l.x = pack_bf16(hid[i].x - __bfloat162float(__float2bfloat16_rn(hid[i].x)),
                        hid[i].y - __bfloat162float(__float2bfloat16_rn(hid[i].y)));
        l.y = pack_bf16(hid[i].z - __bfloat162float(__float2bfloat16_rn(hid[i].z)),
                        hid[i].w - __bfloat162float(__float2bfloat16_rn(hid[i].w)));
        *reinterpret_cast<uint2*>(op + p.split_off + i * os) = l;
      }
    }
  }
}

// CL = 1: single CTA. CL = 2: cluster of 2 along M with TMA multicast of the weight tile. CL = 3: CTA PAIR
// (tcgen05 cta_group::2): one 256 x BN MMA spans both SMs, each SM stages its 128 A rows + half of B, so the operand
// bytes each SM has to ingest per FLOP drop by ~28 % (the measured main-loop limiter); only the leader issues MMAs.
// (original note) CL = thread-block-cluster size along M (1 or 2). With CL = 2 the two CTAs of a cluster work on adjacent M tiles of
// the same N tile and share the weight tile: each loads half of B and TMA-multicasts it into both CTAs, halving the
// L2 -> SM operand traffic for B (the main-loop bottleneck at N tile 160); smem slots are released by multicast commits.
template <int BN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap amap0, const __grid_constant__ CUtensorMap amap1,
               const __grid_constant__ CUtensorMap amap2, const __grid_constant__ CUtensorMap amap3,
               const __grid_constant__ CUtensorMap bmap, const __grid_constant__ GemmKernelParams p) {
  constexpr bool PAIR = (CL == 3 || CL == 4);
  constexpr int NH = (CL == 4) ? 2 : 1;      // BN-wide accumulators per tile (CL = 4: 256 x 2 BN pair tiles)
  constexpr int CLUSTER = (CL == 1) ? 1 : 2;
  using Cfg = GemmCfg<BN, PAIR, NH>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment (checked below)
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_TILE_BYTES;
  float* sEpi = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  constexpr int NACC = Cfg::NACC;
  uint64_t* tfull_bar = bars + 2 * STAGES;          // [NACC]
  uint64_t* tempty_bar = bars + 2 * STAGES + NACC;  // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[tng] gemm_tc: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&amap0);
    tma_prefetch_desc(&bmap);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL == 2 ? 2 : 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], PAIR ? 2 * EPI_WARPS : EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = (CLUSTER > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  if (CLUSTER > 1) cluster_sync_all();  // peer barriers are initialised before any multicast copy / commit targets them

  // work items: (pair of adjacent M tiles) x N tile; CTA `crank` of the cluster takes M tile 2*mp + crank
  const int m_groups = (p.m_tiles + CLUSTER - 1) / CLUSTER;
  const int total_tiles = m_groups * p.n_tiles * p.ksplit;   // split-K: consecutive work items = the K halves of one tile
  const int work0 = blockIdx.x / CLUSTER, work_stride = gridDim.x / CLUSTER;

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = work0; tile < total_tiles; tile += work_stride) {
      const int sp = tile % p.ksplit, t2 = tile / p.ksplit;
      const int tm = (t2 / p.n_tiles) * CLUSTER + crank, tn = t2 % p.n_tiles;
      const int tw = tm % p.tiles_w;
      const int th = (tm / p.tiles_w) % p.tiles_h;
      const int tb = tm / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tb * p.bn;
      // flat K-iteration range of this work item (all of them unless split-K)
      const int k_lo = sp * p.total_kiters / p.ksplit, k_hi = (sp + 1) * p.total_kiters / p.ksplit;
      int kbase = 0;
      for (int gi = 0; gi < p.n_groups; ++gi) {
        const KGroupDev g = p.g[gi];
        const CUtensorMap* am = g.view == 0 ? &amap0 : g.view == 1 ? &amap1 : g.view == 2 ? &amap2 : &amap3;
        const int kb_lo = max(0, k_lo - kbase), kb_hi = min(g.nkb, k_hi - kbase);
        kbase += g.nkb;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (PAIR) {
            // both CTAs load into their own smem; all bytes are credited to the LEADER's full barrier
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_4d_2sm(sA + stage * A_TILE_BYTES, am, &full_bar[stage], g.a_c0 + kb * BK, w0 + g.dw, h0 + g.dh, n0);
#pragma unroll
            for (int nh = 0; nh < NH; ++nh)
              tma_load_2d_2sm(sB + stage * Cfg::B_TILE_BYTES + nh * Cfg::B_HALF_BYTES, &bmap, &full_bar[stage],
                              g.b_k0 + kb * BK, (tn * NH + nh) * BN + crank * (BN / 2));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_4d(sA + stage * A_TILE_BYTES, am, &full_bar[stage], g.a_c0 + kb * BK, w0 + g.dw, h0 + g.dh, n0);
          if (CL == 1) {
            tma_load_2d(sB + stage * Cfg::B_TILE_BYTES, &bmap, &full_bar[stage], g.b_k0 + kb * BK, tn * BN);
          } else {
            // my half of the weight tile, written into BOTH CTAs (each CTA's full barrier sees A + both halves)
            constexpr int HALF_ROWS = BN / 2;
            tma_load_2d_mc(sB + stage * Cfg::B_TILE_BYTES + crank * HALF_ROWS * 128, &bmap, &full_bar[stage],
                           g.b_k0 + kb * BK, tn * BN + crank * HALF_ROWS, static_cast<uint16_t>(3));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && (!PAIR || crank == 0)) {
    // ===================================================== MMA issuer (pair mode: leader CTA only)
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = work0; tile < total_tiles; tile += work_stride, ++it) {
      // accumulator slot of half-tile j = it * NH + nh: j % NACC, in its (j / NACC)-th use
      uint32_t d_tmem[NH];
#pragma unroll
      for (int nh = 0; nh < NH; ++nh) {
        const int j = it * NH + nh;
        mbar_wait(&tempty_bar[j % NACC], ((j / NACC) & 1) ^ 1);
        d_tmem[nh] = tmem_base + (j % NACC) * Cfg::ACC_STRIDE;
      }
      tc_fence_after();
      const int sp = tile % p.ksplit;
      const int nk = (sp + 1) * p.total_kiters / p.ksplit - sp * p.total_kiters / p.ksplit;
      for (int ki = 0; ki < nk; ++ki) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * A_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int nh = 0; nh < NH; ++nh) {
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_TILE_BYTES + nh * Cfg::B_HALF_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzled row: +2 in the (>>4) start-address field
            if (PAIR) umma_bf16_2cta(d_tmem[nh], adesc + 2 * k, bdesc + 2 * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem[nh], adesc + 2 * k, bdesc + 2 * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
          }
        }
        if (CL == 1) umma_commit(&empty_bar[stage]);
        else if (CL == 2) umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>(3));  // frees the slot in both CTAs
        else umma_commit2_mc(&empty_bar[stage], static_cast<uint16_t>(3));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int nh = 0; nh < NH; ++nh) {
        const int j = it * NH + nh;
        if (PAIR) umma_commit2_mc(&tfull_bar[j % NACC], static_cast<uint16_t>(3));  // each CTA drains its own 128 rows
        else umma_commit(&tfull_bar[j % NACC]);
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (8 warps)
    const int ew = warp & 3;            // TMEM lane quarter this warp may access
    const int hf = (warp - 4) >> 2;     // which half of the 32-column chunks
    float* st = sEpi + (warp - 4) * (32 * 32);
    const int rsub = lane >> 3;
    const int mode = (p.res ? 1 : 0) | (p.out_f32 ? 2 : 0) | (p.out_bf16 ? 4 : 0);
    const bool geglu = (p.act == TNG_ACT_GEGLU || p.act == TNG_ACT_GEGLU_TANH);
    int it = 0;
    for (int tile = work0; tile < total_tiles; tile += work_stride, ++it) {
      const int sp = tile % p.ksplit, t2 = tile / p.ksplit;
      const int tm = (t2 / p.n_tiles) * CLUSTER + crank, tn0 = t2 % p.n_tiles;
      const int tw = tm % p.tiles_w;
      const int th = (tm / p.tiles_w) % p.tiles_h;
      const int tb = tm / (p.tiles_w * p.tiles_h);
      // rows of a tile are consecutive output rows; valid rows form a prefix (see host tiling)
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tb * p.bn;
      const long long row_base = (static_cast<long long>(n0) * p.H + h0) * p.W + w0;
      int nvalid;
      if (p.bh == 1 && p.bn == 1) nvalid = min(BM, p.W - w0);
      else if (p.bn == 1) nvalid = min(p.bh, p.H - h0) * p.bw;
      else nvalid = min(p.bn, p.NB - n0) * p.bh * p.bw;
      const int rpi = p.bw * p.bh;  // rows of one image inside a tile
      if (tm >= p.m_tiles) nvalid = 0;  // padding tile of an odd cluster tail
#pragma unroll 1
      for (int nh = 0; nh < NH; ++nh) {
      // half-tile j: accumulator slot j % NACC in its (j / NACC)-th use; output columns of N tile tn * NH + nh
      const int j = it * NH + nh;
      const int as = j % NACC;
      const uint32_t aphase = (j / NACC) & 1;
      const int tn = tn0 * NH + nh;
      const uint32_t taddr = tmem_base + as * Cfg::ACC_STRIDE + (static_cast<uint32_t>(ew * 32) << 16);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const bool full = p.fast_epi && (nvalid == BM) && ((tn + 1) * BN <= p.Ncols);
      if (p.ksplit > 1) {
        epi_tile_splitk<BN>(p, st, row_base, n0, rpi, nvalid, sp, taddr, tn, lane, ew, hf);
      } else if (geglu) {
        // slot i of this lane is row ew*32 + rsub + 4i; rows below nvalid are valid
        if constexpr (BN == 128 || BN == 256) {   // the host only selects these N tiles for GEGLU
          const long long gr0 = row_base + ew * 32 + rsub;
          const int nleft = nvalid - (ew * 32 + rsub);
          if (p.act == TNG_ACT_GEGLU_TANH) {   // T5 front-end (small): one general instantiation
            if (p.split_off > 0) epi_tile_geglu<BN, false, true, true>(p, st, gr0, nleft, taddr, tn, lane, hf);
            else epi_tile_geglu<BN, false, false, true>(p, st, gr0, nleft, taddr, tn, lane, hf);
          } else if (p.split_off > 0) epi_tile_geglu<BN, false, true, false>(p, st, gr0, nleft, taddr, tn, lane, hf);
          else if (nvalid == BM) epi_tile_geglu<BN, true, false, false>(p, st, gr0, nleft, taddr, tn, lane, hf);
          else epi_tile_geglu<BN, false, false, false>(p, st, gr0, nleft, taddr, tn, lane, hf);
        }
      } else if (!full) {
        EpiRows R;
        R.valid = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = ew * 32 + 4 * i + rsub;
          R.img[i] = n0 + rr / rpi;
          R.row[i] = row_base + rr;
          if (rr < nvalid) R.valid |= 1u << i;
        }
        if (p.fast_epi) epi_tile_generic<BN, true>(p, st, R, taddr, tn, lane, hf);
        else epi_tile_generic<BN, false>(p, st, R, taddr, tn, lane, hf);
      } else {
        const long long r0 = row_base + ew * 32 + rsub;
        const int img0 = n0 + (ew * 32 + rsub) / rpi;
        const bool rv_uniform = (rpi % 32) == 0;   // the warp's 32 rows lie in one image
        // image of this warp's 32 consecutive output rows for the GroupNorm statistics (stats_hw % 32 == 0: host check)
        const long long simg = p.col_stats ? (row_base + ew * 32) / p.stats_hw : 0;
        switch (mode) {
          case 2: epi_tile_full<BN, false, true, false>(p, st, r0, img0, rv_uniform, taddr, tn, lane, hf, simg); break;
          case 3: epi_tile_full<BN, true, true, false>(p, st, r0, img0, rv_uniform, taddr, tn, lane, hf, simg); break;
          case 4: epi_tile_full<BN, false, false, true>(p, st, r0, img0, rv_uniform, taddr, tn, lane, hf, simg); break;
          case 5: epi_tile_full<BN, true, false, true>(p, st, r0, img0, rv_uniform, taddr, tn, lane, hf, simg); break;
          case 6: epi_tile_full<BN, false, true, true>(p, st, r0, img0, rv_uniform, taddr, tn, lane, hf, simg); break;
          default: epi_tile_full<BN, true, true, true>(p, st, r0, img0, rv_uniform, taddr, tn, lane, hf, simg); break;
        }
      }
      // all tcgen05.ld of this warp are complete (wait::ld): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote(&tempty_bar[as], 0);  // the leader's MMA issuer owns the accumulators
        else mbar_arrive(&tempty_bar[as]);
      }
      }  // nh
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // the peer may still multicast into / arrive on this CTA until it is done too
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <int BN, int CL>
static int launch_gemm_cl(const CUtensorMap* am, const CUtensorMap& bm, const GemmKernelParams& p, cudaStream_t st) {
  constexpr int CLUSTER = (CL == 1) ? 1 : 2;
  using Cfg = GemmCfg<BN, (CL == 3 || CL == 4), (CL == 4) ? 2 : 1>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(gemm_tc<%d,%d>): %s", BN, CL, cudaGetErrorString(e));
    attr_set = true;
  }
  const int work = ((p.m_tiles + CLUSTER - 1) / CLUSTER) * p.n_tiles * p.ksplit;
  const int slots = num_sms() / CLUSTER;
  const int grid = (work < slots ? work : slots) * CLUSTER;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CL>, am[0], am[1], am[2], am[3], bm, p);
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TNG_ECUDA, "gemm_tc<%d,%d> launch: %s", BN, CL, cudaGetErrorString(e));
  return TNG_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap* am, const CUtensorMap& bm, const GemmKernelParams& p, cudaStream_t st, int cl) {
  if (cl == 4) {
    if constexpr (BN == 160 || BN == 128) return launch_gemm_cl<BN, 4>(am, bm, p, st);
    else return set_error(TNG_EINVAL, "two-accumulator pair tiles exist for N tiles 128 and 160 only");
  }
  if (cl == 3) return launch_gemm_cl<BN, 3>(am, bm, p, st);
  return cl == 2 ? launch_gemm_cl<BN, 2>(am, bm, p, st) : launch_gemm_cl<BN, 1>(am, bm, p, st);
}

static bool is_pow2(long long x) { return x > 0 && (x & (x - 1)) == 0; }

}  // namespace tng

using namespace tng;

// Everything that is decided before a launch: argument checks, the M tiling, the N tile, split-K, the launch mode and
// whether the GroupNorm statistics ride in the epilogue. Shared by tng_conv_gemm and tng_gemm_plan.
static int plan_gemm(const tng_gemm_desc* d, GemmKernelParams& p, int& bn_tile_out, int& cl_out, bool& stats_after_out) {
  if (!d) return set_error(TNG_EINVAL, "null desc");
  if (d->n_aviews < 1 || d->n_aviews > TNG_MAX_AVIEWS) return set_error(TNG_EINVAL, "n_aviews=%d", d->n_aviews);
  if (d->n_groups < 1 || d->n_groups > TNG_MAX_KGROUPS) return set_error(TNG_EINVAL, "n_groups=%d", d->n_groups);
  if (d->W <= 0 || d->H <= 0 || d->NB <= 0 || d->Ncols <= 0) return set_error(TNG_EINVAL, "bad output grid");
  if ((d->ldb > 0 ? d->ldb : d->Ktot) % 8 != 0)
    return set_error(TNG_EINVAL, "B row stride must be a multiple of 8 elements (Ktot=%lld ldb=%lld)", (long long)d->Ktot, (long long)d->ldb);

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
  int ksplit = 1;
  if (d->act == TNG_ACT_GEGLU || d->act == TNG_ACT_GEGLU_TANH) {
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
    else if (N % 160 == 0) {
      bn_tile = 160;
      // under-filled launches (e.g. the 32x2 level of the UNet: 8 M tiles)
      if ((long long)p.m_tiles * (N / 160) * 2 <= num_sms()) {
        static int splitk = -1;   // TNG_GEMM_SPLITK=0 disables (A/B measurements)
        if (splitk < 0) { const char* e = getenv("TNG_GEMM_SPLITK"); splitk = e ? atoi(e) : 1; }
        long long kit = 0;
        for (int i = 0; i < d->n_groups; ++i) kit += d->g[i].nkb;
        const bool can_split = splitk && d->out_f32 && !d->out_bf16 && !d->accumulate && d->act == TNG_ACT_NONE &&
                               d->res != d->out_f32 && kit >= 32 && d->Ncols % 4 == 0;
        if (can_split) ksplit = 2;   // two CTAs per output tile, each reduces half of K (long reductions only)
        else if (N % 128 == 0 && (long long)p.m_tiles * (N / 128) <= num_sms()) bn_tile = 128;  // more, smaller N tiles
      }
    }
    else if (N % 128 == 0) bn_tile = 128;
    else if (N % 64 == 0 && N < 256) bn_tile = 64;
    else bn_tile = 128;
  }
  p.n_tiles = (int)((d->Ncols + bn_tile - 1) / bn_tile);
  p.ksplit = ksplit;

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
  p.fast_epi = (vec && d->Ncols % 4 == 0) ? 1 : 0;
  if ((d->act == TNG_ACT_GEGLU || d->act == TNG_ACT_GEGLU_TANH) && !vec) return set_error(TNG_EINVAL, "GEGLU epilogue needs 16-byte aligned output");
  if (d->gn_stats) {
    if (d->stats_hw <= 0 || (static_cast<long long>(d->W) * d->H * d->NB) % d->stats_hw != 0)
      return set_error(TNG_EINVAL, "gn_stats: the output rows must be whole images of stats_hw pixels");
    if (!d->out_f32 && (d->split_off > 0 || d->act != TNG_ACT_NONE))
      return set_error(TNG_EINVAL, "gn_stats without an fp32 output needs a plain bf16 output (no activation, no hi/lo split)");
  }

  // Launch mode. 1 = one CTA per SM; 4 = CTA pair (tcgen05 cta_group::2) on a 256 x (2 x bn_tile) output tile: each SM
  // stages its 128 A rows and HALF of two weight tiles per K block, i.e. half the L2 -> SM operand bytes per FLOP of
  // mode 1 — the measured main-loop limiter of the long-reduction 3x3 convolutions. Chosen when the reduction is long
  // enough (>= 36 K blocks: measured, a K = 1280 linear with an fp32 residual is slower in this mode) to hide the
  // half-overlapped epilogue, every tile is full and the launch fills at least half the pairs.
  // Modes 2 (cluster-of-2 weight multicast) and 3 (pair on a 256 x bn_tile tile) are kept for A/B measurements
  // (TNG_GEMM_CLUSTER = 1..3 forces a mode; bench.py refuses to run with it set).
  const bool full_m = (p.bh == 1 && p.bn == 1) ? (d->W % BM == 0) : (p.bn == 1 ? (d->H % p.bh == 0) : (d->NB % p.bn == 0));
  int cl = 1;
  {
    static int force = -1;
    if (force < 0) { const char* e = getenv("TNG_GEMM_CLUSTER"); force = e ? atoi(e) : 0; }
    const bool geglu = d->act == TNG_ACT_GEGLU || d->act == TNG_ACT_GEGLU_TANH;
    const bool pair2_ok = (bn_tile == 160 || bn_tile == 128) && d->Ncols % (2 * bn_tile) == 0 && full_m &&
                          p.m_tiles % 2 == 0 && p.fast_epi && !geglu && p.ksplit == 1 && p.total_kiters >= 36 &&
                          static_cast<long long>(p.m_tiles / 2) * (d->Ncols / (2 * bn_tile)) * 2 >= num_sms() / 2;
    if (force == 0 || force == 4) cl = pair2_ok ? 4 : 1;
    else if (force >= 1 && force <= 3) cl = force;
    if (p.m_tiles < 2 || bn_tile < 64) cl = 1;
    if (cl != 1) p.ksplit = 1;   // split-K is implemented for the single-CTA mode only
    else if (p.ksplit > 1 && !p.fast_epi) p.ksplit = 1;
    if (cl == 4) p.n_tiles = static_cast<int>(d->Ncols / (2 * bn_tile));
  }
  // GroupNorm statistics ride in the epilogue when every tile is full (the lean epilogue path), the warp's 32 rows lie
  // in one image and the output is written exactly once; otherwise a separate pass over the output follows the GEMM
  bool stats_after = false;
  if (d->gn_stats) {
    const bool fused = full_m && (d->Ncols % bn_tile == 0) && p.fast_epi && p.ksplit == 1 && !d->accumulate &&
                       (d->stats_hw % 32 == 0) && d->act != TNG_ACT_GEGLU && d->act != TNG_ACT_GEGLU_TANH;
    if (fused) { p.col_stats = d->gn_stats; p.stats_hw = d->stats_hw; }
    else stats_after = true;
  }
  bn_tile_out = bn_tile;
  cl_out = cl;
  stats_after_out = stats_after;
  return TNG_OK;
}

extern "C" int tng_gemm_plan(const tng_gemm_desc* d, int32_t* block_n, int32_t* mode, int32_t* ksplit) {
  GemmKernelParams p;
  int bn_tile = 0, cl = 1;
  bool stats_after = false;
  const int rc = plan_gemm(d, p, bn_tile, cl, stats_after);
  if (rc != TNG_OK) return rc;
  if (block_n) *block_n = bn_tile;
  if (mode) *mode = cl;
  if (ksplit) *ksplit = p.ksplit;
  return TNG_OK;
}

extern "C" int tng_conv_gemm(const tng_gemm_desc* d, void* stream) {
  GemmKernelParams p;
  int bn_tile = 0, cl = 1;
  bool stats_after = false;
  {
    const int rc = plan_gemm(d, p, bn_tile, cl, stats_after);
    if (rc != TNG_OK) return rc;
  }
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
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(cl == 1 ? bn_tile : bn_tile / 2)};
    int rc = encode_tmap_bf16(&bm, d->b, 2, dims, strides, box);
    if (rc) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.ksplit > 1) {   // the partial sums are red.added into a zeroed output
    const size_t rows = static_cast<size_t>(d->W) * d->H * d->NB;
    cudaError_t e = cudaMemset2DAsync(d->out_f32, static_cast<size_t>(d->ld_f32) * 4, 0, static_cast<size_t>(d->Ncols) * 4, rows, st);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaMemset2DAsync(split-K output): %s", cudaGetErrorString(e));
  }
  int rc;
  switch (bn_tile) {
    case 32: rc = launch_gemm<32>(am, bm, p, st, cl); break;
    case 64: rc = launch_gemm<64>(am, bm, p, st, cl); break;
    case 128: rc = launch_gemm<128>(am, bm, p, st, cl); break;
    case 160: rc = launch_gemm<160>(am, bm, p, st, cl); break;
    case 256: rc = launch_gemm<256>(am, bm, p, st, cl); break;
    default: return set_error(TNG_EINVAL, "block_n=%d unsupported", bn_tile);
  }
  if (rc == TNG_OK && stats_after) {
    const long long rows = static_cast<long long>(d->W) * d->H * d->NB;
    // (from the stored output: fp32 when there is one, else the bf16 output — the rounded values the consumer reads)
    if (d->out_f32) rc = launch_col_stats(d->out_f32, TNG_DT_F32, d->Ncols, d->ld_f32, rows / d->stats_hw, d->stats_hw, d->gn_stats, st);
    else rc = launch_col_stats(d->out_bf16, TNG_DT_BF16, d->Ncols, d->ld_bf16, rows / d->stats_hw, d->stats_hw, d->gn_stats, st);
  }
  return rc;
}
