// attention_tc.cu — tcgen05 flash attention for head width 64 (UNet self- and cross-attention), sm_100a.
//
// One CTA = 128 query rows of one (batch, head); 128 threads, thread r owns query row r (= TMEM lane r).
// Per 128-key tile j:
//   S = Q K_j^T    tcgen05.mma 128x128x16, K-major A/B straight from TMA SWIZZLE_128B tiles, fp32 S in TMEM
//   softmax        pass 1: row max of the raw scores (tcgen05.ld); pass 2: p = exp2(s*scale*log2e - m) -> bf16 ->
//                  swizzled smem (the A operand of the next MMA); row sum in fp32. The reference max m is only
//                  advanced when the row max grew by more than 2^8 ("lazy rescale"), so the common tile does no
//                  correction work at all.
//   O += P V_j     tcgen05.mma 128x64x16 accumulating IN TMEM (V consumed as an MN-major B operand, no transpose)
// When the reference max does move, the warp rescales its O rows in place (tcgen05.ld -> scale -> tcgen05.st).
// Issue order on the tensor pipe is  P V_j , Q K_{j+1}^T  back to back right after the softmax of tile j, so the
// S-ready barrier of tile j+1 also certifies that P V_j has drained (P smem and O are safe to touch).
// Resources are sized for TWO CTAs per SM (112 KB smem, 256 TMEM columns): while one CTA is in its softmax the other
// one's MMAs keep the tensor pipe busy. K/V tiles are double-buffered and TMA-prefetched one tile ahead.
// NSPLIT = 2 is the parity mode: every operand carries its bf16 rounding residual and each product is evaluated as
// hi*hi + lo*hi + hi*lo, which restores ~fp32 accuracy on the bf16 tensor cores (1 CTA / SM).
#include "tng_ptx.cuh"
#include "tng_internal.h"
#include <stdlib.h>

namespace tng {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per tile
constexpr int AT_D = 64;     // head width
constexpr int AT_CHUNK = AT_BM * 64 * 2;  // one [128][64] bf16 swizzled chunk = 16 KB
constexpr int AT_NBUF = 2;
constexpr float AT_LAZY = 8.0f;  // rescale only when the row max grew by more than 2^8

struct AttnParams {
  int Lq, Lk, heads;
  int q_col0, q_lo_off, k_col0, k_lo_off, v_col0, v_lo_off;
  const float* kbias;
  __nv_bfloat16* out;
  long long ld_o;
  int split_off;
  float scale_log2e;  // scale * log2(e)
};

template <int NSPLIT>
struct AttnCfg {
  static constexpr int Q_BYTES = NSPLIT * AT_CHUNK;
  static constexpr int KV_BYTES = NSPLIT * AT_CHUNK;      // each of K and V per buffer
  static constexpr int P_BYTES = NSPLIT * 2 * AT_CHUNK;   // [128][128] hi (+ lo)
  static constexpr int SMEM_BYTES = Q_BYTES + AT_NBUF * 2 * KV_BYTES + P_BYTES + 128 + AT_BN * 4;  // + key-bias row
  static constexpr int TMEM_COLS = 256;                   // S: 128, O: 64
};

template <int NSPLIT>
__global__ void __launch_bounds__(128, (NSPLIT == 1) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                    const __grid_constant__ CUtensorMap vmap, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NSPLIT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;                 // [NBUF][KV_BYTES]
  uint8_t* sV = sK + AT_NBUF * Cfg::KV_BYTES;      // [NBUF][KV_BYTES]
  uint8_t* sP = sV + AT_NBUF * Cfg::KV_BYTES;      // hi chunks 0,1 then lo chunks 0,1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* bar_q = bars;             // [1]
  uint64_t* bar_kv = bars + 1;        // [NBUF]
  uint64_t* bar_s = bars + 1 + AT_NBUF;  // [1]  S_j ready (and P V_{j-1} drained)
  uint64_t* bar_o = bars + 2 + AT_NBUF;  // [1]  final P V drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + AT_NBUF);
  float* sbias = reinterpret_cast<float*>(sP + Cfg::P_BYTES + 128);  // per-tile key bias (log2 domain), -inf = masked

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int q0 = blockIdx.x * AT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.Lk + AT_BN - 1) / AT_BN;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[tng] attention: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&qmap);
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&vmap);
    mbar_init(bar_q, 1);
    for (int i = 0; i < AT_NBUF; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base;
  const uint32_t tm_o = tmem_base + 128;

  auto load_kv = [&](int tile) {
    const int buf = tile % AT_NBUF;
    mbar_arrive_expect_tx(&bar_kv[buf], 2 * Cfg::KV_BYTES);
    const int kv0 = tile * AT_BN;
#pragma unroll
    for (int s = 0; s < NSPLIT; ++s) {
      tma_load_3d(sK + buf * Cfg::KV_BYTES + s * AT_CHUNK, &kmap, &bar_kv[buf], p.k_col0 + s * p.k_lo_off + head * AT_D,
                  kv0, b);
      tma_load_3d(sV + buf * Cfg::KV_BYTES + s * AT_CHUNK, &vmap, &bar_kv[buf], p.v_col0 + s * p.v_lo_off + head * AT_D,
                  kv0, b);
    }
  };
  // S = Q K_tile^T  (hi*hi [+ lo*hi + hi*lo]); arrives on bar_s when it (and everything issued before) is done
  auto issue_qk = [&](int tile) {
    const int buf = tile % AT_NBUF;
    constexpr uint32_t idesc = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + buf * Cfg::KV_BYTES);
    constexpr int NT = (NSPLIT == 1) ? 1 : 3;
    const int qsel[3] = {0, 1, 0}, ksel[3] = {0, 0, 1};
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const uint64_t adesc = umma_desc_sw128(qa + qsel[t] * AT_CHUNK, 16, 1024);
      const uint64_t bdesc = umma_desc_sw128(ka + ksel[t] * AT_CHUNK, 16, 1024);
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) {
        umma_bf16(tm_s, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
        acc = 1;
      }
    }
    umma_commit(bar_s);
  };
  // O (+)= P V_tile   (A = P K-major chunks, B = V MN-major: 16 keys per MMA = 2 swizzle atoms = 2048 B)
  auto issue_pv = [&](int tile, uint32_t accumulate, int nk16) {
    const int buf = tile % AT_NBUF;
    constexpr uint32_t idesc = umma_idesc_bf16(AT_BM, AT_D, 0, 1);
    const uint32_t pa = smem_u32(sP), va = smem_u32(sV + buf * Cfg::KV_BYTES);
    constexpr int NT = (NSPLIT == 1) ? 1 : 3;
    const int psel[3] = {0, 1, 0}, vsel[3] = {0, 0, 1};
    uint32_t acc = accumulate;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int k = 0; k < AT_BN / 16; ++k) {
        if (k >= nk16) break;  // keys beyond the written P columns (short last tile)
        const uint64_t adesc =
            umma_desc_sw128(pa + psel[t] * 2 * AT_CHUNK + (k >> 2) * AT_CHUNK, 16, 1024) + 2 * (k & 3);
        const uint64_t bdesc = umma_desc_sw128(va + vsel[t] * AT_CHUNK + k * 2048, 1024, 1024);
        umma_bf16(tm_o, adesc, bdesc, idesc, acc);
        acc = 1;
      }
    }
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
    for (int s = 0; s < NSPLIT; ++s)
      tma_load_3d(sQ + s * AT_CHUNK, &qmap, bar_q, p.q_col0 + s * p.q_lo_off + head * AT_D, q0, b);
    for (int t = 0; t < AT_NBUF && t < n_tiles; ++t) load_kv(t);
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_kv[0], 0);
    tc_fence_after();
    issue_qk(0);
  }

  const int r = tid;  // query row owned by this thread == TMEM lane
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t ts = tm_s + lane_addr;
  const uint32_t to = tm_o + lane_addr;
  float m_ref = -INFINITY;  // reference max used in the exponent (log2 domain)
  float l_run = 0.f;
  const float sc = p.scale_log2e;
  const float* kb = p.kbias ? p.kbias + static_cast<long long>(b) * p.Lk : nullptr;
  constexpr float LOG2E = 1.4426950408889634f;
  uint8_t* prow_base = sP + r * 128;
  const int rsw = r & 7;

  for (int j = 0; j < n_tiles; ++j) {
    __syncwarp();
    mbar_wait(bar_s, j & 1);   // S_j complete; in-order tensor pipe => P V_{j-1} complete as well
    tc_fence_after();
    // K/V buffer of tile j-1 is free now: prefetch tile j+1 into it (tile j+1 == (j-1) + NBUF)
    if (tid == 0 && j >= 1 && j + 1 < n_tiles) load_kv(j + 1);
    const int kv0 = j * AT_BN;
    const bool tail = (kb != nullptr) || (kv0 + AT_BN > p.Lk);
    // columns actually processed in this tile (multiple of 32); P columns beyond are never written nor multiplied
    const int ncols = tail ? min(AT_BN, ((p.Lk - kv0) + 31) & ~31) : AT_BN;
    if (tail) {
      // one bias value per key of the tile, shared through smem: additive mask bias (already * log2e) or -inf
      // for keys beyond Lk -> the softmax passes below need no per-element predicates or global loads
      const int kv = kv0 + tid;
      float bv = -INFINITY;
      if (kv < p.Lk) bv = kb ? kb[kv] * LOG2E : 0.f;
      __syncthreads();  // previous tile's readers are done
      sbias[tid] = bv;
      __syncthreads();
    }
    // ---- pass 1: row maximum (log2 domain)
    float m_tile = -INFINITY;
    if (!tail) {
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_ld_wait();
        // four independent max chains (a single serial chain of 128 dependent FMNMX costs ~4 cycles each)
        float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i]));
          m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        m_tile = fmaxf(m_tile, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      }
      m_tile *= sc;  // scale > 0
    } else {
#pragma unroll 1
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, fmaf(__uint_as_float(v[i]), sc, sbias[c + i]));
          m1 = fmaxf(m1, fmaf(__uint_as_float(v[i + 1]), sc, sbias[c + i + 1]));
          m2 = fmaxf(m2, fmaf(__uint_as_float(v[i + 2]), sc, sbias[c + i + 2]));
          m3 = fmaxf(m3, fmaf(__uint_as_float(v[i + 3]), sc, sbias[c + i + 3]));
        }
        m_tile = fmaxf(m_tile, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      }
    }
    // ---- lazy rescale of the running state (warp-uniform decision; tcgen05.ld/st are warp collectives)
    const bool need = m_tile > m_ref + AT_LAZY;
    if (__any_sync(0xffffffffu, need)) {
      const float m_new = need ? m_tile : m_ref;
      const float f = (j == 0) ? 0.f : ex2_approx(m_ref - m_new);  // 1 for rows that keep their reference
      l_run *= f;
      if (j > 0) {
#pragma unroll
        for (int c = 0; c < AT_D; c += 32) {
          uint32_t v[32];
          tmem_ld32(to + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st32(to + c, v);
        }
        tmem_st_wait();
      }
      m_ref = m_new;
    }
    // ---- pass 2: probabilities -> bf16 (hi/lo) -> swizzled smem
#pragma unroll 1
    for (int c = 0; c < ncols; c += 32) {
      uint32_t v[32];
      tmem_ld32(ts + c, v);
      tmem_ld_wait();
      float pr[32];
      if (!tail) {
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;  // independent partial sums (ILP)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -m_ref));
          pr[i + 1] = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -m_ref));
          pr[i + 2] = ex2_approx(fmaf(__uint_as_float(v[i + 2]), sc, -m_ref));
          pr[i + 3] = ex2_approx(fmaf(__uint_as_float(v[i + 3]), sc, -m_ref));
          l0 += pr[i]; l1 += pr[i + 1]; l2 += pr[i + 2]; l3 += pr[i + 3];
        }
        l_run += (l0 + l1) + (l2 + l3);
      } else {
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sc, sbias[c + i]) - m_ref);       // ex2(-inf) = 0 for masked keys
          pr[i + 1] = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, sbias[c + i + 1]) - m_ref);
          pr[i + 2] = ex2_approx(fmaf(__uint_as_float(v[i + 2]), sc, sbias[c + i + 2]) - m_ref);
          pr[i + 3] = ex2_approx(fmaf(__uint_as_float(v[i + 3]), sc, sbias[c + i + 3]) - m_ref);
          l0 += pr[i]; l1 += pr[i + 1]; l2 += pr[i + 2]; l3 += pr[i + 3];
        }
        l_run += (l0 + l1) + (l2 + l3);
      }
      uint8_t* prow = prow_base + (c >> 6) * AT_CHUNK;
      const int u0 = (c & 63) >> 3;  // first 16-byte unit inside the 128-byte row
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 w;
        w.x = pack_bf16(pr[8 * u + 0], pr[8 * u + 1]);
        w.y = pack_bf16(pr[8 * u + 2], pr[8 * u + 3]);
        w.z = pack_bf16(pr[8 * u + 4], pr[8 * u + 5]);
        w.w = pack_bf16(pr[8 * u + 6], pr[8 * u + 7]);
        *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ rsw) << 4)) = w;
        if (NSPLIT == 2) {
          float lo[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) lo[q] = pr[8 * u + q] - __bfloat162float(__float2bfloat16_rn(pr[8 * u + q]));
          uint4 wl;
          wl.x = pack_bf16(lo[0], lo[1]); wl.y = pack_bf16(lo[2], lo[3]);
          wl.z = pack_bf16(lo[4], lo[5]); wl.w = pack_bf16(lo[6], lo[7]);
          *reinterpret_cast<uint4*>(prow + 2 * AT_CHUNK + (((u0 + u) ^ rsw) << 4)) = wl;
        }
      }
    }
    // P (generic-proxy writes) must be visible to the tensor core (async proxy); all S reads / O rescales are done.
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_pv(j, j > 0 ? 1u : 0u, ncols / 16);
      if (j + 1 < n_tiles) {
        mbar_wait(&bar_kv[(j + 1) % AT_NBUF], ((j + 1) / AT_NBUF) & 1);
        tc_fence_after();
        issue_qk(j + 1);       // commits bar_s after P V_j and Q K_{j+1}^T
      } else {
        umma_commit(bar_o);
      }
    }
  }

  // ---- finalize: O / l -> bf16 (hi/lo)
  __syncwarp();
  mbar_wait(bar_o, 0);
  tc_fence_after();
  const int q = q0 + r;
  const float inv = 1.0f / l_run;
  __nv_bfloat16* op = p.out + (static_cast<long long>(b) * p.Lq + q) * p.ld_o + head * AT_D;
#pragma unroll
  for (int c = 0; c < AT_D; c += 32) {
    uint32_t v[32];
    tmem_ld32(to + c, v);
    tmem_ld_wait();
    if (q < p.Lq) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float y[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) y[t] = __uint_as_float(v[i + t]) * inv;
        uint4 w;
        w.x = pack_bf16(y[0], y[1]); w.y = pack_bf16(y[2], y[3]);
        w.z = pack_bf16(y[4], y[5]); w.w = pack_bf16(y[6], y[7]);
        *reinterpret_cast<uint4*>(op + c + i) = w;
        if (p.split_off > 0) {
          float lo[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) lo[t] = y[t] - __bfloat162float(__float2bfloat16_rn(y[t]));
          uint4 wl;
          wl.x = pack_bf16(lo[0], lo[1]); wl.y = pack_bf16(lo[2], lo[3]);
          wl.z = pack_bf16(lo[4], lo[5]); wl.w = pack_bf16(lo[6], lo[7]);
          *reinterpret_cast<uint4*>(op + p.split_off + c + i) = wl;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =====================================================================================================================
// attention_sub_kernel — same data layout and occupancy as attention_tc_kernel (thread = query row, two CTAs per SM in
// perf mode), but warp-specialised and software-pipelined at 64-key SUB-TILE granularity inside the CTA:
//   * warps 0-3 (128 threads) only do softmax; warp 4 (one elected lane) owns TMA loads and every tcgen05.mma. The
//     main loop has no CTA-wide barrier: softmax -> issuer through bar_p (128 arrivals: "P_g is in smem and S_g has
//     been read"), issuer -> softmax through tcgen05.commit barriers;
//   * S is double-buffered in TMEM (columns [0,64) and [64,128) = the two halves of a 128-key tile); Q K^T of sub-tile
//     g+2 is issued as soon as sub-tile g has been handed over, so the softmax warps find S_{g+1} ready, and P V of
//     sub-tile g runs underneath the softmax of g+1;
//   * single pass over S: the exponent uses the running reference maximum (the lazy-rescale reference, at most 2^8
//     below the true maximum), the sub-tile maximum is tracked on the fly and only when some row of the warp grew by
//     more than 2^8 the warp rescales O / l and recomputes the sub-tile (first sub-tile: explicit maximum pass);
//   * K and V tiles (128 keys) are loaded separately: a K buffer is free after Q K^T of both its halves, a V buffer
//     after P V of both its halves, which keeps every TMA load about two sub-tiles ahead of its first use.
// In-order tensor pipe => the commit that signals S_{g+2} also proves P V_g complete (P half and, for odd g, the V
// buffer reusable); bar_pv is only waited on by the rare rescale path and at the end.
constexpr int AS_SOFTMAX_THREADS = 128;
constexpr int AS_THREADS = AS_SOFTMAX_THREADS + 32;

__device__ __forceinline__ void softmax_bar_sync() {   // named barrier 1: the 128 softmax threads only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

template <int NSPLIT>
__global__ void __launch_bounds__(AS_THREADS, (NSPLIT == 1) ? 2 : 1)
attention_sub_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                     const __grid_constant__ CUtensorMap vmap, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NSPLIT>;
  constexpr int SUB = 64;                          // keys per sub-tile
  constexpr int HALF_BYTES = SUB * 128;            // 64 K rows of a swizzled [128][64] chunk
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;                 // [NBUF][KV_BYTES]
  uint8_t* sV = sK + AT_NBUF * Cfg::KV_BYTES;      // [NBUF][KV_BYTES]
  uint8_t* sP = sV + AT_NBUF * Cfg::KV_BYTES;      // hi chunks 0,1 (= sub-tile halves) then lo chunks 0,1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* bar_q = bars;            // [1]
  uint64_t* bar_k = bars + 1;        // [2] K tile landed
  uint64_t* bar_v = bars + 3;        // [2] V tile landed
  uint64_t* bar_s = bars + 5;        // [2] S half ready
  uint64_t* bar_pv = bars + 7;       // [2] P V of a half drained
  uint64_t* bar_p = bars + 9;        // [2] P half written + S half read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* sbias = reinterpret_cast<float*>(sP + Cfg::P_BYTES + 128);  // per-sub-tile key bias (log2 domain), -inf = masked

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.Lk + AT_BN - 1) / AT_BN;
  const int n_sub = (p.Lk + SUB - 1) / SUB;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[tng] attention: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&qmap);
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&vmap);
    for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_p[0], AS_SOFTMAX_THREADS);
    mbar_init(&bar_p[1], AS_SOFTMAX_THREADS);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base;
  const uint32_t tm_o = tmem_base + 128;

  if (warp == 4) {
    // ================================================================= loader / MMA issuer (one elected lane)
    if (lane == 0) {
      auto load_k = [&](int tile) {
        const int buf = tile % AT_NBUF;
        mbar_arrive_expect_tx(&bar_k[buf], Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s)
          tma_load_3d(sK + buf * Cfg::KV_BYTES + s * AT_CHUNK, &kmap, &bar_k[buf],
                      p.k_col0 + s * p.k_lo_off + head * AT_D, tile * AT_BN, b);
      };
      auto load_v = [&](int tile) {
        const int buf = tile % AT_NBUF;
        mbar_arrive_expect_tx(&bar_v[buf], Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s)
          tma_load_3d(sV + buf * Cfg::KV_BYTES + s * AT_CHUNK, &vmap, &bar_v[buf],
                      p.v_col0 + s * p.v_lo_off + head * AT_D, tile * AT_BN, b);
      };
      // S[g & 1] = Q K_g^T over the 64 keys of sub-tile g (hi*hi [+ lo*hi + hi*lo])
      auto issue_qk = [&](int g) {
        const int buf = (g >> 1) % AT_NBUF, h = g & 1;
        constexpr uint32_t idesc = umma_idesc_bf16(AT_BM, SUB, 0, 0);
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + buf * Cfg::KV_BYTES) + h * HALF_BYTES;
        constexpr int NT = (NSPLIT == 1) ? 1 : 3;
        const int qsel[3] = {0, 1, 0}, ksel[3] = {0, 0, 1};
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint64_t adesc = umma_desc_sw128(qa + qsel[t] * AT_CHUNK, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(ka + ksel[t] * AT_CHUNK, 16, 1024);
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) {
            umma_bf16(tm_s + h * SUB, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&bar_s[h]);
      };
      // O (+)= P_g V_g   (A = P chunk g&1, K-major; B = V rows [64 h, 64 h + 16 nk16), MN-major: 16 keys = 2048 B)
      auto issue_pv = [&](int g, uint32_t accumulate, int nk16) {
        const int buf = (g >> 1) % AT_NBUF, h = g & 1;
        constexpr uint32_t idesc = umma_idesc_bf16(AT_BM, AT_D, 0, 1);
        const uint32_t pa = smem_u32(sP) + h * AT_CHUNK, va = smem_u32(sV + buf * Cfg::KV_BYTES) + h * HALF_BYTES;
        constexpr int NT = (NSPLIT == 1) ? 1 : 3;
        const int psel[3] = {0, 1, 0}, vsel[3] = {0, 0, 1};
        uint32_t acc = accumulate;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
          for (int k = 0; k < SUB / 16; ++k) {
            if (k >= nk16) break;  // keys beyond the written P columns (short last sub-tile)
            const uint64_t adesc = umma_desc_sw128(pa + psel[t] * 2 * AT_CHUNK, 16, 1024) + 2 * k;
            const uint64_t bdesc = umma_desc_sw128(va + vsel[t] * AT_CHUNK + k * 2048, 1024, 1024);
            umma_bf16(tm_o, adesc, bdesc, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&bar_pv[h]);
      };

      mbar_arrive_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
      for (int s = 0; s < NSPLIT; ++s)
        tma_load_3d(sQ + s * AT_CHUNK, &qmap, bar_q, p.q_col0 + s * p.q_lo_off + head * AT_D, q0, b);
      for (int t = 0; t < AT_NBUF && t < n_tiles; ++t) { load_k(t); load_v(t); }
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
      issue_qk(0);
      if (n_sub > 1) issue_qk(1);
      for (int g = 0; g < n_sub; ++g) {
        const int h = g & 1, T = g >> 1;
        const int kv0 = g * SUB;
        const bool tail = (p.kbias != nullptr) || (kv0 + SUB > p.Lk);
        const int ncols = tail ? min(SUB, ((p.Lk - kv0) + 31) & ~31) : SUB;   // same rule as the softmax warps
        mbar_wait(&bar_p[h], T & 1);     // P_g written (and fenced to the async proxy), S_g fully read
        mbar_wait(&bar_v[T % AT_NBUF], (T / AT_NBUF) & 1);
        tc_fence_after();
        issue_pv(g, g > 0 ? 1u : 0u, ncols / 16);
        if (g + 2 < n_sub) {
          const int T2 = (g + 2) >> 1;
          mbar_wait(&bar_k[T2 % AT_NBUF], (T2 / AT_NBUF) & 1);
          tc_fence_after();
          issue_qk(g + 2);       // commits bar_s[h] after P V_g and Q K_{g+2}^T
        }
        if (h == 1) {
          // softmax of g = 2T+1 is over => S_g was complete => both Q K^T of K tile T are done, and (commit order)
          // P V of both halves of tile T-1 had drained before S_g was signalled
          if (T + 2 < n_tiles) load_k(T + 2);
          if (T >= 1 && T + 1 < n_tiles) load_v(T + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================= softmax warps: thread = query row = TMEM lane
    const int r = tid;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t ts = tm_s + lane_addr;
    const uint32_t to = tm_o + lane_addr;
    float m_ref = -INFINITY;  // reference max used in the exponent (log2 domain)
    float l_run = 0.f;
    const float sc = p.scale_log2e;
    const float* kb = p.kbias ? p.kbias + static_cast<long long>(b) * p.Lk : nullptr;
    constexpr float LOG2E = 1.4426950408889634f;
    uint8_t* prow_base = sP + r * 128;
    const int rsw = r & 7;

    for (int g = 0; g < n_sub; ++g) {
      const int h = g & 1, T = g >> 1;
      __syncwarp();
      mbar_wait(&bar_s[h], T & 1);   // S_g complete; in-order tensor pipe => P V_{g-2} complete as well (P half free)
      tc_fence_after();
      const int kv0 = g * SUB;
      const bool tail = (kb != nullptr) || (kv0 + SUB > p.Lk);
      // columns actually processed in this sub-tile (multiple of 32); P columns beyond are never written nor multiplied
      const int ncols = tail ? min(SUB, ((p.Lk - kv0) + 31) & ~31) : SUB;
      if (tail) {
        float bv = -INFINITY;
        if (tid < SUB) {
          const int kv = kv0 + tid;
          if (kv < p.Lk) bv = kb ? kb[kv] * LOG2E : 0.f;
        }
        softmax_bar_sync();  // previous sub-tile's readers are done
        if (tid < SUB) sbias[tid] = bv;
        softmax_bar_sync();
      }
      const uint32_t tsh = ts + h * SUB;
      if (g == 0) {
        // explicit maximum pass for the very first sub-tile (no reference yet)
        float m_tile = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < ncols; c += 32) {
          uint32_t v[32];
          tmem_ld32(tsh + c, v);
          tmem_ld_wait();
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          if (!tail) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
            }
            m_tile = fmaxf(m_tile, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sc);  // scale > 0
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(m0, fmaf(__uint_as_float(v[i]), sc, sbias[c + i]));
              m1 = fmaxf(m1, fmaf(__uint_as_float(v[i + 1]), sc, sbias[c + i + 1]));
              m2 = fmaxf(m2, fmaf(__uint_as_float(v[i + 2]), sc, sbias[c + i + 2]));
              m3 = fmaxf(m3, fmaf(__uint_as_float(v[i + 3]), sc, sbias[c + i + 3]));
            }
            m_tile = fmaxf(m_tile, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          }
        }
        m_ref = m_tile;
      }
      // ---- probabilities relative to m_ref -> bf16 (hi/lo) -> swizzled smem; at most two rounds (see header)
      float lsum = 0.f;
#pragma unroll 1
      for (int round = 0; round < 2; ++round) {
        float xmax = -INFINITY;
        lsum = 0.f;
        uint8_t* prow = prow_base + h * AT_CHUNK;
#pragma unroll 1
        for (int c = 0; c < ncols; c += 32) {
          uint32_t v[32];
          tmem_ld32(tsh + c, v);
          tmem_ld_wait();
          float pr[32];
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;                                  // independent partial sums (ILP)
          float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;          // and partial maxima
          if (!tail) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float a0 = fmaf(__uint_as_float(v[i]), sc, -m_ref), a1 = fmaf(__uint_as_float(v[i + 1]), sc, -m_ref);
              const float a2 = fmaf(__uint_as_float(v[i + 2]), sc, -m_ref), a3 = fmaf(__uint_as_float(v[i + 3]), sc, -m_ref);
              x0 = fmaxf(x0, a0); x1 = fmaxf(x1, a1); x2 = fmaxf(x2, a2); x3 = fmaxf(x3, a3);
              pr[i] = ex2_approx(a0); pr[i + 1] = ex2_approx(a1); pr[i + 2] = ex2_approx(a2); pr[i + 3] = ex2_approx(a3);
              l0 += pr[i]; l1 += pr[i + 1]; l2 += pr[i + 2]; l3 += pr[i + 3];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {   // ex2(-inf) = 0 for masked keys
              const float a0 = fmaf(__uint_as_float(v[i]), sc, sbias[c + i]) - m_ref;
              const float a1 = fmaf(__uint_as_float(v[i + 1]), sc, sbias[c + i + 1]) - m_ref;
              const float a2 = fmaf(__uint_as_float(v[i + 2]), sc, sbias[c + i + 2]) - m_ref;
              const float a3 = fmaf(__uint_as_float(v[i + 3]), sc, sbias[c + i + 3]) - m_ref;
              x0 = fmaxf(x0, a0); x1 = fmaxf(x1, a1); x2 = fmaxf(x2, a2); x3 = fmaxf(x3, a3);
              pr[i] = ex2_approx(a0); pr[i + 1] = ex2_approx(a1); pr[i + 2] = ex2_approx(a2); pr[i + 3] = ex2_approx(a3);
              l0 += pr[i]; l1 += pr[i + 1]; l2 += pr[i + 2]; l3 += pr[i + 3];
            }
          }
          lsum += (l0 + l1) + (l2 + l3);
          xmax = fmaxf(xmax, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)));
          const int u0 = c >> 3;  // first 16-byte unit inside the 128-byte row
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack_bf16(pr[8 * u + 0], pr[8 * u + 1]);
            w.y = pack_bf16(pr[8 * u + 2], pr[8 * u + 3]);
            w.z = pack_bf16(pr[8 * u + 4], pr[8 * u + 5]);
            w.w = pack_bf16(pr[8 * u + 6], pr[8 * u + 7]);
            *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ rsw) << 4)) = w;
            if (NSPLIT == 2) {
              float lo[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) lo[q] = pr[8 * u + q] - __bfloat162float(__float2bfloat16_rn(pr[8 * u + q]));
              uint4 wl;
              wl.x = pack_bf16(lo[0], lo[1]); wl.y = pack_bf16(lo[2], lo[3]);
              wl.z = pack_bf16(lo[4], lo[5]); wl.w = pack_bf16(lo[6], lo[7]);
              *reinterpret_cast<uint4*>(prow + 2 * AT_CHUNK + (((u0 + u) ^ rsw) << 4)) = wl;
            }
          }
        }
        // ---- lazy rescale (warp-uniform decision; tcgen05.ld/st are warp collectives)
        const bool need = xmax > AT_LAZY;
        if (round == 1 || !__any_sync(0xffffffffu, need)) break;
        const float f = need ? ex2_approx(-xmax) : 1.0f;   // = 2^(m_ref - m_new), m_new = m_ref + xmax
        l_run *= f;
        if (g > 0) {
          mbar_wait(&bar_pv[(g - 1) & 1], ((g - 1) >> 1) & 1);   // every P V issued so far has landed in O
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < AT_D; c += 32) {
            uint32_t v[32];
            tmem_ld32(to + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st32(to + c, v);
          }
          tmem_st_wait();
        }
        if (need) m_ref += xmax;
      }
      l_run += lsum;
      // hand-over: P (generic-proxy writes) visible to the tensor core (async proxy); S reads / O rescales are done
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_p[h]);
    }

    // ---- finalize: O / l -> bf16 (hi/lo)
    __syncwarp();
    mbar_wait(&bar_pv[(n_sub - 1) & 1], ((n_sub - 1) >> 1) & 1);
    tc_fence_after();
    const int q = q0 + r;
    const float inv = 1.0f / l_run;
    __nv_bfloat16* op = p.out + (static_cast<long long>(b) * p.Lq + q) * p.ld_o + head * AT_D;
#pragma unroll
    for (int c = 0; c < AT_D; c += 32) {
      uint32_t v[32];
      tmem_ld32(to + c, v);
      tmem_ld_wait();
      if (q < p.Lq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float y[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) y[t] = __uint_as_float(v[i + t]) * inv;
          uint4 w;
          w.x = pack_bf16(y[0], y[1]); w.y = pack_bf16(y[2], y[3]);
          w.z = pack_bf16(y[4], y[5]); w.w = pack_bf16(y[6], y[7]);
          *reinterpret_cast<uint4*>(op + c + i) = w;
          if (p.split_off > 0) {
            float lo[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) lo[t] = y[t] - __bfloat162float(__float2bfloat16_rn(y[t]));
            uint4 wl;
            wl.x = pack_bf16(lo[0], lo[1]); wl.y = pack_bf16(lo[2], lo[3]);
            wl.z = pack_bf16(lo[4], lo[5]); wl.w = pack_bf16(lo[6], lo[7]);
            *reinterpret_cast<uint4*>(op + p.split_off + c + i) = wl;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =====================================================================================================================
// attention_ws_kernel — warp-specialised variant (perf mode, NSPLIT = 1): 8 softmax warps + 1 loader/issuer warp, one
// CTA per SM. Two threads share a query row (warps w and w+4 own the two 64-key halves of the 128-key tile), S is
// double-buffered in TMEM and the issuer keeps the tensor pipe one tile ahead:
//     issue order:  ... P V_{j-1}, Q K_{j+1}^T, P V_j, Q K_{j+2}^T ...   (each pair as soon as P_j is in smem)
// so Q K^T of the next tile and P V of the previous one execute while the softmax warps work on the current tile.
// No __syncthreads in the main loop: softmax -> issuer through an mbarrier with 256 arrivals, issuer -> softmax through
// tcgen05.commit barriers; the two threads of a row exchange their partial row max through smem + a 64-thread named
// barrier.
constexpr int AW_NBUF = 3;
constexpr int AW_SOFTMAX_THREADS = 256;
constexpr int AW_THREADS = AW_SOFTMAX_THREADS + 32;
constexpr int AW_SMEM_BYTES = AT_CHUNK /*Q*/ + AW_NBUF * 2 * AT_CHUNK /*K,V*/ + 2 * AT_CHUNK /*P*/ + 128 /*barriers*/ +
                              AT_BN * 4 /*bias*/ + 2 * 2 * AT_BM * 4 /*row exchange*/;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(AW_THREADS, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                    const __grid_constant__ CUtensorMap vmap, const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_CHUNK;               // [NBUF][16 KB]
  uint8_t* sV = sK + AW_NBUF * AT_CHUNK;     // [NBUF][16 KB]
  uint8_t* sP = sV + AW_NBUF * AT_CHUNK;     // 2 chunks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_CHUNK);
  uint64_t* bar_q = bars;                   // [1]
  uint64_t* bar_kv = bars + 1;              // [NBUF]
  uint64_t* bar_s = bars + 1 + AW_NBUF;     // [2]  S buffer ready
  uint64_t* bar_pv = bars + 3 + AW_NBUF;    // [1]  P V_j drained (P smem / O / K-V buffer reusable)
  uint64_t* bar_p = bars + 4 + AW_NBUF;     // [1]  P_j written + S_j consumed (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + AW_NBUF);
  float* sbias = reinterpret_cast<float*>(sP + 2 * AT_CHUNK + 128);
  float* xch = sbias + AT_BN;               // [2 (tile parity)][2 (half)][128 rows]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.Lk + AT_BN - 1) / AT_BN;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[tng] attention_ws: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(bar_q, 1);
    for (int i = 0; i < AW_NBUF; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(bar_pv, 1);
    mbar_init(bar_p, AW_SOFTMAX_THREADS);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_o = tmem_base + 256;

  if (warp == 8) {
    // ================================================================= loader + MMA issuer (one thread)
    if (lane == 0) {
      tma_prefetch_desc(&qmap);
      tma_prefetch_desc(&kmap);
      tma_prefetch_desc(&vmap);
      auto load_kv = [&](int tile) {
        const int buf = tile % AW_NBUF;
        mbar_arrive_expect_tx(&bar_kv[buf], 2 * AT_CHUNK);
        tma_load_3d(sK + buf * AT_CHUNK, &kmap, &bar_kv[buf], p.k_col0 + head * AT_D, tile * AT_BN, b);
        tma_load_3d(sV + buf * AT_CHUNK, &vmap, &bar_kv[buf], p.v_col0 + head * AT_D, tile * AT_BN, b);
      };
      auto issue_qk = [&](int tile) {
        constexpr uint32_t idesc = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);
        const uint64_t adesc = umma_desc_sw128(smem_u32(sQ), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sK + (tile % AW_NBUF) * AT_CHUNK), 16, 1024);
        const uint32_t d = tmem_base + (tile & 1) * 128;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit(&bar_s[tile & 1]);
      };
      auto issue_pv = [&](int tile, int nk16) {
        constexpr uint32_t idesc = umma_idesc_bf16(AT_BM, AT_D, 0, 1);
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + (tile % AW_NBUF) * AT_CHUNK);
        for (int k = 0; k < nk16; ++k) {
          const uint64_t adesc = umma_desc_sw128(pa + (k >> 2) * AT_CHUNK, 16, 1024) + 2 * (k & 3);
          const uint64_t bdesc = umma_desc_sw128(va + k * 2048, 1024, 1024);
          umma_bf16(tm_o, adesc, bdesc, idesc, (tile > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_pv);
      };
      mbar_arrive_expect_tx(bar_q, AT_CHUNK);
      tma_load_3d(sQ, &qmap, bar_q, p.q_col0 + head * AT_D, q0, b);
      for (int t = 0; t < AW_NBUF && t < n_tiles; ++t) load_kv(t);
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[0], 0);
      tc_fence_after();
      issue_qk(0);
      if (n_tiles > 1) {
        mbar_wait(&bar_kv[1], 0);
        tc_fence_after();
        issue_qk(1);
      }
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(bar_p, j & 1);  // P_j in smem, S_j fully read, O rescaled if needed
        tc_fence_after();
        const int kv0 = j * AT_BN;
        const bool tail = (p.kbias != nullptr) || (kv0 + AT_BN > p.Lk);
        const int ncols = tail ? min(AT_BN, ((p.Lk - kv0) + 31) & ~31) : AT_BN;
        issue_pv(j, ncols / 16);
        if (j + 2 < n_tiles) {
          mbar_wait(&bar_kv[(j + 2) % AW_NBUF], ((j + 2) / AW_NBUF) & 1);
          tc_fence_after();
          issue_qk(j + 2);       // into the S buffer tile j just released
        }
        if (j + AW_NBUF < n_tiles) {
          mbar_wait(bar_pv, j & 1);  // P V_j (and Q K_j^T long before) drained: K/V buffer j % NBUF is free
          load_kv(j + AW_NBUF);
        }
      }
    }
  } else {
    // ================================================================= softmax warps (two threads per query row)
    const int quarter = warp & 3;
    const int hf = warp >> 2;                 // which 64-key half of the tile this thread owns
    const int r = quarter * 32 + lane;        // query row == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t to = tm_o + lane_addr + 32 * hf;   // my 32 of the 64 O columns
    float m_ref = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2e;
    const float* kb = p.kbias ? p.kbias + static_cast<long long>(b) * p.Lk : nullptr;
    constexpr float LOG2E = 1.4426950408889634f;
    uint8_t* prow = sP + hf * AT_CHUNK + r * 128;
    const int rsw = r & 7;
    const int c0 = 64 * hf;

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = j * AT_BN;
      const bool tail = (kb != nullptr) || (kv0 + AT_BN > p.Lk);
      const int ncols = tail ? min(AT_BN, ((p.Lk - kv0) + 31) & ~31) : AT_BN;
      const int cend = min(c0 + 64, ncols);
      if (tail) {
        named_bar_sync(5, AW_SOFTMAX_THREADS);     // previous tile's readers of sbias are done
        if (tid < AT_BN) {
          const int kv = kv0 + tid;
          float bv = -INFINITY;
          if (kv < p.Lk) bv = kb ? kb[kv] * LOG2E : 0.f;
          sbias[tid] = bv;
        }
        named_bar_sync(5, AW_SOFTMAX_THREADS);
      }
      __syncwarp();
      mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_base + (j & 1) * 128 + lane_addr;
      // ---- pass 1: partial row maximum over my key half (log2 domain)
      float m_part = -INFINITY;
#pragma unroll 1
      for (int c = c0; c < cend; c += 32) {
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        if (!tail) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[i]));
            m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
          }
          m_part = fmaxf(m_part, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sc);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(m0, fmaf(__uint_as_float(v[i]), sc, sbias[c + i]));
            m1 = fmaxf(m1, fmaf(__uint_as_float(v[i + 1]), sc, sbias[c + i + 1]));
            m2 = fmaxf(m2, fmaf(__uint_as_float(v[i + 2]), sc, sbias[c + i + 2]));
            m3 = fmaxf(m3, fmaf(__uint_as_float(v[i + 3]), sc, sbias[c + i + 3]));
          }
          m_part = fmaxf(m_part, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        }
      }
      // ---- exchange with the thread that owns the other half of this row
      float* xr = xch + (j & 1) * (2 * AT_BM);
      xr[hf * AT_BM + r] = m_part;
      named_bar_sync(1 + quarter, 64);
      const float m_tile = fmaxf(m_part, xr[(hf ^ 1) * AT_BM + r]);
      // P smem and O are touched below: P V_{j-1} must have drained
      if (j > 0) {
        mbar_wait(bar_pv, (j - 1) & 1);
        tc_fence_after();
      }
      // ---- lazy rescale (identical decision in both threads of a row; warp-uniform execution of tcgen05.ld/st)
      const bool need = m_tile > m_ref + AT_LAZY;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? m_tile : m_ref;
        const float f = (j == 0) ? 0.f : ex2_approx(m_ref - m_new);
        l_run *= f;
        if (j > 0) {
          uint32_t v[32];
          tmem_ld32(to, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st32(to, v);
          tmem_st_wait();
        }
        m_ref = m_new;
      }
      // ---- pass 2: probabilities of my key half -> bf16 -> swizzled smem (A operand of P V)
#pragma unroll 1
      for (int c = c0; c < cend; c += 32) {
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_ld_wait();
        float pr[32];
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        if (!tail) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -m_ref));
            pr[i + 1] = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -m_ref));
            pr[i + 2] = ex2_approx(fmaf(__uint_as_float(v[i + 2]), sc, -m_ref));
            pr[i + 3] = ex2_approx(fmaf(__uint_as_float(v[i + 3]), sc, -m_ref));
            l0 += pr[i]; l1 += pr[i + 1]; l2 += pr[i + 2]; l3 += pr[i + 3];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sc, sbias[c + i]) - m_ref);
            pr[i + 1] = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, sbias[c + i + 1]) - m_ref);
            pr[i + 2] = ex2_approx(fmaf(__uint_as_float(v[i + 2]), sc, sbias[c + i + 2]) - m_ref);
            pr[i + 3] = ex2_approx(fmaf(__uint_as_float(v[i + 3]), sc, sbias[c + i + 3]) - m_ref);
            l0 += pr[i]; l1 += pr[i + 1]; l2 += pr[i + 2]; l3 += pr[i + 3];
          }
        }
        l_run += (l0 + l1) + (l2 + l3);
        const int u0 = ((c - c0) >> 3);  // first 16-byte unit inside my 128-byte P row (0 or 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_bf16(pr[8 * u + 0], pr[8 * u + 1]);
          w.y = pack_bf16(pr[8 * u + 2], pr[8 * u + 3]);
          w.z = pack_bf16(pr[8 * u + 4], pr[8 * u + 5]);
          w.w = pack_bf16(pr[8 * u + 6], pr[8 * u + 7]);
          *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ rsw) << 4)) = w;
        }
      }
      // P (generic-proxy writes) -> async proxy; S reads / O rescale complete -> tell the issuer
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }

    // ---- finalize: O / l -> bf16; the two threads of a row each write 32 of the 64 head columns
    mbar_wait(bar_pv, (n_tiles - 1) & 1);
    tc_fence_after();
    float* xr = xch + (n_tiles & 1) * (2 * AT_BM);
    xr[hf * AT_BM + r] = l_run;
    named_bar_sync(1 + quarter, 64);
    const float inv = 1.0f / (l_run + xr[(hf ^ 1) * AT_BM + r]);
    const int q = q0 + r;
    uint32_t v[32];
    tmem_ld32(to, v);
    tmem_ld_wait();
    if (q < p.Lq) {
      __nv_bfloat16* op = p.out + (static_cast<long long>(b) * p.Lq + q) * p.ld_o + head * AT_D + 32 * hf;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
        *reinterpret_cast<uint4*>(op + i) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_attn_ws(const tng_attn_desc* d, const CUtensorMap& qm, const CUtensorMap& km, const CUtensorMap& vm,
                          const AttnParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(attention_ws): %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((d->Lq + AT_BM - 1) / AT_BM, d->heads, d->batch);
  attention_ws_kernel<<<grid, AW_THREADS, AW_SMEM_BYTES, st>>>(qm, km, vm, p);
  count_launch();
  return check_launch("attention_ws");
}

template <int NSPLIT>
static int launch_attn_sub(const tng_attn_desc* d, const CUtensorMap& qm, const CUtensorMap& km, const CUtensorMap& vm,
                           const AttnParams& p, cudaStream_t st) {
  using Cfg = AttnCfg<NSPLIT>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_sub_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(attention_sub): %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((d->Lq + AT_BM - 1) / AT_BM, d->heads, d->batch);
  attention_sub_kernel<NSPLIT><<<grid, AS_THREADS, Cfg::SMEM_BYTES, st>>>(qm, km, vm, p);
  count_launch();
  return check_launch("attention_sub");
}

template <int NSPLIT>
static int launch_attn(const tng_attn_desc* d, const CUtensorMap& qm, const CUtensorMap& km, const CUtensorMap& vm,
                       const AttnParams& p, cudaStream_t st) {
  using Cfg = AttnCfg<NSPLIT>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((d->Lq + AT_BM - 1) / AT_BM, d->heads, d->batch);
  attention_tc_kernel<NSPLIT><<<grid, 128, Cfg::SMEM_BYTES, st>>>(qm, km, vm, p);
  count_launch();
  return check_launch("attention_tc");
}

}  // namespace tng

using namespace tng;

extern "C" int tng_attention(const tng_attn_desc* d, void* stream) {
  if (!d || !d->q || !d->k || !d->v || !d->out) return set_error(TNG_EINVAL, "attention: null argument");
  if (d->nsplit != 1 && d->nsplit != 2) return set_error(TNG_EINVAL, "attention: nsplit=%d", d->nsplit);
  if (d->batch <= 0 || d->heads <= 0 || d->Lq <= 0 || d->Lk <= 0) return set_error(TNG_EINVAL, "attention: bad sizes");
  if (d->scale <= 0.f) return set_error(TNG_EINVAL, "attention: scale must be positive");
  if (d->ld_o % 8 || d->split_off % 8 || (reinterpret_cast<uintptr_t>(d->out) & 15))
    return set_error(TNG_EINVAL, "attention: output must allow 16-byte stores");
  AttnParams p;
  p.Lq = d->Lq; p.Lk = d->Lk; p.heads = d->heads;
  p.q_col0 = d->q_col0; p.q_lo_off = d->q_lo_off;
  p.k_col0 = d->k_col0; p.k_lo_off = d->k_lo_off;
  p.v_col0 = d->v_col0; p.v_lo_off = d->v_lo_off;
  p.kbias = d->kbias;
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  p.ld_o = d->ld_o; p.split_off = d->split_off;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  CUtensorMap qm, km, vm;
  uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {(uint64_t)d->ld_q, (uint64_t)d->Lq, (uint64_t)d->batch};
    uint64_t str[2] = {(uint64_t)d->ld_q * 2, (uint64_t)d->ld_q * 2 * (uint64_t)d->Lq};
    int rc = encode_tmap_bf16(&qm, d->q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->ld_k, (uint64_t)d->Lk, (uint64_t)d->batch};
    uint64_t str[2] = {(uint64_t)d->ld_k * 2, (uint64_t)d->ld_k * 2 * (uint64_t)d->Lk};
    int rc = encode_tmap_bf16(&km, d->k, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->ld_v, (uint64_t)d->Lk, (uint64_t)d->batch};
    uint64_t str[2] = {(uint64_t)d->ld_v * 2, (uint64_t)d->ld_v * 2 * (uint64_t)d->Lk};
    int rc = encode_tmap_bf16(&vm, d->v, 3, dims, str, box);
    if (rc) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // variants (TNG_ATTN): 4 = sub-tile pipelined kernel (default), 2 = tile-at-a-time kernel, 3 = warp-specialised
  // one-CTA-per-SM kernel (perf mode only; measured slower)
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("TNG_ATTN"); variant = e ? atoi(e) : 4; }
  if (d->nsplit == 2) return variant == 2 ? launch_attn<2>(d, qm, km, vm, p, st) : launch_attn_sub<2>(d, qm, km, vm, p, st);
  if (variant == 3 && d->split_off == 0) return launch_attn_ws(d, qm, km, vm, p, st);
  if (variant == 2) return launch_attn<1>(d, qm, km, vm, p, st);
  return launch_attn_sub<1>(d, qm, km, vm, p, st);
}
