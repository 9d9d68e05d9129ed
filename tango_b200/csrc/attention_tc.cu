// attention_tc.cu — tcgen05 flash attention for head width 64 (UNet self- and cross-attention), sm_100a.
//
// One CTA = 128 query rows of one (batch, head); thread r of the softmax warps owns query row r = TMEM lane r; Q / K / V
// arrive as TMA SWIZZLE_128B tiles; S, P and O live in TMEM; K-major A/B operands straight from the TMA tiles, V consumed
// as an MN-major B operand (no transpose); two CTAs per SM in perf mode. attention_sub_kernel is warp-specialised
// (4 softmax warps + 1 loader / MMA-issuer warp) and software-pipelined over 64-key sub-tiles with S double-buffered in
// TMEM, single-pass softmax against a lazily advanced reference maximum ("lazy rescale": the reference max only moves when
// a row grew by more than 2^8), P handed to the P V MMA through tensor memory (tcgen05.st + TS-mode MMA), three K/V tile
// buffers. See the comment above the kernel.
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

// =====================================================================================================================
// attention_sub_kernel — thread = query row, two CTAs per SM in perf mode, warp-specialised and software-pipelined at
// 64-key SUB-TILE granularity inside the CTA:
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
template <int NSPLIT, int VAR>
struct AttnSubCfg {
  static constexpr bool P_TMEM = (VAR >= 6);
  static constexpr int NB = P_TMEM ? 3 : 2;
  static constexpr int P_SMEM = P_TMEM ? 0 : AttnCfg<NSPLIT>::P_BYTES;
  static constexpr int SMEM_BYTES = AttnCfg<NSPLIT>::Q_BYTES + NB * 2 * AttnCfg<NSPLIT>::KV_BYTES + P_SMEM + 128 + AT_BN * 4;
};
constexpr int AS_SOFTMAX_THREADS = 128;
constexpr int AS_THREADS = AS_SOFTMAX_THREADS + 32;

__device__ __forceinline__ void softmax_bar_sync() {   // named barrier 1: the 128 softmax threads only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

template <int NSPLIT, int VAR>
__global__ void __launch_bounds__(AS_THREADS, (NSPLIT == 1) ? 2 : 1)
attention_sub_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                     const __grid_constant__ CUtensorMap vmap, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NSPLIT>;
  constexpr int SUB = 64;                          // keys per sub-tile
  constexpr int HALF_BYTES = SUB * 128;            // 64 K rows of a swizzled [128][64] chunk
  constexpr bool P_TMEM = (VAR >= 6);              // P handed to the P V MMA through tensor memory instead of smem
  // TMEM columns: S halves [0,128), O [128,192), P (bf16 pairs) hi half 0/1 [192,256), lo half 0/1 [256,320)
  constexpr int TMEM_COLS_SUB = (NSPLIT == 1) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NB = AttnSubCfg<NSPLIT, VAR>::NB;  // K / V tile buffers (3 when P lives in TMEM: its smem pays for them)
  constexpr int P_SMEM = AttnSubCfg<NSPLIT, VAR>::P_SMEM;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;                 // [NB][KV_BYTES]
  uint8_t* sV = sK + NB * Cfg::KV_BYTES;           // [NB][KV_BYTES]
  uint8_t* sP = sV + NB * Cfg::KV_BYTES;           // hi chunks 0,1 (= sub-tile halves) then lo chunks 0,1 (P_SMEM only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_SMEM);
  uint64_t* bar_q = bars;                    // [1]
  uint64_t* bar_k = bars + 1;                // [NB] K tile landed
  uint64_t* bar_v = bars + 1 + NB;           // [NB] V tile landed
  uint64_t* bar_s = bars + 1 + 2 * NB;       // [2] S half ready
  uint64_t* bar_pv = bars + 3 + 2 * NB;      // [2] P V of a half drained
  uint64_t* bar_p = bars + 5 + 2 * NB;       // [2] P half written + S half read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7 + 2 * NB);
  float* sbias = reinterpret_cast<float*>(sP + P_SMEM + 128);  // per-sub-tile key bias (log2 domain), -inf = masked

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
    for (int i = 0; i < 5 + 2 * NB; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_p[0], AS_SOFTMAX_THREADS);
    mbar_init(&bar_p[1], AS_SOFTMAX_THREADS);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS_SUB);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base;
  const uint32_t tm_o = tmem_base + 128;
  const uint32_t tm_p = tmem_base + 192;

  if (warp == 4) {
    // ================================================================= loader / MMA issuer (one elected lane)
    if (lane == 0) {
      auto load_k = [&](int tile) {
        const int buf = tile % NB;
        mbar_arrive_expect_tx(&bar_k[buf], Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s)
          tma_load_3d(sK + buf * Cfg::KV_BYTES + s * AT_CHUNK, &kmap, &bar_k[buf],
                      p.k_col0 + s * p.k_lo_off + head * AT_D, tile * AT_BN, b);
      };
      auto load_v = [&](int tile) {
        const int buf = tile % NB;
        mbar_arrive_expect_tx(&bar_v[buf], Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s)
          tma_load_3d(sV + buf * Cfg::KV_BYTES + s * AT_CHUNK, &vmap, &bar_v[buf],
                      p.v_col0 + s * p.v_lo_off + head * AT_D, tile * AT_BN, b);
      };
      // Descriptors are built once: the issuer is a single serial instruction stream on the critical path of every
      // hand-over, so per-issue work is reduced to a few adds (smem offsets enter the 16-byte address field directly).
      constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BM, SUB, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_D, 0, 1);
      const uint64_t qdesc0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV), 1024, 1024);
      const uint64_t pdesc0 = umma_desc_sw128(smem_u32(sP), 16, 1024);
      constexpr int NT = (NSPLIT == 1) ? 1 : 3;
      // S[g & 1] = Q K_g^T over the 64 keys of sub-tile g (hi*hi [+ lo*hi + hi*lo])
      auto issue_qk = [&](int g) {
        const int buf = (g >> 1) % NB, h = g & 1;
        const uint64_t kd = kdesc0 + static_cast<uint64_t>((buf * Cfg::KV_BYTES + h * HALF_BYTES) >> 4);
        const int qsel[3] = {0, 1, 0}, ksel[3] = {0, 0, 1};
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) {
            umma_bf16(tm_s + h * SUB, qdesc0 + static_cast<uint64_t>((qsel[t] * AT_CHUNK) >> 4) + 2 * k,
                      kd + static_cast<uint64_t>((ksel[t] * AT_CHUNK) >> 4) + 2 * k, idesc_qk, acc);
            acc = 1;
          }
        }
        umma_commit(&bar_s[h]);
      };
      // O (+)= P_g V_g   (A = P chunk g&1, K-major; B = V rows [64 h, 64 h + 16 nk16), MN-major: 16 keys = 2048 B)
      auto issue_pv = [&](int g, uint32_t accumulate, int nk16) {
        const int buf = (g >> 1) % NB, h = g & 1;
        const uint64_t vd = vdesc0 + static_cast<uint64_t>((buf * Cfg::KV_BYTES + h * HALF_BYTES) >> 4);
        const uint64_t pd = pdesc0 + static_cast<uint64_t>((h * AT_CHUNK) >> 4);
        const int psel[3] = {0, 1, 0}, vsel[3] = {0, 0, 1};
        uint32_t acc = accumulate;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
          for (int k = 0; k < SUB / 16; ++k) {
            if (k >= nk16) break;  // keys beyond the written P columns (short last sub-tile)
            const uint64_t bdesc = vd + static_cast<uint64_t>((vsel[t] * AT_CHUNK + k * 2048) >> 4);
            if (P_TMEM) umma_bf16_ts(tm_o, tm_p + psel[t] * 64 + h * 32 + 8 * k, bdesc, idesc_pv, acc);
            else umma_bf16(tm_o, pd + static_cast<uint64_t>((psel[t] * 2 * AT_CHUNK) >> 4) + 2 * k, bdesc, idesc_pv, acc);
            acc = 1;
          }
        }
        umma_commit(&bar_pv[h]);
      };

      mbar_arrive_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
      for (int s = 0; s < NSPLIT; ++s)
        tma_load_3d(sQ + s * AT_CHUNK, &qmap, bar_q, p.q_col0 + s * p.q_lo_off + head * AT_D, q0, b);
      for (int t = 0; t < NB && t < n_tiles; ++t) { load_k(t); load_v(t); }
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
        if (h == 0) mbar_wait(&bar_v[T % NB], (T / NB) & 1);    // first use of V tile T (h == 1 reuses it)
        mbar_wait(&bar_p[h], T & 1);     // P_g written (and fenced / stored), S_g fully read
        tc_fence_after();
        issue_pv(g, g > 0 ? 1u : 0u, ncols / 16);
        if (g + 2 < n_sub) {
          if (h == 0) {                  // first use of K tile T + 1
            mbar_wait(&bar_k[(T + 1) % NB], ((T + 1) / NB) & 1);
            tc_fence_after();
          }
          issue_qk(g + 2);       // commits bar_s[h] after P V_g and Q K_{g+2}^T
        }
        if (h == 1) {
          // softmax of g = 2T+1 is over => S_g was complete => both Q K^T of K tile T are done, and (commit order)
          // P V of both halves of tile T-1 had drained before S_g was signalled
          if (T + NB < n_tiles) load_k(T + NB);
          if (T >= 1 && T - 1 + NB < n_tiles) load_v(T - 1 + NB);
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

    // S of sub-tile g is requested (tcgen05.ld of both 32-column chunks) while P of sub-tile g-1 is still being stored
    // and handed over, so the TMEM round trips of consecutive sub-tiles overlap. The loads are waited for before the
    // loop back-edge: the registers must hold the data before the compiler may move them.
    uint32_t va[32], vb[32];
    auto request_s = [&](int g) {
      const int h = g & 1;
      __syncwarp();
      mbar_wait(&bar_s[h], (g >> 1) & 1);   // S_g complete; in-order tensor pipe => P V_{g-2} complete (P half free)
      tc_fence_after();
      const int left = p.Lk - g * SUB;      // > 0
      tmem_ld32(ts + h * SUB, va);
      if (left > 32) tmem_ld32(ts + h * SUB + 32, vb);
    };
    request_s(0);
    tmem_ld_wait();
    for (int g = 0; g < n_sub; ++g) {
      const int h = g & 1;
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
        // explicit maximum for the very first sub-tile (no reference yet), from the registers already loaded
        float m_tile = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (cc == 1 && ncols <= 32) break;
          const int c = 32 * cc;
          const uint32_t* v = cc ? vb : va;
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
        if (round == 1) {   // the recompute after a rescale reads S_g again
          tmem_ld32(tsh, va);
          if (ncols > 32) tmem_ld32(tsh + 32, vb);
          tmem_ld_wait();
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (cc == 1 && ncols <= 32) break;
          const int c = 32 * cc;
          const uint32_t* v = cc ? vb : va;
          float pr[32];
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;                                  // independent partial sums (ILP)
          float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;          // and partial maxima
          if (!tail) {
            // packed fp32x2 FMA / ADD and 3-input max: ~3 issue slots per element instead of ~4.6
            const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_ref, -m_ref);
            float2 l01 = make_float2(0.f, 0.f), l23 = l01;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float2 a01 = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
              const float2 a23 = ffma2(make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, nm2);
              x0 = fmaxf(x0, fmaxf(a01.x, a01.y));
              x1 = fmaxf(x1, fmaxf(a23.x, a23.y));
              pr[i] = ex2_approx(a01.x); pr[i + 1] = ex2_approx(a01.y);
              pr[i + 2] = ex2_approx(a23.x); pr[i + 3] = ex2_approx(a23.y);
              l01 = fadd2(l01, make_float2(pr[i], pr[i + 1]));
              l23 = fadd2(l23, make_float2(pr[i + 2], pr[i + 3]));
            }
            l0 = l01.x; l1 = l01.y; l2 = l23.x; l3 = l23.y;
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
          if (P_TMEM) {
            // 32 probabilities -> 16 bf16 pairs -> 16 TMEM columns of this row (A operand of the P V MMA)
            uint32_t pk[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) pk[u] = pack_bf16(pr[2 * u], pr[2 * u + 1]);
            tmem_st16(tm_p + lane_addr + h * 32 + 16 * cc, pk);
            if (NSPLIT == 2) {
#pragma unroll
              for (int u = 0; u < 16; ++u)
                pk[u] = pack_bf16(pr[2 * u] - __bfloat162float(__float2bfloat16_rn(pr[2 * u])),
                                  pr[2 * u + 1] - __bfloat162float(__float2bfloat16_rn(pr[2 * u + 1])));
              tmem_st16(tm_p + lane_addr + 64 + h * 32 + 16 * cc, pk);
            }
            continue;
          }
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
      // hand-over: P visible to the tensor core (smem: generic -> async proxy fence; TMEM: stores complete);
      // S reads / O rescales are done
      if (g + 1 < n_sub) request_s(g + 1);
      if (P_TMEM) tmem_st_wait();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_p[h]);
      tmem_ld_wait();
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
    tmem_dealloc(tmem_base, TMEM_COLS_SUB);
  }
}


template <int NSPLIT, int VAR>
static int launch_attn_sub(const tng_attn_desc* d, const CUtensorMap& qm, const CUtensorMap& km, const CUtensorMap& vm,
                           const AttnParams& p, cudaStream_t st) {
  using Cfg = AttnSubCfg<NSPLIT, VAR>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_sub_kernel<NSPLIT, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(attention_sub): %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((d->Lq + AT_BM - 1) / AT_BM, d->heads, d->batch);
  attention_sub_kernel<NSPLIT, VAR><<<grid, AS_THREADS, Cfg::SMEM_BYTES, st>>>(qm, km, vm, p);
  count_launch();
  return check_launch("attention_sub");
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
  if (d->nsplit == 2) return launch_attn_sub<2, 6>(d, qm, km, vm, p, st);
  return launch_attn_sub<1, 6>(d, qm, km, vm, p, st);
}
