// attention_wide.cu — tcgen05 flash attention for ONE head of width 512: the AudioLDM VAE AttnBlock
// (audioldm/variational_autoencoder/modules.py:204-230: softmax(q k^T / sqrt(C)) v over the H*W positions of an image,
// C = 512, 4096 positions for a 10 s clip, 12288 for 30 s). The [HW, HW] score matrix never leaves the SM.
//
// One CTA = 128 query rows of one image x ONE HALF (256 columns) of the value / output width; grid (HW / 128, 2, batch).
//   smem  : Q [128 x 512] bf16 as 8 SWIZZLE_128B chunks (128 KB, loaded once), one K sub-tile [64 keys x 512] (64 KB),
//           one V sub-tile [64 keys x 256] (32 KB)  -> 224 KB, which is why K / V are single-buffered;
//   TMEM  : S [128 x 64] fp32 (64 columns), P [128 x 64] bf16 pairs (32 columns), O [128 x 256] fp32 (256 columns);
//   warps : 0-3 softmax (thread = query row = TMEM lane), warp 4 lane 0 = TMA loads + every tcgen05.mma.
// Per 64-key sub-tile g:  S = sum over the 8 head-dim chunks of Q_c K_c^T (32 MMAs 128x64x16)  ->  softmax warps: one
// pass over S against the lazily advanced reference maximum (rescale O only when a row grew by more than 2^8, as in
// attention_tc.cu), P -> TMEM  ->  O += P V (TS-mode MMAs, V consumed as an MN-major B operand, 64 output columns per
// instruction). K_{g+1} is requested the moment Q K_g^T has drained and V_{g+1} when P V_g has: the K load latency is
// the period of the loop (~2 us per sub-tile), everything else hides under it. Both halves recompute S (the score
// FLOPs double; they are 1/5 of a decoder that is itself < 1 % of a 200-step generation) so that no CTA needs more than
// 512 TMEM columns. bf16 operands only: the parity mode (hi/lo split) keeps the GEMM -> softmax -> GEMM formulation.
#include "tng_ptx.cuh"
#include "tng_internal.h"

namespace tng {

constexpr int AW_BM = 128;                 // queries per CTA
constexpr int AW_SUB = 64;                 // keys per sub-tile
constexpr int AW_D = 512;                  // head width
constexpr int AW_DV = 256;                 // value / output columns per CTA
constexpr int AW_QCH = AW_BM * 128;        // one [128][64] bf16 chunk = 16 KB
constexpr int AW_KCH = AW_SUB * 128;       // one [64][64] bf16 chunk = 8 KB
constexpr int AW_Q_BYTES = (AW_D / 64) * AW_QCH;       // 128 KB
constexpr int AW_K_BYTES = (AW_D / 64) * AW_KCH;       // 64 KB
constexpr int AW_V_BYTES = (AW_DV / 64) * AW_KCH;      // 32 KB
constexpr int AW_SMEM = AW_Q_BYTES + AW_K_BYTES + AW_V_BYTES + 128;
constexpr int AW_THREADS = 160;
constexpr float AW_LAZY = 8.0f;

struct AttnWideParams {
  int L;
  int q_col0, k_col0, v_col0;
  __nv_bfloat16* out;
  long long ld_o;
  float scale_log2e;
};

__global__ void __launch_bounds__(AW_THREADS, 1)
attention_wide_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                      const __grid_constant__ CUtensorMap vmap, const __grid_constant__ AttnWideParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AW_Q_BYTES;
  uint8_t* sV = sK + AW_K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AW_V_BYTES);
  uint64_t* bar_q = bars;        // Q landed
  uint64_t* bar_k = bars + 1;    // K sub-tile landed
  uint64_t* bar_v = bars + 2;    // V sub-tile landed
  uint64_t* bar_s = bars + 3;    // S_g complete (and the K buffer free)
  uint64_t* bar_p = bars + 4;    // P_g written, S_g read (128 arrivals)
  uint64_t* bar_pv = bars + 5;   // P V_g drained (V buffer, P and O safe)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * AW_BM;
  const int half = blockIdx.y;
  const int b = blockIdx.z;
  const int n_sub = p.L / AW_SUB;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[tng] attention_wide: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&qmap);
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&vmap);
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_p, 128);
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
  const uint32_t tm_s = tmem_base;          // 64 columns
  const uint32_t tm_p = tmem_base + 64;     // 32 columns (bf16 pairs)
  const uint32_t tm_o = tmem_base + 128;    // 256 columns

  if (warp == 4) {
    if (lane == 0) {
      // ================================================================= loader / MMA issuer
      auto load_k = [&](int g) {
        mbar_arrive_expect_tx(bar_k, AW_K_BYTES);
#pragma unroll
        for (int c = 0; c < AW_D / 64; ++c)
          tma_load_3d(sK + c * AW_KCH, &kmap, bar_k, p.k_col0 + 64 * c, g * AW_SUB, b);
      };
      auto load_v = [&](int g) {
        mbar_arrive_expect_tx(bar_v, AW_V_BYTES);
#pragma unroll
        for (int c = 0; c < AW_DV / 64; ++c)
          tma_load_3d(sV + c * AW_KCH, &vmap, bar_v, p.v_col0 + half * AW_DV + 64 * c, g * AW_SUB, b);
      };
      constexpr uint32_t idesc_qk = umma_idesc_bf16(AW_BM, AW_SUB, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AW_BM, 64, 0, 1);   // B = V chunk, MN-major, one 64-column atom
      const uint64_t qdesc0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV), 1024, 1024);

      mbar_arrive_expect_tx(bar_q, AW_Q_BYTES);
#pragma unroll
      for (int c = 0; c < AW_D / 64; ++c) tma_load_3d(sQ + c * AW_QCH, &qmap, bar_q, p.q_col0 + 64 * c, q0, b);
      load_k(0);
      load_v(0);
      mbar_wait(bar_q, 0);
      for (int g = 0; g < n_sub; ++g) {
        const uint32_t ph = g & 1;
        // ---- S_g = Q K_g^T over the 512-wide head: 8 chunks x 4 K-steps of 16
        mbar_wait(bar_k, ph);
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < AW_D / 64; ++c) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tm_s, qdesc0 + static_cast<uint64_t>((c * AW_QCH) >> 4) + 2 * k,
                      kdesc0 + static_cast<uint64_t>((c * AW_KCH) >> 4) + 2 * k, idesc_qk, acc);
            acc = 1;
          }
        }
        umma_commit(bar_s);
        mbar_wait(bar_s, ph);                      // Q K_g^T drained: the K buffer is free
        if (g + 1 < n_sub) load_k(g + 1);
        // ---- O (+)= P_g V_g once the softmax warps have handed P_g over
        mbar_wait(bar_p, ph);
        mbar_wait(bar_v, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AW_SUB / 16; ++k) {     // 16 keys per MMA: 8 TMEM columns of P, 16 x 128 B rows of V
#pragma unroll
          for (int n = 0; n < AW_DV / 64; ++n) {
            const uint64_t bdesc = vdesc0 + static_cast<uint64_t>((n * AW_KCH + k * 2048) >> 4);
            umma_bf16_ts(tm_o + 64 * n, tm_p + 8 * k, bdesc, idesc_pv, (g > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_pv);
        mbar_wait(bar_pv, ph);                     // P V_g drained: V buffer and P are free, O is consistent
        if (g + 1 < n_sub) load_v(g + 1);
      }
    }
    __syncwarp();
  } else {
    // ================================================================= softmax warps: thread = query row = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t ts = tm_s + lane_addr;
    const uint32_t tp = tm_p + lane_addr;
    const uint32_t to = tm_o + lane_addr;
    float m_ref = -INFINITY;   // reference maximum used in the exponent (log2 domain)
    float l_run = 0.f;
    const float sc = p.scale_log2e;
    for (int g = 0; g < n_sub; ++g) {
      uint32_t va[32], vb[32];
      __syncwarp();
      mbar_wait(bar_s, g & 1);   // S_g complete; the issuer waited for P V_{g-1} before it issued this Q K^T
      tc_fence_after();
      tmem_ld32(ts, va);
      tmem_ld32(ts + 32, vb);
      tmem_ld_wait();
      if (g == 0) {
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(va[i + 1]), __uint_as_float(vb[i + 1])));
        }
        m_ref = fmaxf(m0, m1) * sc;   // scale > 0
      }
      float lsum = 0.f;
#pragma unroll 1
      for (int round = 0; round < 2; ++round) {
        float xmax = -INFINITY;
        lsum = 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const uint32_t* v = cc ? vb : va;
          float pr[32];
          const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_ref, -m_ref);
          float2 l01 = make_float2(0.f, 0.f), l23 = l01;
          float x0 = -INFINITY, x1 = -INFINITY;
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
          lsum += (l01.x + l01.y) + (l23.x + l23.y);
          xmax = fmaxf(xmax, fmaxf(x0, x1));
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) pk[u] = pack_bf16(pr[2 * u], pr[2 * u + 1]);
          tmem_st16(tp + 16 * cc, pk);
        }
        // lazy rescale (warp-uniform decision: tcgen05.ld / st are warp collectives)
        const bool need = xmax > AW_LAZY;
        if (round == 1 || !__any_sync(0xffffffffu, need)) break;
        const float f = need ? ex2_approx(-xmax) : 1.0f;   // 2^(m_ref - m_new), m_new = m_ref + xmax
        l_run *= f;
        if (g > 0) {   // O holds every P V issued so far (see the wait above)
#pragma unroll 1
          for (int c = 0; c < AW_DV; c += 32) {
            uint32_t w[32];
            tmem_ld32(to + c, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * f);
            tmem_st32(to + c, w);
          }
        }
        tmem_st_wait();
        if (need) m_ref += xmax;
        // the recompute against the new reference reads S_g again (va / vb still hold it)
      }
      l_run += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- finalize: O / l -> bf16
    __syncwarp();
    mbar_wait(bar_pv, (n_sub - 1) & 1);
    tc_fence_after();
    const int q = q0 + tid;
    const float inv = 1.0f / l_run;
    __nv_bfloat16* op = p.out + (static_cast<long long>(b) * p.L + q) * p.ld_o + half * AW_DV;
#pragma unroll 1
    for (int c = 0; c < AW_DV; c += 32) {
      uint32_t v[32];
      tmem_ld32(to + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
        *reinterpret_cast<uint4*>(op + c + i) = w;
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

}  // namespace tng

using namespace tng;

extern "C" int tng_attention_wide(const void* q, int64_t ld_q, int32_t q_col0, const void* k, int64_t ld_k, int32_t k_col0,
                                  const void* v, int64_t ld_v, int32_t v_col0, void* out, int64_t ld_o, int32_t batch,
                                  int32_t L, int32_t dim, float scale, void* stream) {
  if (!q || !k || !v || !out || batch <= 0 || L <= 0) return set_error(TNG_EINVAL, "attention_wide: bad argument");
  if (dim != AW_D) return set_error(TNG_EINVAL, "attention_wide: head width %d unsupported (512 only)", dim);
  if (L % AW_BM != 0) return set_error(TNG_EINVAL, "attention_wide: L = %d must be a multiple of %d", L, AW_BM);
  if (scale <= 0.f) return set_error(TNG_EINVAL, "attention_wide: scale must be positive");
  if (ld_o % 8 || (reinterpret_cast<uintptr_t>(out) & 15)) return set_error(TNG_EINVAL, "attention_wide: output must allow 16-byte stores");
  AttnWideParams p;
  p.L = L; p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_o = ld_o;
  p.scale_log2e = scale * 1.4426950408889634f;
  CUtensorMap qm, km, vm;
  {
    uint64_t dims[3] = {(uint64_t)ld_q, (uint64_t)L, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)ld_q * 2, (uint64_t)ld_q * 2 * (uint64_t)L};
    uint32_t box[3] = {64, AW_BM, 1};
    int rc = encode_tmap_bf16(&qm, q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)ld_k, (uint64_t)L, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)ld_k * 2, (uint64_t)ld_k * 2 * (uint64_t)L};
    uint32_t box[3] = {64, AW_SUB, 1};
    int rc = encode_tmap_bf16(&km, k, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)ld_v, (uint64_t)L, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)ld_v * 2, (uint64_t)ld_v * 2 * (uint64_t)L};
    uint32_t box[3] = {64, AW_SUB, 1};
    int rc = encode_tmap_bf16(&vm, v, 3, dims, str, box);
    if (rc) return rc;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM);
    if (e != cudaSuccess) return set_error(TNG_ECUDA, "cudaFuncSetAttribute(attention_wide): %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid(L / AW_BM, AW_D / AW_DV, batch);
  attention_wide_kernel<<<grid, AW_THREADS, AW_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(qm, km, vm, p);
  count_launch();
  return check_launch("attention_wide");
}
