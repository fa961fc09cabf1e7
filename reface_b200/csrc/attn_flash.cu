// Fused self-attention for the UNet's SpatialTransformer blocks (ldm/modules/attention.py:204-220):
//   O = softmax(scale * Q K^T) V   per (sample, head), never materialising the L x L score matrix.
//
// sm_100a design: one CTA = 128 queries of one (sample, head); KV streamed in tiles of 128 keys.
//   warp0  TMA producer : Q once, then K/V tiles into a 2-stage smem ring (4-D tensor maps over the fused
//                         [N*L, 3C] projection; head slices narrower than 64 columns are zero-filled by TMA)
//   warp1  MMA issuer   : S = Q.K^T  (tcgen05.mma, K-major A/B, fp32 in TMEM cols [0,128))
//                         O_j = P.V  (A = P from smem, B = V as MN-major operand, TMEM cols [128,128+DP))
//   warps2-5 softmax    : thread t owns query row t: tcgen05.ld the S row, online max/sum in registers
//                         (no shuffles), exp2 with the scale folded in, P -> smem (128B-swizzled, fp16),
//                         O_acc = alpha*O_acc + O_j in registers, final O/l -> global fp16.
// Two CTAs are resident per SM for head dim 40 so that one CTA's exp work overlaps the other's MMAs.
#include "engine.h"
#include "ptx.cuh"

namespace rfb {

template <int DP>
struct FlashCfg {
  static constexpr int NC = (DP + 63) / 64;       // 64-column chunks of the head dim
  static constexpr int CHUNK = 128 * 128;          // bytes of one [128 x 64] fp16 tile
  static constexpr int Q_BYTES = NC * CHUNK;
  static constexpr int KV_BYTES = NC * CHUNK;      // per stage, K or V
  static constexpr int P_BYTES = 2 * CHUNK;
  static constexpr int STAGES = 2;
  static constexpr int SMEM = Q_BYTES + 2 * STAGES * KV_BYTES + P_BYTES + 128;  // base is 1024-aligned (checked)
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DP>
__global__ void __launch_bounds__(192, (DP <= 48) ? 2 : 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, long long ldo, int L, int heads,
                  int d, float scale_log2e) {
  using Cfg = FlashCfg<DP>;
  constexpr int NC = Cfg::NC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // warp index via shuffle = provably warp-uniform: TMA / MMA operands stay in uniform registers (no R2UR per issue)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();  // 128-byte swizzle atoms need a 1024-byte aligned tile base
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::Q_BYTES;
  const uint32_t sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  const uint32_t sP = sV + Cfg::STAGES * Cfg::KV_BYTES;
  const uint32_t bars = sP + Cfg::P_BYTES;
  const uint32_t b_q = bars, b_kf = bars + 8, b_ke = bars + 24, b_vf = bars + 40, b_ve = bars + 56;
  const uint32_t b_sfull = bars + 72, b_sfree = bars + 80, b_pfull = bars + 88, b_ofull = bars + 96;
  const uint32_t tptr = bars + 104;

  const int qt = blockIdx.x;       // query tile
  const int z = blockIdx.y;        // n*heads + head
  const int n = z / heads, head = z % heads;
  const int ntiles = L / 128;

  if (threadIdx.x == 0) {
    mbar_init(b_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_kf + 8 * i, 1);
      mbar_init(b_ke + 8 * i, 1);
      mbar_init(b_vf + 8 * i, 1);
      mbar_init(b_ve + 8 * i, 1);
    }
    mbar_init(b_sfull, 1);
    mbar_init(b_sfree, 4);  // one arrival per softmax warp
    mbar_init(b_pfull, 4);
    mbar_init(b_ofull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(b_q, Cfg::Q_BYTES);
      for (int c = 0; c < NC; ++c) tma_load_4d(sQ + c * Cfg::CHUNK, &tmQ, b_q, c * 64, head, qt * 128, n);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1u;
      mbar_wait(b_ke + 8 * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(b_kf + 8 * s, Cfg::KV_BYTES);
        for (int c = 0; c < NC; ++c)
          tma_load_4d(sK + s * Cfg::KV_BYTES + c * Cfg::CHUNK, &tmK, b_kf + 8 * s, c * 64, head, j * 128, n);
      }
      __syncwarp();
      mbar_wait(b_ve + 8 * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(b_vf + 8 * s, Cfg::KV_BYTES);
        for (int c = 0; c < NC; ++c)
          tma_load_4d(sV + s * Cfg::KV_BYTES + c * Cfg::CHUNK, &tmV, b_vf + 8 * s, c * 64, head, j * 128, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = idesc_f16(128, 128);
    const uint32_t idesc_o = idesc_f16(128, DP, 0, 1);  // B (= V) is MN-major
    mbar_wait(b_q, 0);
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1u;
      // ---- S = Q K^T
      mbar_wait(b_kf + 8 * s, ph);
      if (j > 0) mbar_wait(b_sfree, (uint32_t)(j - 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        int first = 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int ksteps = (c == NC - 1) ? ((DP - c * 64) + 15) / 16 : 4;
          const uint64_t da = smem_desc_k_sw128(sQ + c * Cfg::CHUNK);
          const uint64_t db = smem_desc_k_sw128(sK + s * Cfg::KV_BYTES + c * Cfg::CHUNK);
#pragma unroll
          for (int k = 0; k < ksteps; ++k) {
            mma_f16_ss(tS, da + 2u * k, db + 2u * k, idesc_s, first ? 0u : 1u);
            first = 0;
          }
        }
        mma_commit(b_ke + 8 * s);
        mma_commit(b_sfull);
      }
      __syncwarp();
      // ---- O_j = P V
      mbar_wait(b_vf + 8 * s, ph);
      mbar_wait(b_pfull, (uint32_t)j & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = smem_desc_k_sw128(sP + (k >> 2) * Cfg::CHUNK) + 2u * (k & 3);
          // 16 keys further along K = 16 rows of 128 B inside the MN-major tile
          const uint64_t db = smem_desc_mn_sw128(sV + s * Cfg::KV_BYTES + k * 2048, Cfg::CHUNK);
          mma_f16_ss(tO, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        mma_commit(b_ve + 8 * s);
        mma_commit(b_ofull);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax / accumulate warps
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float oacc[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) oacc[i] = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_sfull, (uint32_t)j & 1u);
      tc_fence_after();
      // pass 1: row max of the raw scores (scale > 0 is applied once afterwards); two TMEM loads per wait
      float mraw = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 64) {
        uint32_t s0[32], s1[32];
        tmem_ld32(tS + lane_off + c0, s0);
        tmem_ld32(tS + lane_off + c0 + 32, s1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
      }
      const float mx = fmaxf(m, mraw * scale_log2e);
      const float alpha = fast_exp2(m - mx);  // exp2(-inf) = 0 on the first tile
      m = mx;
      float rs = 0.f;
      // pass 2: p = exp2(s - m) -> fp16 -> swizzled smem (A operand of the P.V MMA)
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 64) {
        uint32_t sv[64];
        tmem_ld32(tS + lane_off + c0, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld32(tS + lane_off + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tmem_ld_wait();
        if (c0 == 64) {  // S fully consumed: let the MMA warp overwrite it with the next tile's scores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_sfree);
        }
        const uint32_t chunk_base = sP + (c0 >> 6) * Cfg::CHUNK + (uint32_t)r * 128u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(sv[8 * u + 2 * i]), scale_log2e, -mx));
            const float p1 = fast_exp2(fmaf(__uint_as_float(sv[8 * u + 2 * i + 1]), scale_log2e, -mx));
            rs += p0 + p1;
            pk[i] = pack_h2(p0, p1);
          }
          const uint32_t unit = (uint32_t)u ^ (uint32_t)(r & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(chunk_base + unit * 16u), "r"(pk[0]), "r"(pk[1]),
                       "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
      }
      l = l * alpha + rs;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
      // O_acc = alpha * O_acc + O_j
      mbar_wait(b_ofull, (uint32_t)j & 1u);
      tc_fence_after();
      {
        uint32_t ov[DP];
#pragma unroll
        for (int c0 = 0; c0 < DP; c0 += 16)
          tmem_ld16(tO + lane_off + c0, *reinterpret_cast<uint32_t(*)[16]>(&ov[c0]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < DP; ++i) oacc[i] = fmaf(oacc[i], alpha, __uint_as_float(ov[i]));
      }
      tc_fence_before();
    }
    // ---- normalise and store: out[(n*L + q0 + r) * ldo + head*d + i]
    const float inv = 1.0f / l;
    __half* op = out + ((long long)n * L + (long long)qt * 128 + r) * ldo + (long long)head * d;
#pragma unroll
    for (int c0 = 0; c0 < DP; c0 += 8) {
      if (c0 < d) {
        uint4 u;
        u.x = pack_h2(oacc[c0] * inv, oacc[c0 + 1] * inv);
        u.y = pack_h2(oacc[c0 + 2] * inv, oacc[c0 + 3] * inv);
        u.z = pack_h2(oacc[c0 + 4] * inv, oacc[c0 + 5] * inv);
        u.w = pack_h2(oacc[c0 + 6] * inv, oacc[c0 + 7] * inv);
        *reinterpret_cast<uint4*>(op + c0) = u;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ------------------------------------------------------------------------------------------------ version 2
// Same tiling and math as attn_flash_kernel, restructured so that the exp (MUFU) pipe is the only bound:
//   * ONE TMEM pass: a softmax thread pulls its whole 128-score row into registers and releases the S tile at once,
//     so the MMA warp issues Q.K^T of tile j+1 while tile j's exponentials are still being computed;
//   * O accumulates IN TMEM across the KV tiles (tcgen05.mma accumulate).  The running maximum is only raised when a
//     row's new maximum exceeds it by more than 2^8 ("lazy rescale"): p stays <= 256 (exact in fp16), the rescale
//     (tcgen05.ld -> * alpha -> tcgen05.st of the O rows) is a rare warp-uniform branch and the P.V MMAs leave the
//     softmax warps' critical path (no per-tile O read-back, no per-tile wait for the MMA);
//   * 3-input max (FMNMX3) for the row maximum.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int DP>
__global__ void __launch_bounds__(192, (DP <= 48) ? 2 : 1)
attn_flash2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, long long ldo, int L, int heads,
                   int d, float scale_log2e) {
  using Cfg = FlashCfg<DP>;
  constexpr int NC = Cfg::NC;
  constexpr float TAU = 8.0f;  // lazy-rescale threshold, log2 units
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::Q_BYTES;
  const uint32_t sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  const uint32_t sP = sV + Cfg::STAGES * Cfg::KV_BYTES;
  const uint32_t bars = sP + Cfg::P_BYTES;
  const uint32_t b_q = bars, b_kf = bars + 8, b_ke = bars + 24, b_vf = bars + 40, b_ve = bars + 56;
  const uint32_t b_sfull = bars + 72, b_sfree = bars + 80, b_pfull = bars + 88, b_pfree = bars + 96;
  const uint32_t tptr = bars + 104;

  const int qt = blockIdx.x;
  const int z = blockIdx.y;
  const int n = z / heads, head = z % heads;
  const int ntiles = L / 128;

  if (threadIdx.x == 0) {
    mbar_init(b_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_kf + 8 * i, 1);
      mbar_init(b_ke + 8 * i, 1);
      mbar_init(b_vf + 8 * i, 1);
      mbar_init(b_ve + 8 * i, 1);
    }
    mbar_init(b_sfull, 1);
    mbar_init(b_sfree, 4);  // one arrival per softmax warp
    mbar_init(b_pfull, 4);
    mbar_init(b_pfree, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(b_q, Cfg::Q_BYTES);
      for (int c = 0; c < NC; ++c) tma_load_4d(sQ + c * Cfg::CHUNK, &tmQ, b_q, c * 64, head, qt * 128, n);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1u;
      mbar_wait(b_ke + 8 * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(b_kf + 8 * s, Cfg::KV_BYTES);
        for (int c = 0; c < NC; ++c)
          tma_load_4d(sK + s * Cfg::KV_BYTES + c * Cfg::CHUNK, &tmK, b_kf + 8 * s, c * 64, head, j * 128, n);
      }
      __syncwarp();
      mbar_wait(b_ve + 8 * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(b_vf + 8 * s, Cfg::KV_BYTES);
        for (int c = 0; c < NC; ++c)
          tma_load_4d(sV + s * Cfg::KV_BYTES + c * Cfg::CHUNK, &tmV, b_vf + 8 * s, c * 64, head, j * 128, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = idesc_f16(128, 128);
    const uint32_t idesc_o = idesc_f16(128, DP, 0, 1);  // B (= V) is MN-major
    mbar_wait(b_q, 0);
    // issue order: S(0), [S(1), PV(0)], [S(2), PV(1)], ... : Q.K^T of the next tile is queued before P.V of this one
    for (int j = 0; j <= ntiles; ++j) {
      if (j < ntiles) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(b_kf + 8 * s, ph);
        if (j > 0) mbar_wait(b_sfree, (uint32_t)(j - 1) & 1u);  // S(j-1) sits in the softmax warps' registers
        tc_fence_after();
        if (elect_one()) {
          int first = 1;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int ksteps = (c == NC - 1) ? ((DP - c * 64) + 15) / 16 : 4;
            const uint64_t da = smem_desc_k_sw128(sQ + c * Cfg::CHUNK);
            const uint64_t db = smem_desc_k_sw128(sK + s * Cfg::KV_BYTES + c * Cfg::CHUNK);
#pragma unroll
            for (int k = 0; k < ksteps; ++k) {
              mma_f16_ss(tS, da + 2u * k, db + 2u * k, idesc_s, first ? 0u : 1u);
              first = 0;
            }
          }
          mma_commit(b_ke + 8 * s);
          mma_commit(b_sfull);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int jj = j - 1, s = jj & 1;
        const uint32_t ph = (uint32_t)(jj >> 1) & 1u;
        mbar_wait(b_vf + 8 * s, ph);
        mbar_wait(b_pfull, (uint32_t)jj & 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = smem_desc_k_sw128(sP + (k >> 2) * Cfg::CHUNK) + 2u * (k & 3);
            const uint64_t db = smem_desc_mn_sw128(sV + s * Cfg::KV_BYTES + k * 2048, Cfg::CHUNK);
            mma_f16_ss(tO, da, db, idesc_o, (jj > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(b_ve + 8 * s);
          mma_commit(b_pfree);  // P consumed, O holds tiles 0..jj
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float m = 0.f, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_sfull, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t sv[128];
      tmem_ld32(tS + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
      tmem_ld32(tS + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
      tmem_ld32(tS + lane_off + 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[64]));
      tmem_ld32(tS + lane_off + 96, *reinterpret_cast<uint32_t(*)[32]>(&sv[96]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_sfree);
      float mr0 = __uint_as_float(sv[0]), mr1 = __uint_as_float(sv[1]);
#pragma unroll
      for (int i = 2; i < 128; i += 4) {
        mr0 = fmax3(mr0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        if (i + 3 < 128) mr1 = fmax3(mr1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      const float mnew = fmaxf(mr0, mr1) * scale_log2e;
      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m = mnew;
      } else if (mnew > m + TAU) {
        need = true;
        alpha = fast_exp2(m - mnew);
        m = mnew;
      }
      const bool any = __any_sync(0xffffffffu, need);
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * i]), scale_log2e, -m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2e, -m));
        rs0 += p0;
        rs1 += p1;
        pk[i] = pack_h2(p0, p1);
      }
      l = fmaf(l, alpha, rs0 + rs1);
      // P.V of the previous tile must have completed: it reads the P buffer and writes O
      if (j > 0) mbar_wait(b_pfree, (uint32_t)(j - 1) & 1u);
      if (any) {  // warp-uniform, rare: bring this warp's O rows to the new maximum
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < DP; c0 += 16) {
          uint32_t ov[16];
          tmem_ld16(tO + lane_off + c0, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tO + lane_off + c0, ov);
        }
        tmem_st_wait();
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t chunk_base = sP + c * Cfg::CHUNK + (uint32_t)r * 128u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t unit = (uint32_t)u ^ (uint32_t)(r & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(chunk_base + unit * 16u), "r"(pk[c * 32 + 4 * u]),
                       "r"(pk[c * 32 + 4 * u + 1]), "r"(pk[c * 32 + 4 * u + 2]), "r"(pk[c * 32 + 4 * u + 3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
    }
    // ---- all P.V MMAs done: normalise and store
    mbar_wait(b_pfree, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    __half* op = out + ((long long)n * L + (long long)qt * 128 + r) * ldo + (long long)head * d;
#pragma unroll
    for (int c0 = 0; c0 < DP; c0 += 16) {
      uint32_t ov[16];
      tmem_ld16(tO + lane_off + c0, ov);
      tmem_ld_wait();
#pragma unroll
      for (int h8 = 0; h8 < 16; h8 += 8) {
        if (c0 + h8 < d) {
          uint4 u;
          u.x = pack_h2(__uint_as_float(ov[h8 + 0]) * inv, __uint_as_float(ov[h8 + 1]) * inv);
          u.y = pack_h2(__uint_as_float(ov[h8 + 2]) * inv, __uint_as_float(ov[h8 + 3]) * inv);
          u.z = pack_h2(__uint_as_float(ov[h8 + 4]) * inv, __uint_as_float(ov[h8 + 5]) * inv);
          u.w = pack_h2(__uint_as_float(ov[h8 + 6]) * inv, __uint_as_float(ov[h8 + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c0 + h8) = u;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int DP>
static void launch_flash(Ctx& c, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, __half* out,
                         long long ldo, int N, int L, int heads, int d, float scale) {
  using Cfg = FlashCfg<DP>;
  static bool attr = false;
  if (!attr) {
    CUDA_OK(cudaFuncSetAttribute(attn_flash_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    CUDA_OK(cudaFuncSetAttribute(attn_flash2_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  dim3 grid((unsigned)(L / 128), (unsigned)(N * heads));
  Ctx::ProfRec rec;
  if (c.profile) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    rec.flops = 4.0 * (double)L * (double)L * (double)d * (double)N * (double)heads;
    rec.kind = 1;
    rec.M = L, rec.N = L, rec.K = d, rec.BN = DP, rec.z = N * heads;
    CUDA_OK(cudaEventRecord(rec.a, c.stream));
  }
  if (c.attn_flash >= 2)
    attn_flash2_kernel<DP><<<grid, 192, Cfg::SMEM, c.stream>>>(tq, tk, tv, out, ldo, L, heads, d,
                                                             scale * 1.4426950408889634f);
  else
    attn_flash_kernel<DP><<<grid, 192, Cfg::SMEM, c.stream>>>(tq, tk, tv, out, ldo, L, heads, d,
                                                            scale * 1.4426950408889634f);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  if (c.profile) {
    CUDA_OK(cudaEventRecord(rec.b, c.stream));
    c.prof.push_back(rec);
  }
}

bool attention_flash(Ctx& c, const __half* qkv, long long ldq, int N, int L, int heads, int d, __half* out,
                     long long ldo, float scale, int q_off, int k_off, int v_off) {
  if (L % 128 != 0 || (d != 40 && d != 80) || (d % 8) != 0) return false;
  const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)L, (uint64_t)N};
  const uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ldq * 2, (uint64_t)L * ldq * 2};
  const uint32_t box[4] = {64, 1, 128, 1};
  CUtensorMap tq = make_tmap(c, qkv + q_off, 4, dims, str, box);
  CUtensorMap tk = make_tmap(c, qkv + k_off, 4, dims, str, box);
  CUtensorMap tv = make_tmap(c, qkv + v_off, 4, dims, str, box);
  if (d == 40) launch_flash<48>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale);
  else launch_flash<80>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale);
  return true;
}

}  // namespace rfb
