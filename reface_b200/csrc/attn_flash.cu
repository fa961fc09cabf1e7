// Fused self-attention (ldm/modules/attention.py:204-220; model.py:178-202; CLIP MHA):
//   O = softmax(scale * Q K^T) V   per (sample, head), never materialising the L x L score matrix.
//
// sm_100a design (all kernels): KV streamed in tiles of 128 keys through a TMA ring (4-D tensor maps over the fused
// [N*L, 3C] projection; head slices narrower than a 64-column chunk are zero-filled by TMA); S = Q.K^T by tcgen05.mma
// into TMEM; the softmax warps pull score rows out of TMEM (tcgen05.ld), exponentiate with the scale folded in
// (ex2.approx), write P back to TENSOR MEMORY as fp16 pairs (tcgen05.st) where it is the A operand of the P.V MMA;
// O accumulates in TMEM across KV tiles with lazy rescaling (the running maximum is only raised when it grows by more
// than 2^8, so p <= 256 stays exact in fp16 and the O read-modify-write is a rare warp-uniform branch).
//   attn_flash3_kernel : 128 queries per CTA, 8 softmax warps (two per TMEM lane quadrant, half a score row each)
//   attn_flash4_kernel : 256 queries per CTA (two query tiles share every K/V tile), 16 softmax warps
// (The first two generations of this kernel -- P through shared memory, O rescaled in registers every tile -- were
// removed in round 2; their measurements are in profiles/r01_attn_flash_*_ncu_full.txt.)
#include <algorithm>

#include "engine.h"
#include "ptx.cuh"

namespace rfb {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

// ------------------------------------------------------------------------------------------------ version 3
// v2 left the exp (MUFU) pipe at 62 % (ncu: profiles/r01s2_attn_v2_ncu.txt): the 4 softmax warps of a CTA spent
// more than half of their time outside the exponentials (TMEM load latency, the P tile's st.shared + proxy fence,
// the max chain), and with only two softmax warps per SM sub-partition nothing covered it.  v3:
//   * EIGHT softmax warps per CTA: two warps share a TMEM lane quadrant, each thread owns HALF a score row (64
//     columns, ~100 registers); the row maximum is exchanged through 4 bytes of smem and a 64-thread named barrier;
//   * P goes straight back to TENSOR MEMORY (tcgen05.st, fp16 pairs) and is the A operand of the P.V MMA
//     (tcgen05.mma with A in TMEM): no smem P tile, no generic->async proxy fence;
//   * POLY of every 8 exponentials are evaluated on the FMA pipe (Cody-Waite range reduction + degree-3 minimax
//     polynomial, |rel err| < 1.1e-4, below the fp16 rounding of P) so that the MUFU pipe is no longer the only
//     resource: with POLY = 2 issue slots and MUFU slots balance.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem, fp16 pairs: lane = row, 8 columns per K=16 step] x B[smem descriptor]
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 2^x for x <= ~16 on the FMA/ALU pipes: n = round(x) through the 1.5*2^23 trick, 2^(x-n) by a degree-3 minimax
// polynomial on [-0.5, 0.5], exponent inserted with an integer add.  x is clamped at -125 (result ~2^-125 ~ 0).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05500902608036995f, f, 0.2422109842300415f);
  p = fmaf(p, f, 0.6932829022407532f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int DP>
struct Flash3Cfg {
  static constexpr int NC = (DP + 63) / 64;
  static constexpr int CHUNK = 128 * 128;
  static constexpr int Q_BYTES = NC * CHUNK;
  static constexpr int KV_BYTES = NC * CHUNK;
  static constexpr int XCHG_BYTES = 2 * 2 * 128 * 4;  // row-max exchange [tile parity][half][row]
  // COMPACT (64 < head dim <= 128, i.e. d = 80): P is written over the first 64 columns of the S tile it was computed
  // from, so S | O fit 256 TMEM columns and TWO CTAs share an SM (16 softmax warps keep the MUFU unit busy; with one CTA
  // the 8 warps of the 32x32 level left it half idle).  Price: Q.K^T of tile j+1 can only be issued after P.V of tile j
  // (it would overwrite P) -- the other CTA's work fills that gap.  One K / V stage each.
  static constexpr bool COMPACT = DP > 64 && DP <= 128;
  // K / V ring depth: two stages where they fit next to Q (head dim <= 64), one for COMPACT and for the 160-wide heads
  // of the 16x16 / 8x8 levels (L <= 256: at most two KV tiles anyway)
  static constexpr int STAGES = (!COMPACT && Q_BYTES + 4 * KV_BYTES + XCHG_BYTES + 128 <= 227 * 1024) ? 2 : 1;
  static constexpr int SMEM = Q_BYTES + 2 * STAGES * KV_BYTES + XCHG_BYTES + 128;
  static constexpr int TMEM_COLS = (DP <= 128) ? 256 : 512;
  static constexpr int P_COL = COMPACT ? 0 : ((DP <= 64) ? 192 : 320);  // S [0,128) | O [128,128+DP) | P 64 columns
  static constexpr int SPLIT = (DP <= 48) ? 24 : (DP <= 64 ? 32 : (DP <= 80 ? 40 : 80));  // output columns of a quadrant's first warp
  static constexpr int MIN_CTAS = (DP <= 128) ? 2 : 1;
  static_assert(COMPACT || (P_COL >= 128 + DP && P_COL + 64 <= TMEM_COLS), "TMEM layout");
  static_assert(128 + DP <= TMEM_COLS, "TMEM layout");
};

template <int DP, int POLY>
__global__ void __launch_bounds__(320, Flash3Cfg<DP>::MIN_CTAS)
attn_flash3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, long long ldo, int L, int heads,
                   int d, float scale_log2e, unsigned long long* __restrict__ dbg) {
  using Cfg = Flash3Cfg<DP>;
  // optional clock64 trace of ONE CTA (option gemm_debug): dbg[tile * 16 + slot], tiles < 32
  const bool trace = dbg != nullptr && blockIdx.x == 7 && blockIdx.y == 37;
#define RFB_STAMP(j, slot)                                                   \
  do {                                                                       \
    if (trace && (j) < 32 && lane == 0) dbg[(j) * 16 + (slot)] = clock64();  \
  } while (0)
  constexpr int NC = Cfg::NC;
  constexpr float TAU = 8.0f;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::Q_BYTES;
  const uint32_t sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  const uint32_t sX = sV + Cfg::STAGES * Cfg::KV_BYTES;
  const uint32_t bars = sX + Cfg::XCHG_BYTES;
  const uint32_t b_q = bars, b_kf = bars + 8, b_ke = bars + 24, b_vf = bars + 40, b_ve = bars + 56;
  const uint32_t b_sfull = bars + 72, b_sfree = bars + 80, b_pfull = bars + 88, b_pfree = bars + 96;
  const uint32_t tptr = bars + 104;

  const int qt = blockIdx.x;
  const int z = blockIdx.y;
  const int n = z / heads, head = z % heads;
  const int ntiles = (L + 127) / 128;  // keys beyond L (zero-filled by TMA) are masked in the last tile

  if (threadIdx.x == 0) {
    mbar_init(b_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_kf + 8 * i, 1);
      mbar_init(b_ke + 8 * i, 1);
      mbar_init(b_vf + 8 * i, 1);
      mbar_init(b_ve + 8 * i, 1);
    }
    mbar_init(b_sfull, 1);
    mbar_init(b_sfree, 8);  // one arrival per softmax warp
    mbar_init(b_pfull, 8);
    mbar_init(b_pfree, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tptr, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: everything above overlapped the previous kernel's tail, nothing below may run before
  // that kernel has completed
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + Cfg::P_COL;

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
      const int s = j % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(j / Cfg::STAGES) & 1u;
      mbar_wait(b_ke + 8 * s, ph ^ 1u);
      RFB_STAMP(j, 0);
      if (elect_one()) {
        mbar_expect_tx(b_kf + 8 * s, Cfg::KV_BYTES);
        for (int c = 0; c < NC; ++c)
          tma_load_4d(sK + s * Cfg::KV_BYTES + c * Cfg::CHUNK, &tmK, b_kf + 8 * s, c * 64, head, j * 128, n);
      }
      __syncwarp();
      mbar_wait(b_ve + 8 * s, ph ^ 1u);
      RFB_STAMP(j, 1);
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
    for (int j = 0; j <= ntiles; ++j) {
      // COMPACT: P(j-1) occupies the S columns, so P.V(j-1) is issued BEFORE Q.K^T(j) (tcgen05.mma executes in issue order)
      auto issue_pv = [&]() {
        if (j > 0) {
        const int jj = j - 1, s = jj % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(jj / Cfg::STAGES) & 1u;
        mbar_wait(b_vf + 8 * s, ph);
        RFB_STAMP(jj, 12);
        mbar_wait(b_pfull, (uint32_t)jj & 1u);
        RFB_STAMP(jj, 3);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t db = smem_desc_mn_sw128(sV + s * Cfg::KV_BYTES + k * 2048, Cfg::CHUNK);
            mma_f16_ts(tO, tP + 8u * k, db, idesc_o, (jj > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(b_ve + 8 * s);
          mma_commit(b_pfree);
        }
        __syncwarp();
        }
      };
      if (Cfg::COMPACT) issue_pv();
      if (j < ntiles) {
        const int s = j % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(j / Cfg::STAGES) & 1u;
        mbar_wait(b_kf + 8 * s, ph);
        RFB_STAMP(j, 10);
        if (j > 0) mbar_wait(b_sfree, (uint32_t)(j - 1) & 1u);
        RFB_STAMP(j, 2);
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
      if (!Cfg::COMPACT) issue_pv();
    }
  } else {
    // ------------------------------------------------------------------ softmax warps 2..9
    const int q = warp & 3;            // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int hh = (warp - 2) >> 2;    // which 64-column half of the score row this thread owns
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem_raw + (sX - base));  // [parity][half][row]
    float m = 0.f, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_sfull, (uint32_t)j & 1u);
      if (warp == 2) RFB_STAMP(j, 4);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(tS + lane_off + 64 * hh, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
      tmem_ld32(tS + lane_off + 64 * hh + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
      tmem_ld_wait();
      if (warp == 2) RFB_STAMP(j, 5);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_sfree);
      if ((L & 127) && j == ntiles - 1) {  // ragged sequence (CLIP: 257 tokens; 8x8 maps: 64): keys >= L score -inf
        const int valid = L - j * 128 - 64 * hh;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) sv[i] = 0xff800000u;
      }
      float mr0 = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
      float mr1 = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
        mr0 = fmax3(mr0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mr1 = fmax3(mr1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      const float mh = fmaxf(mr0, mr1);
      xch[((j & 1) * 2 + hh) * 128 + r] = mh;
      named_bar_sync(1 + q, 64);
      if (warp == 2) RFB_STAMP(j, 6);
      const float mo = xch[((j & 1) * 2 + (hh ^ 1)) * 128 + r];
      const float mnew = fmaxf(mh, mo) * scale_log2e;
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
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = fmaf(__uint_as_float(sv[2 * i]), scale_log2e, -m);
        const float x1 = fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2e, -m);
        const float p0 = (((2 * i) & 7) < POLY) ? poly_exp2(x0) : fast_exp2(x0);
        const float p1 = (((2 * i + 1) & 7) < POLY) ? poly_exp2(x1) : fast_exp2(x1);
        rs0 += p0;
        rs1 += p1;
        pk[i] = pack_h2(p0, p1);
      }
      l = fmaf(l, alpha, rs0 + rs1);
      if (warp == 2) RFB_STAMP(j, 7);
      if (j > 0) mbar_wait(b_pfree, (uint32_t)(j - 1) & 1u);  // P.V(j-1) done: P free, O stable
      if (warp == 2) RFB_STAMP(j, 8);
      tc_fence_after();
      if (any && hh == 0) {  // warp-uniform, rare: bring the quadrant's O rows to the new maximum
#pragma unroll
        for (int c0 = 0; c0 < DP; c0 += 16) {
          uint32_t ov[16];
          tmem_ld16(tO + lane_off + c0, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tO + lane_off + c0, ov);
        }
      }
      tmem_st32(tP + lane_off + 32 * hh, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
      if (warp == 2) RFB_STAMP(j, 9);
    }
    // ---- all P.V MMAs done: row sums of the two halves, normalise, store
    mbar_wait(b_pfree, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    xch[hh * 128 + r] = l;  // parity-0 slots: their last use (tile ntiles-2 or earlier) is behind >= 1 named barrier
    named_bar_sync(1 + q, 64);
    const float inv = 1.0f / (l + xch[(hh ^ 1) * 128 + r]);
    __half* op = out + ((long long)n * L + (long long)qt * 128 + r) * ldo + (long long)head * d;
    const int c_begin = hh ? Cfg::SPLIT : 0, c_end = hh ? d : Cfg::SPLIT;
    const bool row_ok = qt * 128 + r < L;  // query rows beyond L exist only as zero-filled padding
#pragma unroll
    for (int cc = 0; cc < Cfg::SPLIT; cc += 8) {
      const int c0 = c_begin + cc;
      if (c0 < c_end) {  // warp-uniform
        uint32_t ov[8];
        tmem_ld8(tO + lane_off + c0, ov);
        tmem_ld_wait();
        uint4 u;
        u.x = pack_h2(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
        u.y = pack_h2(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
        u.z = pack_h2(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
        u.w = pack_h2(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
        if (row_ok) *reinterpret_cast<uint4*>(op + c0) = u;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
#undef RFB_STAMP
}

template <int DP>
static void launch_flash3(Ctx& c, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, __half* out,
                          long long ldo, int N, int L, int heads, int d, float scale) {
  using Cfg = Flash3Cfg<DP>;
  static const char* keys[] = {"flash3_48", "flash3_64", "flash3_80", "flash3_160"};
  if (c.first_use(keys[DP == 48 ? 0 : DP == 64 ? 1 : DP == 80 ? 2 : 3])) {
    CUDA_OK(cudaFuncSetAttribute(attn_flash3_kernel<DP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    CUDA_OK(cudaFuncSetAttribute(attn_flash3_kernel<DP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  }
  dim3 grid((unsigned)((L + 127) / 128), (unsigned)(N * heads));
  Ctx::ProfRec rec;
  if (c.profile) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    rec.flops = rec.flops_exec = 4.0 * (double)L * (double)L * (double)d * (double)N * (double)heads;
    rec.kind = 1;
    rec.M = L, rec.N = L, rec.K = d, rec.BN = DP, rec.z = N * heads;
    CUDA_OK(cudaEventRecord(rec.a, c.stream));
  }
  const float sl = scale * 1.4426950408889634f;
  unsigned long long* dbg = nullptr;
  if (c.gemm_debug) {
    if (!c.dbg_buf) CUDA_OK(cudaMalloc((void**)&c.dbg_buf, (size_t)c.num_sms * 8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c.dbg_buf, 0, (size_t)c.num_sms * 8 * sizeof(unsigned long long), c.stream));
    dbg = c.dbg_buf;
  }
  if (c.attn_poly) launch_pdl(c, attn_flash3_kernel<DP, 2>, grid, dim3(320), Cfg::SMEM, tq, tk, tv, out, ldo, L, heads, d, sl, dbg);
  else launch_pdl(c, attn_flash3_kernel<DP, 0>, grid, dim3(320), Cfg::SMEM, tq, tk, tv, out, ldo, L, heads, d, sl, dbg);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  if (c.profile) {
    CUDA_OK(cudaEventRecord(rec.b, c.stream));
    c.prof.push_back(rec);
  }
}


// ------------------------------------------------------------------------------------------------ version 4
// clock64 trace of v3 (scripts/attn_trace.py, profiles/r01s2_attn_v3_trace.txt): a K or V tile takes 3000-3800 cycles
// from TMA issue to mbarrier completion (128 strided 80-byte rows per box; 296 resident CTAs re-read all of K and V
// for every 128 queries), longer than a whole softmax tile -- the kernel was bound by K/V delivery, not by MUFU.
// v4 gives one CTA TWO query tiles (256 queries, one CTA per SM): every K/V tile is fetched once for both (half the
// L2->SM traffic and TMA row requests per unit of work) into a 4-deep ring, and 16 softmax warps (two groups of 8,
// organised as in v3) work on S0/S1 while one MMA warp feeds both.  TMEM (all 512 columns):
//   S0 [0,128) S1 [128,256) | O0 [256,320) O1 [320,384) | P0 [384,448) P1 [448,512)
template <int DP, int NS>
struct Flash4Cfg {
  static_assert(DP <= 64, "v4 keeps two O tiles of <= 64 columns in TMEM");
  static constexpr int CHUNK = 128 * 128;
  static constexpr int Q_BYTES = 2 * CHUNK;
  static constexpr int XCHG_BYTES = 2 * 2 * 2 * 128 * 4;  // [group][tile parity][half][row]
  static constexpr int SMEM = Q_BYTES + 2 * NS * CHUNK + XCHG_BYTES + 384;
};

template <int DP, int POLY, int NS, int PINGPONG>
__global__ void __launch_bounds__(576, 1)
attn_flash4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, long long ldo, int L, int heads,
                   int d, float scale_log2e, unsigned long long* __restrict__ dbg, int stagger) {
  using Cfg = Flash4Cfg<DP, NS>;
  constexpr float TAU = 8.0f;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::Q_BYTES;
  const uint32_t sV = sK + NS * Cfg::CHUNK;
  const uint32_t sX = sV + NS * Cfg::CHUNK;
  const uint32_t bars = sX + Cfg::XCHG_BYTES;
  // barrier map (8 bytes each): q | kf[NS] ke[NS] vf[NS] ve[NS] | sfull[2] sfree[2] pfull[2] pfree[2]
  const uint32_t b_q = bars, b_kf = bars + 8, b_ke = b_kf + 8 * NS, b_vf = b_ke + 8 * NS, b_ve = b_vf + 8 * NS;
  const uint32_t b_sfull = b_ve + 8 * NS, b_sfree = b_sfull + 16, b_pfull = b_sfree + 16, b_pfree = b_pfull + 16;
  const uint32_t tptr = b_pfree + 16;
  // MUFU turn-taking (see the softmax warps): b_turn[q][g] = "group g of lane quadrant q may start its exponentials"
  const uint32_t b_turn = tptr + 16;
  static_assert(8 + 4 * 8 * NS + 64 + 16 + 64 <= 384, "barrier area");

  const int qt = blockIdx.x;  // pair of query tiles
  const int z = blockIdx.y;
  const int n = z / heads, head = z % heads;
  const int ntiles = L / 128;
  const bool trace = dbg != nullptr && blockIdx.x == 3 && blockIdx.y == 37;
#define RFB_STAMP(j, slot)                                                   \
  do {                                                                       \
    if (trace && (j) < 32 && lane == 0) dbg[(j) * 16 + (slot)] = clock64();  \
  } while (0)
#define RFB_STAMP1(j, slot) /* the second query tile's warp of the same sub-partition: second half of the buffer */ \
  do {                                                                                                            \
    if (trace && (j) < 32 && lane == 0) dbg[512 + (j) * 16 + (slot)] = clock64();                                 \
  } while (0)

  if (threadIdx.x == 0) {
    mbar_init(b_q, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(b_kf + 8 * i, 1);
      mbar_init(b_ke + 8 * i, 1);
      mbar_init(b_vf + 8 * i, 1);
      mbar_init(b_ve + 8 * i, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(b_sfull + 8 * g, 1);
      mbar_init(b_sfree + 8 * g, 8);  // one arrival per softmax warp of the group
      mbar_init(b_pfull + 8 * g, 8);
      mbar_init(b_pfree + 8 * g, 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(b_turn + 8 * i, 2);  // the two column-half warps of the other group arrive
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: everything above overlapped the previous kernel's tail, nothing below may run before
  // that kernel has completed
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(b_q, Cfg::Q_BYTES);
      tma_load_4d(sQ, &tmQ, b_q, 0, head, qt * 256, n);
      tma_load_4d(sQ + Cfg::CHUNK, &tmQ, b_q, 0, head, qt * 256 + 128, n);
    }
    __syncwarp();
    uint32_t st = 0, ph = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(b_ke + 8 * st, ph ^ 1u);
      RFB_STAMP(j, 0);
      if (elect_one()) {
        mbar_expect_tx(b_kf + 8 * st, Cfg::CHUNK);
        tma_load_4d(sK + st * Cfg::CHUNK, &tmK, b_kf + 8 * st, 0, head, j * 128, n);
      }
      __syncwarp();
      mbar_wait(b_ve + 8 * st, ph ^ 1u);
      RFB_STAMP(j, 1);
      if (elect_one()) {
        mbar_expect_tx(b_vf + 8 * st, Cfg::CHUNK);
        tma_load_4d(sV + st * Cfg::CHUNK, &tmV, b_vf + 8 * st, 0, head, j * 128, n);
      }
      __syncwarp();
      if (++st == NS) st = 0, ph ^= 1u;
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = idesc_f16(128, 128);
    const uint32_t idesc_o = idesc_f16(128, DP, 0, 1);  // B (= V) is MN-major
    constexpr int KSTEPS = (DP + 15) / 16;
    mbar_wait(b_q, 0);
    uint32_t st = 0, ph = 0;      // ring position of tile j (Q.K^T)
    uint32_t st2 = 0, ph2 = 0;    // ring position of tile j-1 (P.V)
    for (int j = 0; j <= ntiles; ++j) {
      if (j < ntiles) {
        mbar_wait(b_kf + 8 * st, ph);
        RFB_STAMP(j, 10);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (j > 0) mbar_wait(b_sfree + 8 * g, (uint32_t)(j - 1) & 1u);
          if (g == 0) RFB_STAMP(j, 2);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = smem_desc_k_sw128(sQ + g * Cfg::CHUNK);
            const uint64_t db = smem_desc_k_sw128(sK + st * Cfg::CHUNK);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              mma_f16_ss(tmem_base + 128u * g, da + 2u * k, db + 2u * k, idesc_s, k > 0 ? 1u : 0u);
            if (g == 1) mma_commit(b_ke + 8 * st);
            mma_commit(b_sfull + 8 * g);
          }
          __syncwarp();
        }
        if (++st == NS) st = 0, ph ^= 1u;
      }
      if (j > 0) {
        const int jj = j - 1;
        mbar_wait(b_vf + 8 * st2, ph2);
        RFB_STAMP(jj, 12);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(b_pfull + 8 * g, (uint32_t)jj & 1u);
          if (g == 0) RFB_STAMP(jj, 3);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t db = smem_desc_mn_sw128(sV + st2 * Cfg::CHUNK + k * 2048, Cfg::CHUNK);
              mma_f16_ts(tmem_base + 256u + 64u * g, tmem_base + 384u + 64u * g + 8u * k, db, idesc_o,
                         (jj > 0 || k > 0) ? 1u : 0u);
            }
            if (g == 1) mma_commit(b_ve + 8 * st2);
            mma_commit(b_pfree + 8 * g);
          }
          __syncwarp();
        }
        if (++st2 == NS) st2 = 0, ph2 ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps 2..17
    const int g = (warp - 2) >> 3;           // query tile of this warp
    const int q = warp & 3;                  // TMEM lane quadrant (hardware: warp id % 4)
    const int hh = ((warp - 2) & 7) >> 2;    // 64-column half of the score row
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem_base + 128u * g, tO = tmem_base + 256u + 64u * g, tP = tmem_base + 384u + 64u * g;
    const uint32_t bsfull = b_sfull + 8 * g, bsfree = b_sfree + 8 * g, bpfull = b_pfull + 8 * g, bpfree = b_pfree + 8 * g;
    float* xch = reinterpret_cast<float*>(smem_raw + (sX - base)) + g * 512;  // [parity][half][row]
    const int bar_id = 1 + g * 4 + q;
    const bool tw = (warp == 2), tw1 = (warp == 10);
    float m = 0.f, l = 0.f;
    if (g == 1 && stagger > 0) {
      // start the second query tile's softmax half a tile late: its TMEM load / max / P-store phases then overlap
      // the first tile's exponentials instead of leaving the MUFU pipe idle in both groups at once
      const long long t0 = clock64();
      while (clock64() - t0 < stagger) {
      }
    }
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(bsfull, (uint32_t)j & 1u);
      if (tw) RFB_STAMP(j, 4);
      if (tw1) RFB_STAMP1(j, 4);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(tS + lane_off + 64 * hh, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
      tmem_ld32(tS + lane_off + 64 * hh + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
      tmem_ld_wait();
      if (tw) RFB_STAMP(j, 5);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bsfree);
      float mr0 = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
      float mr1 = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
        mr0 = fmax3(mr0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mr1 = fmax3(mr1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      const float mh = fmaxf(mr0, mr1);
      xch[((j & 1) * 2 + hh) * 128 + r] = mh;
      named_bar_sync(bar_id, 64);
      if (tw) RFB_STAMP(j, 6);
      if (tw1) RFB_STAMP1(j, 6);
      const float mo = xch[((j & 1) * 2 + (hh ^ 1)) * 128 + r];
      const float mnew = fmaxf(mh, mo) * scale_log2e;
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
      uint32_t pk[32];
      // MUFU turn-taking.  The four softmax warps of a lane quadrant (two per query tile) share one SM sub-partition and
      // its 4-lane MUFU unit.  Left alone they run in lock-step -- all four in the exponential phase (MUFU saturated),
      // then all four in the TMEM load / row-max / P store phases (MUFU idle for ~600 of every ~2700 cycles, the 76 %
      // of profiles/r01s2_attn_flash4_ncu_full.txt): a start-up stagger does not survive, the group that gets ahead
      // runs alone at twice the rate until the MMA warp's in-order waits pin its lead at exactly one tile = in phase
      // again.  So the two groups take explicit turns: group 1 starts its exponentials when group 0 has finished those of
      // the same tile, group 0 when group 1 has finished the previous tile's; each group's load / max / store phases
      // then overlap the other group's exponentials.
      if (PINGPONG) {
        if (g == 1) mbar_wait(b_turn + 8 * (q * 2 + 1), (uint32_t)j & 1u);
        else if (j > 0) mbar_wait(b_turn + 8 * (q * 2), (uint32_t)(j - 1) & 1u);
      }
      if (tw) RFB_STAMP(j, 13);
      if (tw1) RFB_STAMP1(j, 13);
      // POLY of every 8 exponentials (whole fp16 pairs) go to the FMA pipe (option attn_poly; measured slower in every
      // arrangement -- interleaved per element, opposite order in the two column halves, and (round 2) ONE of the four
      // warps of a sub-partition doing all its exponentials by polynomial: 814-880 us against 700-730 us, that warp
      // becomes the straggler every tile waits for; profiles/r02_s3_attn_micro.txt).  A warp issues in order, so a warp that is
      // blocked on a full MUFU queue cannot reach polynomial work further down its stream: the two column halves of a
      // quadrant (which share an SM sub-partition) therefore run the two kinds in OPPOSITE order.
      auto exp_pair = [&](int i, bool poly) {
        const float x0 = fmaf(__uint_as_float(sv[2 * i]), scale_log2e, -m);
        const float x1 = fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2e, -m);
        const float p0 = poly ? poly_exp2(x0) : fast_exp2(x0);
        const float p1 = poly ? poly_exp2(x1) : fast_exp2(x1);
        rs0 += p0;
        rs1 += p1;
        pk[i] = pack_h2(p0, p1);
      };
      if (POLY == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) exp_pair(i, false);
      } else if (hh == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((2 * i) & 7) < POLY) exp_pair(i, true);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((2 * i) & 7) >= POLY) exp_pair(i, false);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((2 * i) & 7) >= POLY) exp_pair(i, false);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((2 * i) & 7) < POLY) exp_pair(i, true);
      }
      l = fmaf(l, alpha, rs0 + rs1);
      if (PINGPONG) {  // hand the MUFU unit to the other group's warps of this quadrant
        __syncwarp();
        if (lane == 0) mbar_arrive(b_turn + 8 * (q * 2 + (g ^ 1)));
      }
      if (tw) RFB_STAMP(j, 7);
      if (tw1) RFB_STAMP1(j, 7);
      if (j > 0) mbar_wait(bpfree, (uint32_t)(j - 1) & 1u);  // P.V(j-1) done: P free, O stable
      if (tw) RFB_STAMP(j, 8);
      tc_fence_after();
      if (any && hh == 0) {  // warp-uniform, rare
#pragma unroll
        for (int c0 = 0; c0 < DP; c0 += 16) {
          uint32_t ov[16];
          tmem_ld16(tO + lane_off + c0, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tO + lane_off + c0, ov);
        }
      }
      tmem_st32(tP + lane_off + 32 * hh, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bpfull);
      if (tw) RFB_STAMP(j, 9);
      if (tw1) RFB_STAMP1(j, 9);
    }
    mbar_wait(bpfree, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    xch[hh * 128 + r] = l;
    named_bar_sync(bar_id, 64);
    const float inv = 1.0f / (l + xch[(hh ^ 1) * 128 + r]);
    __half* op = out + ((long long)n * L + (long long)qt * 256 + g * 128 + r) * ldo + (long long)head * d;
    constexpr int SPLIT = 24;
    const int c_begin = hh ? SPLIT : 0, c_end = hh ? d : SPLIT;
#pragma unroll
    for (int cc = 0; cc < SPLIT; cc += 8) {
      const int c0 = c_begin + cc;
      if (c0 < c_end) {  // warp-uniform
        uint32_t ov[8];
        tmem_ld8(tO + lane_off + c0, ov);
        tmem_ld_wait();
        uint4 u;
        u.x = pack_h2(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
        u.y = pack_h2(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
        u.z = pack_h2(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
        u.w = pack_h2(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
        *reinterpret_cast<uint4*>(op + c0) = u;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
#undef RFB_STAMP
#undef RFB_STAMP1
}

template <int DP>
static void launch_flash4(Ctx& c, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, __half* out,
                          long long ldo, int N, int L, int heads, int d, float scale) {
  constexpr int NS = 4;
  using Cfg = Flash4Cfg<DP, NS>;
  if (c.first_use("flash4")) {
    CUDA_OK(cudaFuncSetAttribute(attn_flash4_kernel<DP, 0, NS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    CUDA_OK(cudaFuncSetAttribute(attn_flash4_kernel<DP, 0, NS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    CUDA_OK(cudaFuncSetAttribute(attn_flash4_kernel<DP, 2, NS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    CUDA_OK(cudaFuncSetAttribute(attn_flash4_kernel<DP, 2, NS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  }
  dim3 grid((unsigned)(L / 256), (unsigned)(N * heads));
  Ctx::ProfRec rec;
  if (c.profile) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    rec.flops = rec.flops_exec = 4.0 * (double)L * (double)L * (double)d * (double)N * (double)heads;
    rec.kind = 1;
    rec.M = L, rec.N = L, rec.K = d, rec.BN = DP, rec.z = N * heads;
    CUDA_OK(cudaEventRecord(rec.a, c.stream));
  }
  const float sl = scale * 1.4426950408889634f;
  unsigned long long* dbg = nullptr;
  if (c.gemm_debug) {
    if (!c.dbg_buf) CUDA_OK(cudaMalloc((void**)&c.dbg_buf, (size_t)c.num_sms * 8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c.dbg_buf, 0, (size_t)c.num_sms * 8 * sizeof(unsigned long long), c.stream));
    dbg = c.dbg_buf;
  }
#define RFB_FLASH4(POLY_, PP_)                                                                                    \
  launch_pdl(c, attn_flash4_kernel<DP, POLY_, NS, PP_>, grid, dim3(576), Cfg::SMEM, tq, tk, tv, out, ldo, L, heads, d, sl, dbg, \
             c.attn_stagger)
  if (c.attn_poly) {
    if (c.attn_pingpong) RFB_FLASH4(2, 1);
    else RFB_FLASH4(2, 0);
  } else {
    if (c.attn_pingpong) RFB_FLASH4(0, 1);
    else RFB_FLASH4(0, 0);
  }
#undef RFB_FLASH4
  CUDA_OK(cudaGetLastError());
  c.launches++;
  if (c.profile) {
    CUDA_OK(cudaEventRecord(rec.b, c.stream));
    c.prof.push_back(rec);
  }
}

bool attention_flash(Ctx& c, const __half* qkv, long long ldq, int N, int L, int heads, int d, __half* out,
                     long long ldo, float scale, int q_off, int k_off, int v_off, int hs) {
  // fused kernels: head dims 40 / 64 / 80 / 160 (UNet 64^2 / CLIP / 32^2 / 16^2 + 8^2 levels), any sequence length (keys
  // beyond L are masked); d = 512 (the VAE's single-head AttnBlock) stays on the materialised path
  if ((d != 40 && d != 64 && d != 80 && d != 160) || L < 1) return false;
  if (hs <= 0) hs = d;
  // padded head slices (hs > d, zero filled): the whole 64-column box is in bounds and every row is one aligned line
  const int ncols = d <= 64 ? 64 : (d <= 128 ? 128 : 192);
  const uint64_t dims[4] = {(uint64_t)std::min(hs, ncols), (uint64_t)heads, (uint64_t)L, (uint64_t)N};
  const uint64_t str[3] = {(uint64_t)hs * 2, (uint64_t)ldq * 2, (uint64_t)L * ldq * 2};
  const uint32_t box[4] = {64, 1, 128, 1};
  CUtensorMap tq = make_tmap(c, qkv + q_off, 4, dims, str, box);
  CUtensorMap tk = make_tmap(c, qkv + k_off, 4, dims, str, box);
  CUtensorMap tv = make_tmap(c, qkv + v_off, 4, dims, str, box);
  if (c.attn_flash >= 4 && d == 40 && L % 256 == 0) {
    launch_flash4<48>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale);
    return true;
  }
  switch (d) {
    case 40: launch_flash3<48>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale); break;
    case 64: launch_flash3<64>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale); break;
    case 80: launch_flash3<80>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale); break;
    default: launch_flash3<160>(c, tq, tk, tv, out, ldo, N, L, heads, d, scale); break;
  }
  return true;
}

}  // namespace rfb
