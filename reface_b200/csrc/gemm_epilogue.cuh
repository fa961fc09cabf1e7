// Epilogue of the persistent GEMM kernels: TMEM accumulators -> fused bias / time-embedding row / GEGLU / activation /
// residual -> fp16 rows in HBM, with COALESCED global traffic and the global-load latencies taken off the critical path.
//
// tcgen05.ld 32x32b hands every thread one accumulator ROW, so a direct store makes each warp instruction touch
// 32 different 128-byte lines (measured: ~2500 cycles of LSU time per 128x160 tile, and the same again for the
// residual read; the K=320 GEMMs of the UNet were epilogue-bound at 2.2-2.7x their HBM time).  Each epilogue warp owns
// two 4 KB shared-memory staging tiles (32 rows x 128 B, XOR-swizzled in 16-byte units) and a 512 B bias strip:
//   prefetch (BEFORE the accumulators are ready): residual rows of the warp's <=2 column chunks -> staging tiles with
//            8 lanes per row (4 lines per instruction instead of 32); bias columns -> smem strip
//   drain    : tcgen05.ld -> + bias (smem broadcast) [+ row vector] [GEGLU / activation] + staged residual ->
//              fp16 row back into the staging tile -> coalesced 16-byte stores
// (clock64 profile of the first version: 30 % of the epilogue time was the exposed residual-load latency, 37-49 % the
// per-chunk bias loads + math; the accumulator loads themselves were ~1 %.)
#pragma once
#include "gemm_tc.cuh"

namespace rfb {

static constexpr int EPI_STAGE_BYTES = 4096;               // one staging tile
static constexpr int EPI_WARP_BYTES = 4096 + 512;          // staging tile + bias strip per epilogue warp

// EPI_LEAN: EPI_FAST's arithmetic for launches whose EVERY tile meets the lean drain's conditions (checked on the host)
enum EpiMode { EPI_FAST = 0, EPI_GEGLU = 1, EPI_GENERIC = 2, EPI_LEAN = 3 };

// exact-erf GELU with a cheap erf (Abramowitz-Stegun 7.1.26, |err| < 1.5e-7: far below the fp16 output ulp)
// exact-erf GELU in 8 instructions and one MUFU: with u = |x| and E(u) = 1 - Phi(u) = 0.5 * erfc(u / sqrt 2),
//   gelu(x) = x * Phi(x) = max(x, 0) - u * E(u)          (both signs)
// and E(u) = 2^-q(u) with q a degree-5 polynomial fitted to -log2(0.5 erfc(u / sqrt 2)) so that the ABSOLUTE error of
// u * E(u) is minimal: |gelu error| < 5.1e-7 for every fp32 input (q is increasing for all u >= 0, so E -> 0 for large
// |x| and nothing overflows).  The GEGLU FF GEMMs at 64^2 (K = 320) are bound by the epilogue's instruction issue: the
// previous Abramowitz-Stegun erf needed ~17 instructions and two MUFUs per element.
__device__ __forceinline__ float gelu_exp2poly(float x) {
  const float u = fabsf(x);
  float q = fmaf(0.0004687140753958374f, u, -0.007054118439555168f);
  q = fmaf(q, u, 0.05175532400608063f);
  q = fmaf(q, u, 0.46006664633750916f);
  q = fmaf(q, u, 1.1507560014724731f);
  q = fmaf(q, u, 1.0000418424606323f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q));
  return fmaf(-u, e, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = 1.0f - p * t * __expf(-z * z);
  const float erf = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf);
}

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// Row of the OUTPUT tensor that accumulator row gm is written to: the identity, except for the folded upsample
// convolution, whose tile rows are input pixels (n, y, x) and whose results belong to output pixel (2y+py, 2x+px).
__device__ __forceinline__ long long epi_out_row(const GemmArgs& g, long long gm, int z) {
  if (!g.up) return gm;
  const long long n = gm >> g.up_hwlog2;
  const int r = (int)(gm & ((1ll << g.up_hwlog2) - 1));
  const int y = r >> g.up_wlog2, x = r & ((1 << g.up_wlog2) - 1);
  return (n << (g.up_hwlog2 + 2)) + ((long long)(2 * y + (z >> 1)) << (g.up_wlog2 + 1)) + (2 * x + (z & 1));
}

struct EpiTile {  // warp-uniform description of one output tile for one epilogue warp
  int ncols, NO, ocol_tile, halfN, z;
  long long m_base, zoff, r_base;  // r_base: first row of the residual tensor for this warp (m_base unless res_mod)
  bool vec_ok;
};

template <int MODE>
__device__ __forceinline__ EpiTile epi_tile_info(const GemmArgs& g, int q, int m_tile, int n_tile, int z) {
  EpiTile t;
  t.halfN = g.BN >> 1;
  t.ncols = (MODE == EPI_GEGLU) ? t.halfN : g.BN;
  t.NO = (MODE == EPI_GEGLU) ? (g.N >> 1) : g.N;
  t.ocol_tile = (MODE == EPI_GEGLU) ? n_tile * t.halfN : n_tile * g.BN;
  t.m_base = (long long)m_tile * GEMM_BM + q * 32;
  t.r_base = g.res_mod > 0 ? t.m_base % g.res_mod : t.m_base;
  t.z = z;
  t.zoff = (g.zdiv == 1) ? (long long)z * g.zs_outer  // plain / conv / per-sample batches: no division per tile
                         : (long long)(z / g.zdiv) * g.zs_outer + (long long)(z % g.zdiv) * g.zs_inner;
  t.vec_ok = g.out != nullptr && (t.NO & 7) == 0 && (g.ldo & 7) == 0 && (t.zoff & 7) == 0 &&
             (!g.res || (g.ldr & 7) == 0);
  return t;
}

// One tile AHEAD: this warp's bias columns into 4 registers per lane (2 chunks x 2 floats) and the residual lines of
// the tile into L2 -- issued before the current tile is drained, so the latencies are off the critical path.
template <int MODE, int NP>
__device__ __forceinline__ void epilogue_lookahead(const GemmArgs& g, const EpiTile& t, int lane, int half,
                                                   float (&nb)[4]) {
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int c0 = half * 64 + ci * (NP * 64);  // `half` = part index of this warp inside its lane quadrant
    nb[2 * ci] = nb[2 * ci + 1] = 0.f;
    if (c0 >= t.ncols) continue;
    const int ocol0 = t.ocol_tile + c0;
    if (g.bias && MODE != EPI_GEGLU) {
      const int c1 = ocol0 + lane, c2 = ocol0 + 32 + lane;
      if (c1 < g.N && lane < t.ncols - c0) nb[2 * ci] = __ldg(g.bias + c1);
      if (c2 < g.N && 32 + lane < t.ncols - c0) nb[2 * ci + 1] = __ldg(g.bias + c2);
    }
    if (g.res && t.vec_ok && ocol0 < t.NO) {
      const long long gm = t.m_base + lane;
      if (gm < g.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.res + t.zoff + (t.r_base + lane) * g.ldr + ocol0));
    }
  }
}

// Phase A -- runs while the MMAs of this tile are still in flight.
//   wbuf: this warp's smem region (staging tile, then the bias strip); nb: bias values from epilogue_lookahead
template <int MODE, int NP>
__device__ __forceinline__ void epilogue_prefetch(const GemmArgs& g, const EpiTile& t, uint32_t wbuf, int lane, int half,
                                                  int n_tile, const float (&nb)[4]) {
  const uint32_t bias_s = wbuf + EPI_STAGE_BYTES;
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int c0 = half * 64 + ci * (NP * 64);  // `half` = part index of this warp inside its lane quadrant
    if (c0 >= t.ncols) break;
    const int ocol0 = t.ocol_tile + c0;
    const int cvalid = min(64, min(t.ncols - c0, t.NO - ocol0));
    if (cvalid <= 0) continue;
    if (g.bias && MODE != EPI_GEGLU) {  // GEGLU reads its two bias vectors directly in epilogue_drain
      sts32f(bias_s + (uint32_t)(ci * 64 + lane) * 4u, nb[2 * ci]);
      sts32f(bias_s + (uint32_t)(ci * 64 + 32 + lane) * 4u, nb[2 * ci + 1]);
    }
    if (g.res && t.vec_ok) {
      // first chunk: residual rows straight into the (idle) staging tile; second chunk: pull its lines into L2
      // now, the staging tile is refilled when chunk 0 has been stored (see epilogue_drain)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), unit = lane & 7;
        const long long gm = t.m_base + row;
        const bool ok = gm < g.M && unit * 8 < cvalid;
        const __half* src = g.res + t.zoff + (t.r_base + row) * g.ldr + ocol0 + unit * 8;
        if (ci == 0) {
          uint4 val = make_uint4(0u, 0u, 0u, 0u);
          if (ok) val = __ldg(reinterpret_cast<const uint4*>(src));
          sts128(wbuf + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4), val);
        } else if (ok && unit == 0) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
        }
      }
    }
  }
  __syncwarp();
}

// Lean drain for the common case -- EPI_FAST, vectorised fp16 output, identity row map, M % 128 == 0, N % BN == 0,
// BN % 16 == 0, row vector (if any) float4-readable and constant over 32-row groups (engine.cu: epi_lean_ok).  The generic drain below spends ~980 warp
// instructions per tile on ~64 output columns per lane, most of them addressing, predication and register moves
// (profiles/r01s2_gemm_k320_ncu_full.txt: 19 % arithmetic); the clock64 role breakdown of round 2
// (profiles/r02_gemm_role_trace_before_lean_epilogue.txt) shows the K <= 640 GEMMs waiting on exactly this: epilogue warps
// 94 % busy (drain 61-86 %), the MMA warp idle 32-44 % of the time for a free accumulator stage.  Here every address is
// hoisted out of the loops, there are no per-element guards and the accumulators are loaded straight into float registers.
template <int NP>
__device__ __forceinline__ void epilogue_drain_lean(const GemmArgs& g, const EpiTile& t, uint32_t trow, uint32_t wbuf, int lane,
                                                    int half) {
  const uint32_t bias_s = wbuf + EPI_STAGE_BYTES;
  const int sw = lane & 7, unit = lane & 7, r0 = lane >> 3;
  const uint32_t my_row = wbuf + (uint32_t)lane * 128u;
  // per-lane addresses of the coalesced pass (8 lanes per row, 4 rows per instruction): row r0 + 4 i, 16-byte unit `unit`
  const uint32_t ld_off = wbuf + (uint32_t)r0 * 128u + (uint32_t)((unit ^ (r0 & 7)) << 4);  // + i * 512: (r0 + 4i) & 7 ...
  __half* const out_base = g.out + t.zoff + (t.m_base + r0) * g.ldo + t.ocol_tile + unit * 8;
  const long long out_step = 4 * g.ldo;
  const __half* const res_base = g.res ? g.res + t.zoff + (t.r_base + r0) * g.ldr + t.ocol_tile + unit * 8 : nullptr;
  const long long res_step = 4 * g.ldr;
  const float* const rv = g.rowvec ? g.rowvec + (long long)t.z * g.rowvec_zs + (t.m_base / g.rows_per_vec) * g.ldv + t.ocol_tile
                                   : nullptr;
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int c0 = half * 64 + ci * (NP * 64);
    if (c0 >= t.ncols) break;
    const int cvalid = min(64, t.ncols - c0);  // 64 or 32 (warp-uniform)
    if (ci == 1 && g.res) {  // chunk 0's residual was staged by epilogue_prefetch; stage this chunk's now
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + r0;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (unit * 8 < cvalid) val = __ldg(reinterpret_cast<const uint4*>(res_base + i * res_step + c0));
        sts128(wbuf + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4), val);
      }
      __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h * 32 >= cvalid) break;
      float v[32];
      tmem_ld32f(trow + (uint32_t)(c0 + h * 32), v);
      tmem_ld_wait();
      if (g.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = lds128f(bias_s + (uint32_t)(ci * 64 + h * 32 + 4 * j) * 4u);
          v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
        }
      }
      if (rv) {
        const float4* r4 = reinterpret_cast<const float4*>(rv + c0 + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(r4 + j);
          v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
        }
      }
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const uint32_t a = my_row + (uint32_t)(((h * 4 + u4) ^ sw) << 4);
        if (g.res) {
          const uint4 u = lds128(a);
          const float2 f0 = unpack_h2(u.x), f1 = unpack_h2(u.y), f2 = unpack_h2(u.z), f3 = unpack_h2(u.w);
          v[u4 * 8 + 0] += f0.x, v[u4 * 8 + 1] += f0.y, v[u4 * 8 + 2] += f1.x, v[u4 * 8 + 3] += f1.y;
          v[u4 * 8 + 4] += f2.x, v[u4 * 8 + 5] += f2.y, v[u4 * 8 + 6] += f3.x, v[u4 * 8 + 7] += f3.y;
        }
        uint4 o;
        o.x = pack_h2(v[u4 * 8 + 0], v[u4 * 8 + 1]);
        o.y = pack_h2(v[u4 * 8 + 2], v[u4 * 8 + 3]);
        o.z = pack_h2(v[u4 * 8 + 4], v[u4 * 8 + 5]);
        o.w = pack_h2(v[u4 * 8 + 6], v[u4 * 8 + 7]);
        sts128(a, o);
      }
    }
    __syncwarp();
    // staging tile -> HBM: 8 lanes x 16 B per row, 4 rows per instruction; statistics from the same registers
    uint4 val[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + r0;
      val[i] = lds128(wbuf + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4));
    }
    if (unit * 8 < cvalid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(out_base + i * out_step + c0) = val[i];
    }
    if (g.stats) {
      float cs[8], cq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] = cq[j] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t w[4] = {val[i].x, val[i].y, val[i].z, val[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_h2(w[e]);
          cs[2 * e] += f.x, cq[2 * e] = fmaf(f.x, f.x, cq[2 * e]);
          cs[2 * e + 1] += f.y, cq[2 * e + 1] = fmaf(f.y, f.y, cq[2 * e + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
        cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 8);
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
        cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 16);
      }
      if (lane < 8 && unit * 8 < cvalid) {
        float4* sp = reinterpret_cast<float4*>(g.stats + ((t.m_base >> 5) * g.N + t.ocol_tile + c0 + unit * 8) * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) sp[j] = make_float4(cs[2 * j], cq[2 * j], cs[2 * j + 1], cq[2 * j + 1]);
      }
    }
  }
  __syncwarp();  // staging tile / bias strip may be refilled by the next tile's prefetch
}

// Phase B -- drains this warp's 32 accumulator rows (TMEM lanes q*32..q*32+31).
//   trow : TMEM address of (lane quadrant, accumulator stage, column 0)
template <int MODE, int NP>
__device__ __forceinline__ void epilogue_drain(const GemmArgs& g, const EpiTile& t, uint32_t trow, uint32_t wbuf, int lane,
                                               int half, int n_tile) {
  if (MODE == EPI_LEAN) {  // compile-time: the lean kernel carries no generic drain (registers)
    epilogue_drain_lean<NP>(g, t, trow, wbuf, lane, half);
    return;
  }
  const int BN = g.BN;
  const long long m = t.m_base + lane;  // the row this thread owns next to TMEM
  const bool row_ok = m < g.M;
  const float* rv = (g.rowvec && row_ok) ? g.rowvec + (long long)t.z * g.rowvec_zs + (m / g.rows_per_vec) * g.ldv : nullptr;
  const uint32_t bias_s = wbuf + EPI_STAGE_BYTES;
  const int sw = lane & 7;

#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int c0 = half * 64 + ci * (NP * 64);  // `half` = part index of this warp inside its lane quadrant
    if (c0 >= t.ncols) break;
    const int ocol0 = t.ocol_tile + c0;
    const int cvalid = min(64, min(t.ncols - c0, t.NO - ocol0));  // warp-uniform
    if (cvalid <= 0) continue;
    const uint32_t stage = wbuf;
    const uint32_t my_row = stage + (uint32_t)lane * 128u;  // row-per-thread view of the staging tile
    if (ci == 1 && g.res && t.vec_ok) {
      __syncwarp();  // chunk 0's rows have been read out of the staging tile
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), unit = lane & 7;
        const long long gm = t.m_base + row;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (gm < g.M && unit * 8 < cvalid)
          val = __ldg(reinterpret_cast<const uint4*>(g.res + t.zoff + (t.r_base + row) * g.ldr + ocol0 + unit * 8));
        sts128(stage + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4), val);
      }
      __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h * 32 >= cvalid) break;  // warp-uniform
      const int cc = c0 + h * 32;   // column inside the tile
      const int oc = ocol0 + h * 32;
      uint32_t acc[32];
      float v[32];
      tmem_ld32(trow + (uint32_t)cc, acc);
      if (MODE == EPI_GEGLU) {
        uint32_t gat[32];
        tmem_ld32(trow + (uint32_t)(t.halfN + cc), gat);
        const float4* bx = reinterpret_cast<const float4*>(g.bias + n_tile * BN + cc);
        const float4* bg = reinterpret_cast<const float4*>(g.bias + n_tile * BN + t.halfN + cc);
        float4 b1[8], b2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b1[j] = __ldg(bx + j), b2[j] = __ldg(bg + j);  // in flight with the TMEM loads
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j + 0] = (__uint_as_float(acc[4 * j + 0]) + b1[j].x) * gelu_exp2poly(__uint_as_float(gat[4 * j + 0]) + b2[j].x);
          v[4 * j + 1] = (__uint_as_float(acc[4 * j + 1]) + b1[j].y) * gelu_exp2poly(__uint_as_float(gat[4 * j + 1]) + b2[j].y);
          v[4 * j + 2] = (__uint_as_float(acc[4 * j + 2]) + b1[j].z) * gelu_exp2poly(__uint_as_float(gat[4 * j + 2]) + b2[j].z);
          v[4 * j + 3] = (__uint_as_float(acc[4 * j + 3]) + b1[j].w) * gelu_exp2poly(__uint_as_float(gat[4 * j + 3]) + b2[j].w);
        }
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        if (MODE == EPI_GENERIC) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= g.alpha;
        }
        if (g.bias) {  // staged by epilogue_prefetch (zeros beyond N)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = lds128f(bias_s + (uint32_t)(ci * 64 + h * 32 + 4 * j) * 4u);
            v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
          }
        }
        if (rv) {
          if (oc + 32 <= g.N && (g.ldv & 3) == 0) {
            const float4* r4 = reinterpret_cast<const float4*>(rv + oc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(r4 + j);
              v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (oc + j < g.N) v[j] += __ldg(rv + oc + j);
          }
        }
        if (MODE == EPI_GENERIC && g.act) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = oc + j;
            v[j] = apply_act(v[j], g.act, (g.act == ACT_PRELU && col < g.N) ? __ldg(g.act_param + col) : 0.f);
          }
        }
      }
      if (MODE == EPI_GENERIC && g.ksplit) {
        // split-K: the raw fp32 accumulators of this K slice -> out32[z][M][N], through the (otherwise idle) staging
        // tile so that the stores are coalesced: a row of 32 floats is exactly one 128-byte staging row
#pragma unroll
        for (int u4 = 0; u4 < 8; ++u4)
          sts128(my_row + (uint32_t)((u4 ^ sw) << 4),
                 make_uint4(__float_as_uint(v[4 * u4]), __float_as_uint(v[4 * u4 + 1]), __float_as_uint(v[4 * u4 + 2]),
                            __float_as_uint(v[4 * u4 + 3])));
        __syncwarp();
        float* part = g.out32 + (long long)t.z * g.M * g.N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3), unit = lane & 7;
          const long long gm = t.m_base + row;
          const uint4 val = lds128(stage + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4));
          if (gm < g.M && oc + unit * 4 < g.N) *reinterpret_cast<uint4*>(part + gm * g.N + oc + unit * 4) = val;
        }
        __syncwarp();
        continue;
      }
      if (t.vec_ok) {
        // own row: add the staged residual, then overwrite the same 16-byte units with the fp16 result
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const uint32_t a = my_row + (uint32_t)(((h * 4 + u4) ^ sw) << 4);
          if (g.res) {
            const uint4 u = lds128(a);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_h2(w[e]);
              v[u4 * 8 + 2 * e] += f.x;
              v[u4 * 8 + 2 * e + 1] += f.y;
            }
          }
          if (MODE == EPI_GENERIC && g.relu_after_res) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u4 * 8 + e] = fmaxf(v[u4 * 8 + e], 0.f);
          }
          uint4 o;
          o.x = pack_h2(v[u4 * 8 + 0], v[u4 * 8 + 1]);
          o.y = pack_h2(v[u4 * 8 + 2], v[u4 * 8 + 3]);
          o.z = pack_h2(v[u4 * 8 + 4], v[u4 * 8 + 5]);
          o.w = pack_h2(v[u4 * 8 + 6], v[u4 * 8 + 7]);
          sts128(a, o);
        }
      } else if (row_ok) {
        // scalar fallback (odd widths / strides, fp32 strided outputs)
        if (g.res) {
          const __half* rp = g.res + t.zoff + (t.r_base + lane) * g.ldr + oc;
          for (int j = 0; j < 32; ++j)
            if (oc + j < t.NO) v[j] += __half2float(rp[j]);
        }
        if (MODE == EPI_GENERIC && g.relu_after_res) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (g.out) {
          __half* op = g.out + t.zoff + epi_out_row(g, m, t.z) * g.ldo + oc;
          for (int j = 0; j < 32; ++j)
            if (oc + j < t.NO) op[j] = __float2half_rn(v[j]);
        }
        if (MODE == EPI_GENERIC && g.out32) {
          float* op = g.out32 + (m / g.o32_rpn) * g.o32_sn + (m % g.o32_rpn) * g.o32_sp;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (oc + j < t.NO) op[(long long)(oc + j) * g.o32_sc] = v[j];
        }
      }
    }
    // staging tile -> HBM, coalesced: 8 lanes x 16 B per row, 4 rows per instruction
    if (t.vec_ok) {
      __syncwarp();
      uint4 val[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), unit = lane & 7;
        val[i] = lds128(stage + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), unit = lane & 7;
        const long long gm = t.m_base + row;
        if (gm < g.M && unit * 8 < cvalid)
          *reinterpret_cast<uint4*>(g.out + t.zoff + epi_out_row(g, gm, t.z) * g.ldo + ocol0 + unit * 8) = val[i];
      }
      if (MODE == EPI_FAST && g.stats) {
        // GroupNorm statistics of the tile this warp has just produced (the consumer's first read of the tensor is gone):
        // every lane holds 8 rows x 8 channels of the FINAL fp16 values; per-channel sum / sum of squares over those rows,
        // then over the 4 lanes that share the channel unit (fixed order: bitwise reproducible, batch independent).
        float cs[8], cq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] = cq[j] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3);
          if (t.m_base + row < g.M) {
            const uint32_t w[4] = {val[i].x, val[i].y, val[i].z, val[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_h2(w[e]);
              cs[2 * e] += f.x, cq[2 * e] = fmaf(f.x, f.x, cq[2 * e]);
              cs[2 * e + 1] += f.y, cq[2 * e + 1] = fmaf(f.y, f.y, cq[2 * e + 1]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
          cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 8);
          cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
          cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 16);
        }
        const int unit = lane & 7;
        if (lane < 8 && unit * 8 < cvalid) {
          const long long prow = g.up ? (t.m_base >> 5) * 4 + t.z : (t.m_base >> 5);
          float4* sp = reinterpret_cast<float4*>(g.stats + (prow * g.N + ocol0 + unit * 8) * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j) sp[j] = make_float4(cs[2 * j], cq[2 * j], cs[2 * j + 1], cq[2 * j + 1]);
        }
      }
    }
  }
  __syncwarp();  // staging tiles / bias strip may be refilled by the next tile's prefetch
}

}  // namespace rfb
