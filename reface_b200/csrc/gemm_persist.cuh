// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution (second generation of gemm_tc_kernel).
//
// One CTA per SM loops over output tiles (static round-robin, N-tiles fastest so that CTAs running at the same
// time share the A tile through L2).  320 threads:
//   warp0    TMA producer  (smem ring of `stages` x {A 128x64, B BNx64} fp16 tiles, 128B swizzle)
//   warp1    TMEM owner + single-thread tcgen05.mma issuer; TWO accumulator stages of BN fp32 columns each, so the
//            MMAs of tile i+1 run while tile i is being drained
//   warps2-9 epilogue: two warps per TMEM lane quadrant, each draining every other 32-column chunk
//            (tcgen05.ld -> bias / time-embedding row vector / GEGLU / activation / residual -> 16-byte stores)
// The first-generation kernel spent 3-10x the MMA time in a 4-warp, fully predicated epilogue that could not
// overlap with the main loop; here full tiles take a branch-free path with 16-byte bias/residual traffic.
#pragma once
#include "gemm_tc.cuh"

namespace rfb {

static constexpr int GEMMP_THREADS = 320;
static constexpr int GEMMP_EPI_WARPS = 8;

__host__ __device__ inline size_t gemmp_smem_bytes(int stages, int BN) {
  return 1024 + (size_t)stages * (GEMM_A_STAGE_BYTES + (size_t)BN * 128) + 16 * stages + 128;
}

enum EpiMode { EPI_FAST = 0, EPI_GEGLU = 1, EPI_GENERIC = 2 };

// exact-erf GELU with a cheap erf (Abramowitz-Stegun 7.1.26, |err| < 1.5e-7: far below the fp16 output ulp)
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

template <int MODE>
__global__ void __launch_bounds__(GEMMP_THREADS, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g,
                    const int m_tiles, const int n_tiles, const int total_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform (see gemm_pair.cuh)
  const int lane = threadIdx.x & 31;
  const int S = g.stages;
  const int BN = g.BN;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)S * GEMM_A_STAGE_BYTES;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128u;
  const uint32_t bars = sB + (uint32_t)S * b_stage_bytes;  // full[S], empty[S]
  const uint32_t bar_accf = bars + 16u * S;                // acc_full[2]
  const uint32_t bar_acce = bar_accf + 16u;                // acc_empty[2]
  const uint32_t tptr = bar_acce + 16u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bars + 8u * i, 1);
      mbar_init(bars + 8u * (S + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accf + 8u * i, 1);
      mbar_init(bar_acce + 8u * i, GEMMP_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const int per_z = m_tiles * n_tiles;

  if (warp == 0) {
    {
      // ------------------------------------------------------------ TMA producer (converged warp, elected issue)
      const uint32_t tx = GEMM_A_STAGE_BYTES + b_stage_bytes;
      uint32_t st = 0, sp = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / per_z;
        const int rem = tile - z * per_z;
        const int m_tile = rem / n_tiles, n_tile = rem - m_tile * n_tiles;
        int cw = 0, ch = 0, cn = 0;
        if (g.a_mode == A_CONV3) {
          if (g.bimg > 1) {
            cn = m_tile * g.bimg;
          } else {
            const int per_img = g.tiles_w * g.tiles_h;
            cn = m_tile / per_img;
            const int r2 = m_tile - cn * per_img;
            ch = (r2 / g.tiles_w) * g.bh;
            cw = (r2 % g.tiles_w) * g.bw;
          }
        }
        const int m0 = m_tile * GEMM_BM, n0 = n_tile * BN;
        for (int kb = 0; kb < g.nk; ++kb) {
          const uint32_t s = st, ph = sp;
          if (++st == (uint32_t)S) st = 0, sp ^= 1u;
          mbar_wait(bars + 8u * (S + s), ph ^ 1u);
          const uint32_t full = bars + 8u * s;
          if (elect_one()) {
            mbar_expect_tx(full, tx);
            const uint32_t dA = sA + s * GEMM_A_STAGE_BYTES;
            const uint32_t dB = sB + s * b_stage_bytes;
            switch (g.a_mode) {
              case A_PLAIN: tma_load_2d(dA, &tmA, full, kb * GEMM_BK, m0); break;
              case A_CONV3: {
                const int tap = kb / g.cblocks;
                const int cb = kb - tap * g.cblocks;
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                tma_load_4d(dA, &tmA, full, cb * GEMM_BK, cw + dx, ch + dy, cn);
              } break;
              case A_BATCH3: tma_load_3d(dA, &tmA, full, kb * GEMM_BK, m0, z); break;
              default: tma_load_4d(dA, &tmA, full, kb * GEMM_BK, z % g.heads, m0, z / g.heads); break;
            }
            switch (g.b_mode) {
              case B_PLAIN: tma_load_2d(dB, &tmB, full, kb * GEMM_BK, n0); break;
              case B_BATCH3: tma_load_3d(dB, &tmB, full, kb * GEMM_BK, n0, z); break;
              default: tma_load_4d(dB, &tmB, full, kb * GEMM_BK, z % g.heads, n0, z / g.heads); break;
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      // ------------------------------------------------------------ MMA issuer (converged warp, elected issue)
      const uint32_t idesc = idesc_f16(GEMM_BM, (uint32_t)BN);
      uint32_t ti = 0, st = 0, sp = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait(bar_acce + 8u * as, aph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * 256u;
        for (int kb = 0; kb < g.nk; ++kb) {
          const uint32_t s = st, ph = sp;
          if (++st == (uint32_t)S) st = 0, sp ^= 1u;
          mbar_wait(bars + 8u * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = smem_desc_k_sw128(sA + s * GEMM_A_STAGE_BYTES);
            const uint64_t db = smem_desc_k_sw128(sB + s * b_stage_bytes);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              mma_f16_ss(tacc, da + 2u * k, db + 2u * k, idesc, (uint32_t)((kb | k) != 0));
            mma_commit(bars + 8u * (S + s));
            if (kb == g.nk - 1) mma_commit(bar_accf + 8u * as);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..9)
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant
    const int half = e >> 2;  // which 32-column chunks this warp drains
    const int r = q * 32 + lane;
    const int halfN = BN >> 1;
    const int ncols = (MODE == EPI_GEGLU) ? halfN : BN;
    const int NO = (MODE == EPI_GEGLU) ? (g.N >> 1) : g.N;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      const int z = tile / per_z;
      const int rem = tile - z * per_z;
      const int m_tile = rem / n_tiles, n_tile = rem - m_tile * n_tiles;
      const long long m = (long long)m_tile * GEMM_BM + r;
      const bool row_ok = m < g.M;
      const long long zoff = (long long)(z / g.zdiv) * g.zs_outer + (long long)(z % g.zdiv) * g.zs_inner;
      const float* rv = (g.rowvec && row_ok) ? g.rowvec + (m / g.rows_per_vec) * g.ldv : nullptr;
      mbar_wait(bar_accf + 8u * as, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + as * 256u + ((uint32_t)(q * 32) << 16);
      for (int c0 = half * 32; c0 < ncols; c0 += 64) {
        const int ocol0 = ((MODE == EPI_GEGLU) ? n_tile * halfN : n_tile * BN) + c0;
        uint32_t acc[32];
        float v[32];
        tmem_ld32(trow + (uint32_t)c0, acc);
        if (MODE == EPI_GEGLU) {
          uint32_t gat[32];
          tmem_ld32(trow + (uint32_t)(halfN + c0), gat);
          tmem_ld_wait();
          const float4* bx = reinterpret_cast<const float4*>(g.bias + n_tile * BN + c0);
          const float4* bg = reinterpret_cast<const float4*>(g.bias + n_tile * BN + halfN + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b1 = __ldg(bx + j), b2 = __ldg(bg + j);
            v[4 * j + 0] = (__uint_as_float(acc[4 * j + 0]) + b1.x) * gelu_erf_fast(__uint_as_float(gat[4 * j + 0]) + b2.x);
            v[4 * j + 1] = (__uint_as_float(acc[4 * j + 1]) + b1.y) * gelu_erf_fast(__uint_as_float(gat[4 * j + 1]) + b2.y);
            v[4 * j + 2] = (__uint_as_float(acc[4 * j + 2]) + b1.z) * gelu_erf_fast(__uint_as_float(gat[4 * j + 2]) + b2.z);
            v[4 * j + 3] = (__uint_as_float(acc[4 * j + 3]) + b1.w) * gelu_erf_fast(__uint_as_float(gat[4 * j + 3]) + b2.w);
          }
        } else {
          tmem_ld_wait();
          const bool fullc = (ocol0 + 32 <= NO);
          if (MODE == EPI_FAST && fullc) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
            if (g.bias) {
              const float4* b4 = reinterpret_cast<const float4*>(g.bias + ocol0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
              }
            }
            if (rv) {
              const float4* r4 = reinterpret_cast<const float4*>(rv + ocol0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(r4 + j);
                v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = ocol0 + j;
              float x = __uint_as_float(acc[j]) * g.alpha;
              if (col < g.N) {
                if (g.bias) x += __ldg(g.bias + col);
                if (rv) x += __ldg(rv + col);
                if (g.act) x = apply_act(x, g.act, g.act == ACT_PRELU ? __ldg(g.act_param + col) : 0.f);
              }
              v[j] = x;
            }
          }
        }
        if (!row_ok) continue;
        const bool fullc = (ocol0 + 32 <= NO);
        if (g.res) {
          const __half* rp = g.res + zoff + m * g.ldr + ocol0;
          if (fullc) {
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {
              const uint4 u = *reinterpret_cast<const uint4*>(rp + grp * 8);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_h2(w[t]);
                v[grp * 8 + 2 * t] += f.x;
                v[grp * 8 + 2 * t + 1] += f.y;
              }
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (ocol0 + j < NO) v[j] += __half2float(rp[j]);
          }
        }
        if (g.out) {
          __half* op = g.out + zoff + m * g.ldo + ocol0;
          if (fullc) {
#pragma unroll
            for (int grp = 0; grp < 4; ++grp) {
              uint4 u;
              u.x = pack_h2(v[grp * 8 + 0], v[grp * 8 + 1]);
              u.y = pack_h2(v[grp * 8 + 2], v[grp * 8 + 3]);
              u.z = pack_h2(v[grp * 8 + 4], v[grp * 8 + 5]);
              u.w = pack_h2(v[grp * 8 + 6], v[grp * 8 + 7]);
              *reinterpret_cast<uint4*>(op + grp * 8) = u;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (ocol0 + j < NO) op[j] = __float2half_rn(v[j]);
          }
        }
        if (MODE == EPI_GENERIC && g.out32) {
          float* op = g.out32 + (m / g.o32_rpn) * g.o32_sn + (m % g.o32_rpn) * g.o32_sp;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ocol0 + j < NO) op[(long long)(ocol0 + j) * g.o32_sc] = v[j];
        }
      }
      // this warp has finished reading the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8u * as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace rfb
