// tcgen05 / TMEM / TMA GEMM for sm_100a.  One kernel serves every dense contraction on REFace's
// inference path: nn.Linear / 1x1 conv (plain 2-D A), 3x3 stride-1 conv as implicit GEMM (A tiles are
// shifted NHWC boxes fetched by TMA, zero fill = padding), and the batched Q.K^T / P.V products.
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] (fp16, K-major, smem SW128)  x  B[BN x K]^T (fp16, K-major)
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = TMEM alloc + single-thread MMA issuer,
// warps2-5 = epilogue (tcgen05.ld -> bias/time-emb/activation/GEGLU/residual -> fp16 or fp32 stores).
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), one tmem_full mbarrier (MMA -> epilogue).
#pragma once
#include "ptx.cuh"

namespace rfb {

enum AMode { A_PLAIN = 0, A_CONV3 = 1, A_BATCH3 = 2, A_HEADS4 = 3 };
enum BMode { B_PLAIN = 0, B_BATCH3 = 2, B_HEADS4 = 3 };
enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_QGELU = 3, ACT_PRELU = 4, ACT_RELU = 5, ACT_SIGMOID = 6 };

struct GemmArgs {
  int M, N, nk;  // rows (per batch z), valid output columns (of the B tensor), number of 64-wide k blocks
  int BN, stages, tmem_cols, kmerge;
  int a_mode, b_mode;
  // A_CONV3 geometry: tile = bimg images x bh rows x bw cols (=128 output pixels), K = taps x cblocks x 64
  int cblocks, bw, bh, bimg, tiles_w, tiles_h;
  int cstride, cpad_l, cpad_t;  // conv stride (1 or 2) and left / top padding: input coordinate = stride * out + tap - pad
  int heads;  // *_HEADS4: z = n*heads + head
  // epilogue
  float alpha;
  const float* bias;    // [N] (packed order for GEGLU) or nullptr
  const float* rowvec;  // per-sample vector added per column: rowvec[(m / rows_per_vec) * ldv + col]
  int rows_per_vec, ldv;
  const float* act_param;  // PReLU slopes [N]
  int act, geglu;
  int relu_after_res;  // ReLU applied AFTER the residual add (ResNet BasicBlock: relu(shortcut + bn(conv)))
  const __half* res;  // residual [M, ldr] added after activation
  long long ldr;
  __half* out;  // fp16 [M, ldo]
  long long ldo;
  float* out32;  // optional fp32 strided output: (m / o32_rpn) * o32_sn + (m % o32_rpn) * o32_sp + col * o32_sc
  long long o32_sn, o32_sp, o32_sc;
  int o32_rpn;
  // split-K: blockIdx.z = K slice; k-block index = z * nk + kb; raw fp32 partial tiles go to out32[z][M][N]
  int ksplit;
  // batch (blockIdx.z) offsets in elements for out/res: (z / zdiv) * zs_outer + (z % zdiv) * zs_inner
  long long zs_outer, zs_inner;
  int zdiv;
  // optional instrumentation (option "gemm_debug"): per CTA 8 x u64 clock64 totals
  // [0] MMA thread total, [1] waiting on full[], [2] waiting on acc_empty[], [3] producer waiting on empty[],
  // [4] epilogue warp2 waiting on acc_full[], [5] epilogue warp2 total, [6] k-stages issued, [7] tiles
  unsigned long long* dbg;
};

static constexpr int GEMM_THREADS = 192;
static constexpr int GEMM_BM = 128;
static constexpr int GEMM_BK = 64;
static constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KiB

__host__ __device__ inline size_t gemm_smem_bytes(int stages, int BN) {
  return 1024 + (size_t)stages * (GEMM_A_STAGE_BYTES + (size_t)BN * 128) + 16 * stages + 64;
}

// 1 / (1 + 2^(-k x)) in four instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP); the ftz approximations need no range
// fix-ups (2^.. -> inf gives 0, -> 0 gives 1) and are accurate to ~3 ulp, far below the fp16 output.  `x / (1 + __expf(-x))`
// compiled to an IEEE division (~25 instructions per element in an epilogue that is bound by instruction issue).
__device__ __forceinline__ float fast_sigmoid_scaled(float x, float k_log2e) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -k_log2e));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float apply_act(float v, int act, float p) {
  switch (act) {
    case ACT_SILU: return v * fast_sigmoid_scaled(v, 1.4426950408889634f);
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    case ACT_QGELU: return v * fast_sigmoid_scaled(v, 1.702f * 1.4426950408889634f);
    case ACT_PRELU: return v >= 0.f ? v : v * p;
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: return fast_sigmoid_scaled(v, 1.4426950408889634f);
    default: return v;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = g.stages;
  const int BN = g.BN;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)S * GEMM_A_STAGE_BYTES;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128u;
  const uint32_t bars = sB + (uint32_t)S * b_stage_bytes;
  const uint32_t bar_tfull = bars + 16u * S;
  const uint32_t tptr = bar_tfull + 8u;

  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int z = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bars + 8u * i, 1);        // full[i]
      mbar_init(bars + 8u * (S + i), 1);  // empty[i]
    }
    mbar_init(bar_tfull, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tptr, (uint32_t)g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int cw = 0, ch = 0, cn = 0;
      if (g.a_mode == A_CONV3) {
        if (g.bimg > 1) {
          cn = m_tile * g.bimg;
        } else {
          const int per_img = g.tiles_w * g.tiles_h;
          cn = m_tile / per_img;
          const int rem = m_tile - cn * per_img;
          ch = (rem / g.tiles_w) * g.bh;
          cw = (rem % g.tiles_w) * g.bw;
        }
      }
      const int m0 = m_tile * GEMM_BM;
      const int n0 = n_tile * BN;
      const uint32_t tx = GEMM_A_STAGE_BYTES + b_stage_bytes;
      for (int kb = 0; kb < g.nk; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(bars + 8u * (S + s), ph ^ 1u);
        const uint32_t full = bars + 8u * s;
        mbar_expect_tx(full, tx);
        const uint32_t dA = sA + (uint32_t)s * GEMM_A_STAGE_BYTES;
        const uint32_t dB = sB + (uint32_t)s * b_stage_bytes;
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
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      const uint32_t idesc = idesc_f16(GEMM_BM, (uint32_t)BN);
      for (int kb = 0; kb < g.nk; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(bars + 8u * s, ph);
        tc_fence_after();
        const uint64_t da = smem_desc_k_sw128(sA + (uint32_t)s * GEMM_A_STAGE_BYTES);
        const uint64_t db = smem_desc_k_sw128(sB + (uint32_t)s * b_stage_bytes);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advance 16 halfs = 32 B along K inside the 128 B swizzle atom: +2 in the (addr>>4) field
          mma_f16_ss(tmem_base, da + 2u * k, db + 2u * k, idesc, (uint32_t)((kb | k) != 0));
        }
        mma_commit(bars + 8u * (S + s));  // frees the smem slot when these MMAs retire
      }
      mma_commit(bar_tfull);
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may touch
    const int r = q * 32 + lane;
    const long long m = (long long)m_tile * GEMM_BM + r;
    const bool row_ok = m < g.M;
    const long long zoff = (long long)(z / g.zdiv) * g.zs_outer + (long long)(z % g.zdiv) * g.zs_inner;
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int halfN = BN >> 1;
    const int ncols = g.geglu ? halfN : BN;
    const float* rv = (g.rowvec && row_ok) ? g.rowvec + (m / g.rows_per_vec) * g.ldv : nullptr;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t acc[32];
      float v[32];
      tmem_ld32(trow + (uint32_t)c0, acc);
      if (g.geglu) {
        uint32_t gat[32];
        tmem_ld32(trow + (uint32_t)(halfN + c0), gat);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int cx = n_tile * BN + c0 + j, cg = cx + halfN;
          float xv = __uint_as_float(acc[j]) * g.alpha;
          float gv = __uint_as_float(gat[j]) * g.alpha;
          if (g.bias) {
            xv += (cx < g.N) ? __ldg(g.bias + cx) : 0.f;
            gv += (cg < g.N) ? __ldg(g.bias + cg) : 0.f;
          }
          v[j] = xv * apply_act(gv, ACT_GELU, 0.f);
        }
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n_tile * BN + c0 + j;
          float x = __uint_as_float(acc[j]) * g.alpha;
          if (col < g.N) {
            if (g.bias) x += __ldg(g.bias + col);
            if (rv) x += __ldg(rv + col);
            if (g.act) x = apply_act(x, g.act, g.act == ACT_PRELU ? __ldg(g.act_param + col) : 0.f);
          }
          v[j] = x;
        }
      }
      if (!row_ok) continue;
      const int ocol0 = (g.geglu ? n_tile * halfN : n_tile * BN) + c0;
      const int NO = g.geglu ? (g.N >> 1) : g.N;
      if (g.res) {
        const __half* rp = g.res + zoff + m * g.ldr + ocol0;
#pragma unroll
        for (int grp = 0; grp < 4; ++grp) {
          if (ocol0 + grp * 8 + 8 <= NO) {
            const uint4 u = *reinterpret_cast<const uint4*>(rp + grp * 8);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = unpack_h2(w[t]);
              v[grp * 8 + 2 * t] += f.x;
              v[grp * 8 + 2 * t + 1] += f.y;
            }
          } else {
            for (int j = 0; j < 8; ++j)
              if (ocol0 + grp * 8 + j < NO) v[grp * 8 + j] += __half2float(rp[grp * 8 + j]);
          }
        }
      }
      if (g.out) {
        __half* op = g.out + zoff + m * g.ldo + ocol0;
#pragma unroll
        for (int grp = 0; grp < 4; ++grp) {
          if (ocol0 + grp * 8 + 8 <= NO) {
            uint4 u;
            u.x = pack_h2(v[grp * 8 + 0], v[grp * 8 + 1]);
            u.y = pack_h2(v[grp * 8 + 2], v[grp * 8 + 3]);
            u.z = pack_h2(v[grp * 8 + 4], v[grp * 8 + 5]);
            u.w = pack_h2(v[grp * 8 + 6], v[grp * 8 + 7]);
            *reinterpret_cast<uint4*>(op + grp * 8) = u;
          } else {
            for (int j = 0; j < 8; ++j)
              if (ocol0 + grp * 8 + j < NO) op[grp * 8 + j] = __float2half_rn(v[grp * 8 + j]);
          }
        }
      }
      if (g.out32) {
        float* op = g.out32 + (m / g.o32_rpn) * g.o32_sn + (m % g.o32_rpn) * g.o32_sp;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (ocol0 + j < NO) op[(long long)(ocol0 + j) * g.o32_sc] = v[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

}  // namespace rfb
