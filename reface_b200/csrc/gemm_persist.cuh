// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution, one CTA per SM (used for the batched /
// per-head operand modes and for problems too small for CTA pairs; the 2-CTA variant lives in gemm_pair.cuh).
//
// One CTA per SM loops over output tiles (static round-robin, N-tiles fastest so that CTAs running at the same
// time share the A tile through L2).  320 threads:
//   warp0    TMA producer  (smem ring of `stages` x {A 128x64, B BNx64} fp16 tiles, 128B swizzle)
//   warp1    TMEM owner + tcgen05.mma issuer; TWO accumulator stages of BN fp32 columns each, so the MMAs of tile
//            i+1 run while tile i is being drained
//   warps2-9 epilogue (gemm_epilogue.cuh): two warps per TMEM lane quadrant, 64-column chunks, coalesced HBM traffic
// The TMA and MMA roles run as CONVERGED warps with one elected lane issuing: their operands are then warp-uniform
// and live in uniform registers; a lane-0-only loop made ptxas emit ELECT + R2UR.BROADCAST chains before every
// UTCHMMA/UTMALDG (~130 cycles per MMA issue).
#pragma once
#include "gemm_epilogue.cuh"

namespace rfb {

static constexpr int GEMMP_THREADS = 320;
static constexpr int GEMMP_EPI_WARPS = 8;

// NP = epilogue warps per TMEM lane quadrant (2 for MMA-bound shapes; 3 for short-K, epilogue-bound GEMMs whose
// per-warp latency chains need more warps in flight)
__host__ __device__ inline size_t gemmp_smem_bytes(int stages, int BN, int np = 2) {
  return 2048 + (size_t)stages * (GEMM_A_STAGE_BYTES + (size_t)BN * 128) + 16 * stages + 128 +
         (size_t)(4 * np) * EPI_WARP_BYTES;
}

// (z, m_tile, n_tile) of the tiles a CTA visits (tile = blockIdx.x + i * gridDim.x), advanced with adds and compares:
// decomposing the linear tile index with runtime integer divisions cost ~280 instructions per epilogue warp and
// tile (28 % of everything the K=320 GEMMs executed, profiles/r01s2_gemm_k320_ncu_full.txt).
struct TileIter {
  int z, m_tile, n_tile, step_z, step_m, step_n;
  __device__ __forceinline__ void init(int tile, int stride, int m_tiles, int n_tiles) {
    const int per_z = m_tiles * n_tiles;
    z = tile / per_z;
    int rem = tile - z * per_z;
    m_tile = rem / n_tiles, n_tile = rem - m_tile * n_tiles;
    step_z = stride / per_z;
    rem = stride - step_z * per_z;
    step_m = rem / n_tiles, step_n = rem - step_m * n_tiles;
  }
  __device__ __forceinline__ void next(int m_tiles, int n_tiles) {
    n_tile += step_n, m_tile += step_m, z += step_z;
    if (n_tile >= n_tiles) n_tile -= n_tiles, ++m_tile;
    if (m_tile >= m_tiles) m_tile -= m_tiles, ++z;
  }
};

template <int MODE, int NP>
__global__ void __launch_bounds__(64 + NP * 128, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const GemmArgs g, const int m_tiles, const int n_tiles,
                    const int total_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
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
  const uint32_t epi_stage = (tptr + 16u + 1023u) & ~1023u;  // 8 x 4 KB staging tiles

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bars + 8u * i, 1);
      mbar_init(bars + 8u * (S + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accf + 8u * i, 1);
      mbar_init(bar_acce + 8u * i, 4 * NP);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.nk1) tma_prefetch_desc(&tmA2);
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
    // ------------------------------------------------------------ TMA producer (converged warp, elected issue)
    const uint32_t tx = GEMM_A_STAGE_BYTES + b_stage_bytes;
    uint32_t st = 0, sp = 0;
    long long t_dbg = 0;  // option gemm_debug: cycles this role spent waiting (see GemmArgs::dbg)
    TileIter it;
    it.init(blockIdx.x, gridDim.x, m_tiles, n_tiles);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it.next(m_tiles, n_tiles)) {
      const int z = it.z, m_tile = it.m_tile, n_tile = it.n_tile;
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
      const int m0b = g.a2_mod ? m0 % g.a2_mod : m0;  // row of the second A source (tiles never straddle its end)
      const int py = z >> 1, px = z & 1;              // output phase of the folded upsample convolution
      for (int kb = 0; kb < g.nk; ++kb) {
        const uint32_t s = st, ph = sp;
        if (++st == (uint32_t)S) st = 0, sp ^= 1u;
        const long long tw0 = g.dbg ? clock64() : 0;
        mbar_wait(bars + 8u * (S + s), ph ^ 1u);
        if (g.dbg) t_dbg += clock64() - tw0;
        const uint32_t full = bars + 8u * s;
        if (elect_one()) {
          mbar_expect_tx(full, tx);
          const uint32_t dA = sA + s * GEMM_A_STAGE_BYTES;
          const uint32_t dB = sB + s * b_stage_bytes;
          const int kg = g.ksplit ? z * g.nk + kb : kb;  // split-K: this unit's slice of the K range
          switch (g.a_mode) {
            case A_PLAIN:
              if (g.nk1 && kg >= g.nk1) tma_load_2d(dA, &tmA2, full, (kg - g.nk1) * GEMM_BK, m0b);
              else tma_load_2d(dA, &tmA, full, kg * GEMM_BK, m0);
              break;
            case A_CONV3: {
              const int tap = kg / g.cblocks;
              const int cb = kg - tap * g.cblocks;
              int dy, dx;
              if (g.up) dy = (tap >> 1) - 1 + py, dx = (tap & 1) - 1 + px;
              else dy = tap / 3 - g.cpad_t, dx = tap % 3 - g.cpad_l;
              tma_load_4d(dA, &tmA, full, cb * GEMM_BK, cw * g.cstride + dx, ch * g.cstride + dy, cn);
            } break;
            case A_BATCH3: tma_load_3d(dA, &tmA, full, kb * GEMM_BK, m0, z); break;
            default: tma_load_4d(dA, &tmA, full, kb * GEMM_BK, z % g.heads, m0, z / g.heads); break;
          }
          switch (g.b_mode) {
            case B_PLAIN: tma_load_2d(dB, &tmB, full, kg * GEMM_BK, n0); break;
            case B_BATCH3: tma_load_3d(dB, &tmB, full, kb * GEMM_BK, n0, z); break;
            default: tma_load_4d(dB, &tmB, full, kb * GEMM_BK, z % g.heads, n0, z / g.heads); break;
          }
        }
        __syncwarp();
      }
    }
    if (g.dbg && lane == 0) g.dbg[(size_t)blockIdx.x * 8 + 3] = (unsigned long long)t_dbg;
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, elected issue)
    const uint32_t idesc = idesc_f16(GEMM_BM, (uint32_t)BN);
    uint32_t ti = 0, st = 0, sp = 0;
    long long t_full = 0, t_acc = 0;
    const long long t_begin = g.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      long long tw0 = g.dbg ? clock64() : 0;
      mbar_wait(bar_acce + 8u * as, aph ^ 1u);  // epilogue has drained this accumulator stage
      if (g.dbg) t_acc += clock64() - tw0;
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * 256u;
      for (int kb = 0; kb < g.nk; ++kb) {
        const uint32_t s = st, ph = sp;
        if (++st == (uint32_t)S) st = 0, sp ^= 1u;
        tw0 = g.dbg ? clock64() : 0;
        mbar_wait(bars + 8u * s, ph);
        if (g.dbg) t_full += clock64() - tw0;
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
    pdl_trigger();  // this CTA's MMAs are all issued: the next kernel may be scheduled behind the grid's last epilogues
    if (g.dbg && lane == 0) {
      unsigned long long* d = g.dbg + (size_t)blockIdx.x * 8;
      d[0] = (unsigned long long)(clock64() - t_begin), d[1] = (unsigned long long)t_full, d[2] = (unsigned long long)t_acc, d[7] = ti;
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..9)
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant
    const int half = e >> 2;  // which 64-column chunks this warp drains
    const uint32_t stage = epi_stage + (uint32_t)e * EPI_WARP_BYTES;  // this warp's staging tile + bias strip
    uint32_t ti = 0;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    long long t_wait = 0, t_pre = 0;
    const long long t_ebegin = g.dbg ? clock64() : 0;
    TileIter it;
    it.init(blockIdx.x, gridDim.x, m_tiles, n_tiles);
    if ((int)blockIdx.x < total_tiles) {
      const EpiTile e0 = epi_tile_info<MODE>(g, q, it.m_tile, it.n_tile, it.z);
      epilogue_lookahead<MODE, NP>(g, e0, lane, half, nb);
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      const int n_tile = it.n_tile;
      const EpiTile et = epi_tile_info<MODE>(g, q, it.m_tile, it.n_tile, it.z);
      const long long tp0 = g.dbg ? clock64() : 0;
      epilogue_prefetch<MODE, NP>(g, et, stage, lane, half, n_tile, nb);  // bias / residual while the MMAs still run
      it.next(m_tiles, n_tiles);
      if (tile + (int)gridDim.x < total_tiles) {  // next tile's bias -> registers, residual lines -> L2
        const EpiTile en = epi_tile_info<MODE>(g, q, it.m_tile, it.n_tile, it.z);
        epilogue_lookahead<MODE, NP>(g, en, lane, half, nb);
      }
      const long long tw0 = g.dbg ? clock64() : 0;
      mbar_wait(bar_accf + 8u * as, aph);
      if (g.dbg) t_wait += clock64() - tw0, t_pre += tw0 - tp0;
      tc_fence_after();
      const uint32_t trow = tmem_base + as * 256u + ((uint32_t)(q * 32) << 16);
      epilogue_drain<MODE, NP>(g, et, trow, stage, lane, half, n_tile);
      // this warp has finished reading the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8u * as);
    }
    if (g.dbg && warp == 2 && lane == 0) {
      unsigned long long* d = g.dbg + (size_t)blockIdx.x * 8;
      d[4] = (unsigned long long)t_wait, d[5] = (unsigned long long)(clock64() - t_ebegin), d[6] = (unsigned long long)t_pre;
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
