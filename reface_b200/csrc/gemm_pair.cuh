// 2-CTA (tcgen05 cta_group::2) persistent GEMM / implicit-GEMM convolution -- the main tensor-core kernel.
//
// A CTA PAIR (cluster of 2 on one TPC) computes a 256 x BN tile with one UMMA (M=256): each CTA loads its own 128 A
// rows and only HALF of the B tile (BN/2 rows); the tensor cores read both halves across the pair, so operand bytes
// per FLOP drop by 27-33 % versus the 1-CTA kernel (which ncu showed pinned at the ~6-7 KB/cycle L2->SM fabric limit).
//
// Protocol (per CTA: warp0 TMA producer, warp1 TMEM owner + MMA issuer (leader only), warps2-9 epilogue):
//   full[s]      leader's barrier only (count 1): leader posts expect_tx for BOTH CTAs' bytes; both CTAs' TMA loads
//                complete_tx on it (cta_group::2 loads with the peer bit of the mbarrier address cleared)
//   empty[s]     one per CTA: released by the leader's tcgen05.commit (multicast to both CTAs)
//   acc_full[a]  one per CTA: multicast commit after the last k-block of a tile
//   acc_empty[a] leader's barrier (count 2 x 8 epilogue warps): the peer's warps arrive remotely (mapa)
// TMA / MMA roles are converged warps with one elected issuing lane (operands in uniform registers, see
// gemm_persist.cuh).  kmerge = number of 64-wide k-blocks per pipeline stage.
#pragma once
#include "gemm_persist.cuh"

namespace rfb {

__host__ __device__ inline size_t gemm2_smem_bytes(int stages, int BN, int kmerge) {
  return 2048 + (size_t)stages * kmerge * (GEMM_A_STAGE_BYTES + (size_t)(BN / 2) * 128) + 16 * stages + 128 +
         (size_t)GEMMP_EPI_WARPS * EPI_WARP_BYTES;
}

template <int MODE>
__global__ void __launch_bounds__(GEMMP_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g,
                 const int m_pairs, const int n_tiles, const int total_pairs) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int S = g.stages;
  const int BN = g.BN;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const int KM = g.kmerge;
  const uint32_t a_stage_bytes = (uint32_t)KM * GEMM_A_STAGE_BYTES;
  const uint32_t b_sub_bytes = (uint32_t)(BN / 2) * 128u;
  const uint32_t b_stage_bytes = (uint32_t)KM * b_sub_bytes;
  const uint32_t sB = base + (uint32_t)S * a_stage_bytes;
  const uint32_t bars = sB + (uint32_t)S * b_stage_bytes;  // full[S], empty[S]
  const uint32_t bar_accf = bars + 16u * S;
  const uint32_t bar_acce = bar_accf + 16u;
  const uint32_t tptr = bar_acce + 16u;
  const uint32_t epi_stage = (tptr + 16u + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bars + 8u * i, 1);
      mbar_init(bars + 8u * (S + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accf + 8u * i, 1);
      mbar_init(bar_acce + 8u * i, 2 * GEMMP_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  cluster_sync_all();  // barriers of both CTAs initialised before anyone signals them
  if (warp == 1) tmem_alloc_2cta(tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int per_z = m_pairs * n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    uint32_t st = 0, sp = 0;
    long long t_empty = 0, t_issue = 0;
    for (int pt = pair_id; pt < total_pairs; pt += num_pairs) {
      const int z = pt / per_z;
      const int rem = pt - z * per_z;
      const int mp = rem / n_tiles, n_tile = rem - mp * n_tiles;
      const int m_tile = 2 * mp + (int)rank;
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
      const int m0 = m_tile * GEMM_BM;
      const int n0 = n_tile * BN + (int)rank * (BN / 2);
      for (int kb0 = 0; kb0 < g.nk; kb0 += KM) {
        const uint32_t s = st, ph = sp;  // running stage / phase (no runtime division in the hot loop)
        if (++st == (uint32_t)S) st = 0, sp ^= 1u;
        const int nv = min(KM, g.nk - kb0);
        const long long t0 = g.dbg ? clock64() : 0;
        mbar_wait(bars + 8u * (S + s), ph ^ 1u);
        const long long ti0 = g.dbg ? clock64() : 0;
        if (g.dbg) t_empty += ti0 - t0;
        const uint32_t full = bars + 8u * s;
        if (elect_one()) {
          if (leader) mbar_expect_tx(full, 2u * (uint32_t)nv * (GEMM_A_STAGE_BYTES + b_sub_bytes));
          for (int j = 0; j < nv; ++j) {
            const int kb = kb0 + j;
            const uint32_t dA = sA + s * a_stage_bytes + (uint32_t)j * GEMM_A_STAGE_BYTES;
            const uint32_t dB = sB + s * b_stage_bytes + (uint32_t)j * b_sub_bytes;
            if (g.a_mode == A_PLAIN) {
              tma_load_2d_2sm(dA, &tmA, full, kb * GEMM_BK, m0);
            } else {
              const int tap = kb / g.cblocks;
              const int cb = kb - tap * g.cblocks;
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              tma_load_4d_2sm(dA, &tmA, full, cb * GEMM_BK, cw + dx, ch + dy, cn);
            }
            tma_load_2d_2sm(dB, &tmB, full, kb * GEMM_BK, n0);
          }
        }
        __syncwarp();
        if (g.dbg) t_issue += clock64() - ti0;
      }
    }
    if (g.dbg && lane == 0) {
      g.dbg[(size_t)blockIdx.x * 8 + 3] = (unsigned long long)t_empty;
      g.dbg[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)t_issue;
    }
  } else if (warp == 1) {
    if (leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA)
      const uint32_t idesc = idesc_f16(256, (uint32_t)BN);
      uint32_t it = 0, ti = 0, st = 0, sp = 0;
      long long t_full = 0, t_acc = 0;
      const long long t_begin = clock64();
      for (int pt = pair_id; pt < total_pairs; pt += num_pairs, ++ti) {
        const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
        long long t0 = g.dbg ? clock64() : 0;
        mbar_wait(bar_acce + 8u * as, aph ^ 1u);
        if (g.dbg) t_acc += clock64() - t0;
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * 256u;
        for (int kb0 = 0; kb0 < g.nk; kb0 += KM, ++it) {
          const uint32_t s = st, ph = sp;
          if (++st == (uint32_t)S) st = 0, sp ^= 1u;
          const int nv = min(KM, g.nk - kb0);
          t0 = g.dbg ? clock64() : 0;
          mbar_wait(bars + 8u * s, ph);
          if (g.dbg) t_full += clock64() - t0;
          tc_fence_after();
          const bool last = kb0 + KM >= g.nk;
          if (elect_one()) {
            for (int j = 0; j < nv; ++j) {
              const uint64_t da = smem_desc_k_sw128(sA + s * a_stage_bytes + (uint32_t)j * GEMM_A_STAGE_BYTES);
              const uint64_t db = smem_desc_k_sw128(sB + s * b_stage_bytes + (uint32_t)j * b_sub_bytes);
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k)
                mma_f16_ss_2cta(tacc, da + 2u * k, db + 2u * k, idesc, (uint32_t)(((kb0 + j) | k) != 0));
            }
            mma_commit_2cta_mc(bars + 8u * (S + s), 3);           // smem slot free in both CTAs
            if (last) mma_commit_2cta_mc(bar_accf + 8u * as, 3);  // accumulators ready in both CTAs
          }
          __syncwarp();
        }
      }
      if (g.dbg && lane == 0) {
        unsigned long long* d = g.dbg + (size_t)blockIdx.x * 8;
        d[0] = (unsigned long long)(clock64() - t_begin), d[1] = (unsigned long long)t_full;
        d[2] = (unsigned long long)t_acc, d[7] = ti;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..9, both CTAs)
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const uint32_t stage = epi_stage + (uint32_t)e * EPI_WARP_BYTES;  // this warp's staging tile + bias strip
    uint32_t ti = 0;
    long long t_wait = 0, t_pre = 0;
    const long long t_ebegin = clock64();
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    auto tile_info = [&](int pt, int& m_tile, int& n_tile, int& z) {
      z = pt / per_z;
      const int rem = pt - z * per_z;
      const int mp = rem / n_tiles;
      n_tile = rem - mp * n_tiles;
      m_tile = 2 * mp + (int)rank;
      return epi_tile_info<MODE>(g, q, m_tile, n_tile, z);
    };
    if (pair_id < total_pairs) {
      int a, b, c2;
      const EpiTile e0 = tile_info(pair_id, a, b, c2);
      epilogue_lookahead<MODE, 2>(g, e0, lane, half, nb);
    }
    for (int pt = pair_id; pt < total_pairs; pt += num_pairs, ++ti) {
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      int m_tile, n_tile, z;
      const EpiTile et = tile_info(pt, m_tile, n_tile, z);
      const long long tp0 = g.dbg ? clock64() : 0;
      epilogue_prefetch<MODE, 2>(g, et, stage, lane, half, n_tile, nb);  // bias / residual while the MMAs still run
      if (pt + num_pairs < total_pairs) {  // next tile's bias -> registers, residual lines -> L2
        int a, b, c2;
        const EpiTile en = tile_info(pt + num_pairs, a, b, c2);
        epilogue_lookahead<MODE, 2>(g, en, lane, half, nb);
      }
      const long long tw0 = g.dbg ? clock64() : 0;
      mbar_wait(bar_accf + 8u * as, aph);
      if (g.dbg) t_wait += clock64() - tw0, t_pre += tw0 - tp0;
      tc_fence_after();
      const uint32_t trow = tmem_base + as * 256u + ((uint32_t)(q * 32) << 16);
      epilogue_drain<MODE, 2>(g, et, trow, stage, lane, half, n_tile);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_acce + 8u * as);
        else mbar_arrive_remote(bar_acce + 8u * as, 0);
      }
    }
    if (g.dbg && warp == 2 && lane == 0) {
      g.dbg[(size_t)blockIdx.x * 8 + 4] = (unsigned long long)t_wait;
      g.dbg[(size_t)blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - t_ebegin);
      if (!leader) g.dbg[(size_t)blockIdx.x * 8 + 0] = (unsigned long long)t_pre;  // peer's MMA slots are free
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer may still be signalling the leader's barriers / reading its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace rfb
