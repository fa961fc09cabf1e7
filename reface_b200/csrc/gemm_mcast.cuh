// Cluster variant of the persistent tcgen05 GEMM / implicit-GEMM convolution with TMA MULTICAST of the weight tile.
//
// The CS CTAs that work on the same (N tile, K slice) and on CS consecutive M tiles form a thread-block cluster: each loads
// its own 128 x 64 A tile, but only 1/CS of the B tile, with cp.async.bulk.tensor ... .multicast::cluster, so that the
// slice lands in the shared memory of all CS CTAs.  Two uses (launch_gemm / conv3x3_t in engine.cu):
//  * every long-K launch (K >= 1024) that fills the GPU, as clusters of 2: the clock64 role trace of the 1-CTA kernel
//    shows its MMA warp waiting 20-37 % of the time for operands while the ring is not full -- operand delivery bounds
//    these launches, and fetching each weight tile once per pair of CTAs takes 9-14 % off them
//    (profiles/r02_gemm_role_trace_long_k.txt, r02_gemm_mcast_long_k_shapes.txt);
//  * the split-K slices of the 3x3 convolutions of the <= 8x8 maps (M = 64 rows per sample: few M tiles, a weight
//    operand that dwarfs the activation operand), as clusters of up to 8 (time-neutral there, kept for the L2 traffic).
//
// Protocol per CTA (warp0 TMA producer of the activation tiles, warp1 MMA issuer, warps 2..9 epilogue, warp 10 TMA
// producer of the weight slices):
//   full[s]   count 2 (one arrive.expect_tx per producer warp) + transaction bytes of the WHOLE stage (A tile + all CS
//             slices of B: every peer's multicast signals the barrier at the same offset in every destination CTA)
//   empty[s]  count CS: a slot may be overwritten by any peer, so every CTA's MMA warp releases it in ALL CTAs
//             (tcgen05.commit ... .multicast::cluster) and a producer refills it only when all CS consumers are done
//   acc_full / acc_empty: CTA-local, as in the 1-CTA kernel
// The CTAs of a cluster walk the same sequence of (K slice, M group, N tile) units in lock step; A tiles, accumulators
// and epilogues stay private, so the arithmetic (and every bit of the result) is that of gemm_persist_kernel.
//
// Measured and removed again (profiles/r02_gemm_cluster_variants.txt): sharing the ACTIVATION tile between two N tiles as
// well (clusters of 1 x 2 and 2 x 2) and issuing the 4-D activation box as two half boxes from two warps -- the
// activation-side TMA instruction occupies its warp for ~300 clocks per K block whatever its row count.
#pragma once
#include "gemm_persist.cuh"

namespace rfb {

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// completion of all prior MMAs of this thread arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// units: (z, m_group, n_tile), n_tile fastest; a cluster takes unit blockIdx.x / CS + i * (gridDim.x / CS)
// warps: 0 TMA (activations), 1 MMA, 2..9 epilogue, 10 TMA (weights)
static constexpr int GEMMC_B_WARP = GEMMP_THREADS / 32;
static constexpr int GEMMC_THREADS = GEMMP_THREADS + 32;

template <int MODE, int CS>
__global__ void __launch_bounds__(GEMMC_THREADS, 1)
gemm_mcast_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g,
                  const int m_groups, const int n_tiles, const int total_units) {
  constexpr int NP = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int S = g.stages;
  const int BN = g.BN;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)S * GEMM_A_STAGE_BYTES;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128u;
  const uint32_t b_slice_rows = (uint32_t)BN / CS;
  const uint32_t bars = sB + (uint32_t)S * b_stage_bytes;  // full[S], empty[S]
  const uint32_t bar_accf = bars + 16u * S;
  const uint32_t bar_acce = bar_accf + 16u;
  const uint32_t tptr = bar_acce + 16u;
  const uint32_t epi_stage = (tptr + 16u + 1023u) & ~1023u;
  constexpr uint16_t MASK = (uint16_t)((1u << CS) - 1u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bars + 8u * i, 2);  // the two producer warps
      mbar_init(bars + 8u * (S + i), CS);
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
  }
  cluster_sync_all();  // every CTA's barriers exist before a peer's multicast / commit can signal them
  if (warp == 1) tmem_alloc(tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: everything above overlapped the previous kernel's tail, nothing below may run before
  // that kernel has completed
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  const int per_z = m_groups * n_tiles;

  if (warp == 0 || warp == GEMMC_B_WARP) {
    // ------------------------------------------------------------ TMA producers: warp 0 loads the activation tiles,
    // warp GEMMC_B_WARP the weight slices.  Two warps because a TMA instruction occupies its issuing warp for ~80
    // (2-D box) to ~240 (4-D conv box) clocks (clock64 trace: one warp issuing both spent 72 % of a conv launch inside
    // them, 322 of the 441 clocks per K block, and the MMA warp waited 25 % of its time for operands): issued from two
    // warps the two latencies overlap.  Both arrive on full[s] (count 2) with their own byte counts.
    const bool load_a = warp == 0;
    const uint32_t tx = load_a ? GEMM_A_STAGE_BYTES : b_stage_bytes;
    uint32_t st = 0, sp = 0;
    long long t_empty = 0, t_issue = 0;  // clock64 role counters as in gemm_persist_kernel (option gemm_debug)
    for (int u = cluster_id; u < total_units; u += num_clusters) {
      const int z = u / per_z;
      const int rem = u - z * per_z;
      const int mg = rem / n_tiles, n_tile = rem - mg * n_tiles;
      const int m_tile = mg * CS + (int)rank;
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
      const int n0 = n_tile * BN + (int)(rank * b_slice_rows);
      for (int kb = 0; kb < g.nk; ++kb) {
        const uint32_t s = st, ph = sp;
        if (++st == (uint32_t)S) st = 0, sp ^= 1u;
        const long long tw0 = g.dbg ? clock64() : 0;
        mbar_wait(bars + 8u * (S + s), ph ^ 1u);  // all CS consumers have released this slot
        const long long ti0 = g.dbg ? clock64() : 0;
        if (g.dbg) t_empty += ti0 - tw0;
        const uint32_t full = bars + 8u * s;
        if (elect_one()) {
          mbar_expect_tx(full, tx);
          const uint32_t dA = sA + s * GEMM_A_STAGE_BYTES;
          const uint32_t dB = sB + s * b_stage_bytes + rank * b_slice_rows * 128u;
          const int kg = g.ksplit ? z * g.nk + kb : kb;
          if (!load_a) {
            tma_load_2d_mc(dB, &tmB, full, kg * GEMM_BK, n0, MASK);  // this CTA's slice of the weight tile -> all CTAs
          } else if (g.a_mode == A_PLAIN) {
            tma_load_2d(dA, &tmA, full, kg * GEMM_BK, m0);
          } else {
            const int tap = kg / g.cblocks;
            const int cb = kg - tap * g.cblocks;
            const int dy = tap / 3 - g.cpad_t, dx = tap % 3 - g.cpad_l;
            tma_load_4d(dA, &tmA, full, cb * GEMM_BK, cw + dx, ch + dy, cn);
          }
        }
        __syncwarp();
        if (g.dbg) t_issue += clock64() - ti0;
      }
    }
    if (g.dbg && lane == 0 && warp == 0) {
      g.dbg[(size_t)blockIdx.x * 8 + 3] = (unsigned long long)t_empty;
      g.dbg[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)t_issue;  // expect_tx + the activation TMA instruction
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = idesc_f16(GEMM_BM, (uint32_t)BN);
    uint32_t ti = 0, st = 0, sp = 0;
    long long t_full = 0, t_acc = 0;
    const long long t_begin = g.dbg ? clock64() : 0;
    for (int u = cluster_id; u < total_units; u += num_clusters, ++ti) {
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      long long tw0 = g.dbg ? clock64() : 0;
      mbar_wait(bar_acce + 8u * as, aph ^ 1u);
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
          mma_commit_mc(bars + 8u * (S + s), MASK);  // slot free in every CTA of the cluster once these MMAs retire
          if (kb == g.nk - 1) mma_commit(bar_accf + 8u * as);
        }
        __syncwarp();
      }
    }
    pdl_trigger();  // as in gemm_persist_kernel
    if (g.dbg && lane == 0) {
      unsigned long long* d = g.dbg + (size_t)blockIdx.x * 8;
      d[0] = (unsigned long long)(clock64() - t_begin), d[1] = (unsigned long long)t_full, d[2] = (unsigned long long)t_acc, d[7] = ti;
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..9), CTA-local
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const uint32_t stage = epi_stage + (uint32_t)e * EPI_WARP_BYTES;
    uint32_t ti = 0;
    long long t_wait = 0;
    const long long t_ebegin = g.dbg ? clock64() : 0;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    auto tile_of = [&](int u, int& n_tile) {
      const int z = u / per_z;
      const int rem = u - z * per_z;
      const int mg = rem / n_tiles;
      n_tile = rem - mg * n_tiles;
      return epi_tile_info<MODE>(g, q, mg * CS + (int)rank, n_tile, z);
    };
    if (cluster_id < total_units) {
      int nt;
      const EpiTile e0 = tile_of(cluster_id, nt);
      epilogue_lookahead<MODE, NP>(g, e0, lane, half, nb);
    }
    for (int u = cluster_id; u < total_units; u += num_clusters, ++ti) {
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      int n_tile;
      const EpiTile et = tile_of(u, n_tile);
      epilogue_prefetch<MODE, NP>(g, et, stage, lane, half, n_tile, nb);
      if (u + num_clusters < total_units) {
        int nt;
        const EpiTile en = tile_of(u + num_clusters, nt);
        epilogue_lookahead<MODE, NP>(g, en, lane, half, nb);
      }
      const long long tw0 = g.dbg ? clock64() : 0;
      mbar_wait(bar_accf + 8u * as, aph);
      if (g.dbg) t_wait += clock64() - tw0;
      tc_fence_after();
      const uint32_t trow = tmem_base + as * 256u + ((uint32_t)(q * 32) << 16);
      epilogue_drain<MODE, NP>(g, et, trow, stage, lane, half, n_tile);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8u * as);
    }
    if (g.dbg && warp == 2 && lane == 0) {
      g.dbg[(size_t)blockIdx.x * 8 + 4] = (unsigned long long)t_wait;
      g.dbg[(size_t)blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - t_ebegin);
    }
  }
  tc_fence_before();
  cluster_sync_all();  // peers may still be writing this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace rfb
