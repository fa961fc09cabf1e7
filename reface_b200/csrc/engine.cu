// Host-side engine: context, weight packing, TMA descriptor construction and kernel launchers.
#include "engine.h"

#include <cudaTypedefs.h>

#include <algorithm>
#include <cstring>

#include "elem.cuh"
#include "gemm_pair.cuh"
#include "gemm_mcast.cuh"

namespace rfb {

// ------------------------------------------------------------------------------------------ context
void* Ctx::alloc(size_t bytes) {
  size_t off = (arena_off + 255) & ~size_t(255);
  if (off + bytes > arena_cap)
    throw std::runtime_error("activation arena exhausted: need " + std::to_string(off + bytes) + " of " +
                             std::to_string(arena_cap) + " bytes (raise arena_bytes in rfb_init)");
  arena_off = off + bytes;
  arena_peak = std::max(arena_peak, arena_off);
  return arena + off;
}
void* Ctx::dmalloc(size_t bytes) {
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
  owned.push_back(p);
  return p;
}
const Param& Ctx::param(const std::string& name) const {
  auto it = params.find(name);
  if (it == params.end()) throw std::runtime_error("missing parameter: " + name);
  if (!it->second.f32)
    throw std::runtime_error("parameter " + name + " was released after packing (rfb_release_packed_originals): register it again");
  it->second.packed = true;
  return it->second;
}
const float* Ctx::pf(const std::string& name) const {
  const Param& p = param(name);
  p.packed = false, p.pinned = true;
  return p.f32;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

#define LAUNCH_CHECK(c)          \
  do {                           \
    CUDA_OK(cudaGetLastError()); \
    (c).launches++;              \
  } while (0)

// ------------------------------------------------------------------------------------------ packing
ConvW pack_conv(Ctx& c, const std::string& wname, const std::string& bname, const float* oscale) {
  const Param& p = c.param(wname);
  RFB_CHECK(p.shape.size() == 4 && p.shape[2] == p.shape[3], "conv weight must be [O,I,k,k]");
  ConvW w;
  w.cout = (int)p.shape[0], w.cin = (int)p.shape[1], w.ksz = (int)p.shape[2], w.taps = w.ksz * w.ksz;
  w.cin_p = w.cin;
  w.kp = round_up(w.taps * w.cin, 64);
  const int cout_p = round_up(w.cout, 32);
  w.w = (__half*)c.dmalloc((size_t)cout_p * w.kp * sizeof(__half));
  CUDA_OK(cudaMemsetAsync(w.w, 0, (size_t)cout_p * w.kp * sizeof(__half), c.stream));
  pack_conv_w_kernel<<<grid_for((long long)w.cout * w.kp), 256, 0, c.stream>>>(p.f32, w.w, w.cout, w.cin, w.taps,
                                                                               w.cin_p, w.kp, oscale);
  LAUNCH_CHECK(c);
  w.b = bname.empty() ? nullptr : c.pf(bname);
  return w;
}

LinW pack_linear_rows(Ctx& c, const std::vector<std::string>& wnames) {
  LinW w;
  int total = 0;
  for (auto& n : wnames) {
    const Param& p = c.param(n);
    RFB_CHECK(p.shape.size() >= 2, "linear weight must be at least 2-D");
    int in = 1;
    for (size_t i = 1; i < p.shape.size(); ++i) in *= (int)p.shape[i];
    RFB_CHECK(w.in == 0 || w.in == in, "fused linear weights must share the input width");
    w.in = in;
    total += (int)p.shape[0];
  }
  w.out = total;
  w.kp = round_up(w.in, 64);
  const int out_p = round_up(w.out, 32);
  w.w = (__half*)c.dmalloc((size_t)out_p * w.kp * sizeof(__half));
  CUDA_OK(cudaMemsetAsync(w.w, 0, (size_t)out_p * w.kp * sizeof(__half), c.stream));
  int row = 0;
  for (auto& n : wnames) {
    const Param& p = c.param(n);
    const int rows = (int)p.shape[0];
    pack_rows_f16_kernel<<<grid_for((long long)rows * w.kp), 256, 0, c.stream>>>(p.f32, w.w + (size_t)row * w.kp, rows,
                                                                                 w.in, w.kp);
    LAUNCH_CHECK(c);
    row += rows;
  }
  return w;
}
LinW pack_linear(Ctx& c, const std::string& wname, const std::string& bname) {
  LinW w = pack_linear_rows(c, {wname});
  w.b = bname.empty() ? nullptr : c.pf(bname);
  return w;
}

// GEGLU projection [2*inner, in]: rows permuted so that every BN-wide output tile holds BN/2 value rows
// followed by the matching BN/2 gate rows (the epilogue multiplies them in registers).
__global__ void pack_geglu_kernel(const float* __restrict__ w, const float* __restrict__ b, __half* __restrict__ wo,
                                  float* __restrict__ bo, int inner, int in, int kp, int BN) {
  const int hb = BN / 2;
  const long long total = (long long)2 * inner * kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kp);
    const int r = (int)(i / kp);  // packed row
    const int tile = r / BN, j = r % BN;
    const int src = (j < hb) ? tile * hb + j : inner + tile * hb + (j - hb);
    wo[i] = (k < in) ? __float2half_rn(w[(long long)src * in + k]) : __float2half_rn(0.f);
    if (k == 0) bo[r] = b[src];
  }
}
LinW pack_geglu(Ctx& c, const std::string& wname, const std::string& bname, int BN) {
  const Param& p = c.param(wname);
  LinW w;
  w.out = (int)p.shape[0], w.in = (int)p.shape[1], w.kp = round_up(w.in, 64);
  const int inner = w.out / 2;
  RFB_CHECK(w.out % BN == 0 && BN % 64 == 0, "GEGLU tile must divide the projection width");
  w.w = (__half*)c.dmalloc((size_t)w.out * w.kp * sizeof(__half));
  float* bo = (float*)c.dmalloc((size_t)w.out * sizeof(float));
  pack_geglu_kernel<<<grid_for((long long)w.out * w.kp), 256, 0, c.stream>>>(p.f32, c.pf(bname), w.w, bo, inner, w.in,
                                                                             w.kp, BN);
  LAUNCH_CHECK(c);
  w.b = bo;
  return w;
}
Lin32 lin32(Ctx& c, const std::string& wname, const std::string& bname) {
  const Param& p = c.param(wname);
  p.pinned = true;
  Lin32 w;
  w.w = p.f32, w.out = (int)p.shape[0], w.in = (int)(p.numel / p.shape[0]);
  w.b = bname.empty() ? nullptr : c.pf(bname);
  return w;
}

// ------------------------------------------------------------------------------------------ TMA descriptors
CUtensorMap make_tmap(Ctx& c, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                             const uint32_t* box, const uint32_t* elem_strides) {
  CUtensorMap m;
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  RFB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i) RFB_CHECK(strides_b[i] % 16 == 0, "TMA strides must be multiples of 16 bytes");
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(c.encode_fn);
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

int pick_bn(Ctx& c, long long M, int N, bool geglu, int K, bool allow16) {
  if (c.force_bn) return c.force_bn;
  if (geglu) return (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  if (N <= 32) return 32;
  if (c.gemm_wave_bn && K >= 1024) {
    // Long-K GEMMs are MMA-bound: choose the tile width that minimises (waves over the SMs) x (time of one tile).
    // One 128 x bn x 16 MMA takes max(bn/2, 32 + bn/4) cycles (tensor pipe vs. the 128 B/clk of smem operand reads);
    // e.g. M=4096, N=1280: bn=256 -> 160 tiles = 2 waves of 128-cycle steps, bn=144 -> 288 tiles = 2 waves of 72.
    // Widths that are only multiples of 16 need the vectorised epilogue (allow16).  The K order of every output
    // element's accumulation does not depend on bn, so results stay bitwise independent of the batch size.
    const long long mt = (M + GEMM_BM - 1) / GEMM_BM;
    int best = 0;
    double best_cost = 0;
    for (int bn = 256; bn >= 32; bn -= 16) {
      if ((bn & 31) && !allow16) continue;
      const long long nt = (N + bn - 1) / bn;
      const long long waves = (mt * nt + c.num_sms - 1) / c.num_sms;
      const double step = std::max(bn / 2.0, 32.0 + bn / 4.0);
      const double cost = (double)waves * (step * (K / 16.0) + 8.0 * bn + 1500.0);
      if (!best || cost < best_cost) best = bn, best_cost = cost;
    }
    return best;
  }
  static const int cand[] = {256, 224, 192, 160, 128, 96, 64, 32};
  int best = 32;
  long long best_cost = -1;
  for (int bn : cand) {
    const long long padded = (long long)((N + bn - 1) / bn) * bn;
    if (best_cost < 0 || padded < best_cost) best_cost = padded, best = bn;
  }
  // small problems: prefer more CTAs over the widest tile
  const long long mt = (M + GEMM_BM - 1) / GEMM_BM;
  while (best > 64 && mt * ((N + best - 1) / best) < c.num_sms) {
    int next = 0;
    for (int bn : cand)
      if (bn < best && (long long)((N + bn - 1) / bn) * bn <= best_cost + 32) { next = bn; break; }
    if (!next) break;
    best = next;
  }
  return best;
}

// EPI_FAST launches whose every tile is full and vectorisable take the kernel with the lean drain only (gemm_epilogue.cuh)
static bool epi_lean_ok(const Ctx& c, const GemmArgs& g) {
  const long long zo = g.zs_outer | g.zs_inner;
  return c.gemm_lean && g.out != nullptr && !g.up && g.M % GEMM_BM == 0 && g.N % g.BN == 0 && g.BN % 16 == 0 &&
         (g.N & 7) == 0 && (g.ldo & 7) == 0 && (zo & 7) == 0 && (!g.res || (g.ldr & 7) == 0) &&
         (!g.rowvec || ((g.ldv & 3) == 0 && g.rows_per_vec % 32 == 0 && (g.rowvec_zs & 3) == 0));
}

// Bplain/kp/nrows_w: the plain 2-D weight operand (lets the 2-CTA kernel rebuild the B map with a half-height box)
static void launch_gemm(Ctx& c, const CUtensorMap& tmA, const CUtensorMap& tmB, GemmArgs g, dim3 grid, double kalg,
                        const __half* Bplain = nullptr, int kp = 0, int nrows_w = 0, const CUtensorMap* tmA2p = nullptr,
                        int mcast_cs = 0, double kalg_ref = 0) {
  const CUtensorMap& tmA2 = tmA2p ? *tmA2p : tmA;
  if (g.ctw <= 0) g.ctw = 3;
  RFB_CHECK(!g.up || (!g.res && !g.rowvec && !g.out32 && !g.ksplit), "folded upsample conv: plain fp16 epilogue only");
  RFB_CHECK(!g.nk1 || (g.a_mode == A_PLAIN && tmA2p), "two-source A needs the plain operand mode");
  const int stage_bytes = GEMM_A_STAGE_BYTES + g.BN * 128;
  int stages = c.force_stages ? c.force_stages : std::max(2, std::min(6, c.gemm_smem_budget / stage_bytes));
  stages = std::min(stages, std::max(1, g.nk));
  g.stages = stages;
  g.tmem_cols = g.BN <= 32 ? 32 : g.BN <= 64 ? 64 : g.BN <= 128 ? 128 : 256;
  RFB_CHECK(g.BN % 16 == 0 && g.BN >= 32 && g.BN <= 256, "BN must be a multiple of 16 in [32,256]");
  if (g.zdiv <= 0) g.zdiv = 1;
  if (g.rows_per_vec <= 0) g.rows_per_vec = 1;
  if (g.o32_rpn <= 0) g.o32_rpn = 1;
  if (g.heads <= 0) g.heads = 1;
  Ctx::ProfRec rec;
  if (c.profile) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    rec.flops_exec = 2.0 * (double)g.M * (double)g.N * kalg * (double)grid.z;
    rec.flops = kalg_ref > 0 ? 2.0 * (double)g.M * (double)g.N * kalg_ref * (double)grid.z : rec.flops_exec;
    rec.kind = 0;
    rec.M = g.M, rec.N = g.N, rec.K = (int)kalg, rec.BN = g.BN, rec.z = (int)grid.z;
    rec.mode = g.a_mode * 100 + (g.geglu ? 10 : 0) + (g.res ? 1 : 0) + (g.rowvec ? 2 : 0) + (g.out32 ? 4 : 0);
    CUDA_OK(cudaEventRecord(rec.a, c.stream));
  }
  g.dbg = nullptr;
  if (c.gemm_debug) {  // per-CTA clock64 role breakdown of this launch (rfb_debug_read)
    // raw cudaMalloc on purpose: Ctx::owned is rolled back by the single-op test entry points
    if (!c.dbg_buf) CUDA_OK(cudaMalloc((void**)&c.dbg_buf, (size_t)c.num_sms * 8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c.dbg_buf, 0, (size_t)c.num_sms * 8 * sizeof(unsigned long long), c.stream));
    g.dbg = c.dbg_buf;
  }
  const bool pair_ok = c.gemm_pair && g.cstride <= 1 && !g.up && !g.nk1 && Bplain != nullptr && grid.z == 1 && grid.x >= 2 &&
                       (g.a_mode == A_PLAIN || g.a_mode == A_CONV3) && g.b_mode == B_PLAIN && g.nk >= c.gemm_pair_min_nk &&
                       (long long)grid.x * grid.y >= c.num_sms / 2;
  // Long-K launches are bound by L2->SM operand delivery (clock64 role trace: the MMA warp waits 26-37 % of its time for
  // operands while the ring is not full; profiles/r02_gemm_role_trace_long_k.txt): pairs of CTAs on consecutive M tiles
  // share every weight tile by TMA multicast (each fetches half of it), -9..-14 % per launch.  A work-distribution
  // choice from the launch shape only; the arithmetic of a tile is the 1-CTA kernel's.
  CUtensorMap tmB_mc;
  const CUtensorMap* tmBp = &tmB;
  if (mcast_cs == 0 && c.gemm_mcast_big > 1 && Bplain != nullptr && grid.z == 1 && !g.nk1 && !g.up && g.cstride <= 1 &&
      g.b_mode == B_PLAIN && (g.a_mode == A_PLAIN || g.a_mode == A_CONV3) && g.nk >= c.gemm_mcast_min_nk &&
      grid.x % c.gemm_mcast_big == 0 && g.BN % c.gemm_mcast_big == 0 && (g.BN / c.gemm_mcast_big) % 8 == 0 &&
      (long long)grid.x * grid.y >= c.num_sms && !(c.gemm_pair && g.nk >= c.gemm_pair_min_nk)) {
    mcast_cs = c.gemm_mcast_big;
    const uint64_t db[2] = {(uint64_t)kp, (uint64_t)nrows_w};
    const uint64_t sb[1] = {(uint64_t)kp * 2};
    const uint32_t bb[2] = {64, (uint32_t)(g.BN / mcast_cs)};
    tmB_mc = make_tmap(c, Bplain, 2, db, sb, bb);
    tmBp = &tmB_mc;
  }
  if (mcast_cs > 1) {
    // cluster of mcast_cs CTAs on consecutive M tiles sharing every weight tile by TMA multicast (gemm_mcast.cuh);
    // the B map has a box of BN / mcast_cs rows
    RFB_CHECK(grid.x % mcast_cs == 0 && (g.BN / mcast_cs) % 8 == 0 && g.BN % mcast_cs == 0 && g.b_mode == B_PLAIN &&
                  (g.a_mode == A_PLAIN || (g.a_mode == A_CONV3 && g.cstride == 1 && !g.up)) && !g.nk1,
              "multicast GEMM: shape not supported");
    const bool fast = g.act == 0 && !g.relu_after_res && g.out32 == nullptr && g.alpha == 1.0f;
    const int budget = (227 - 3) * 1024 - 4 * 2 * EPI_WARP_BYTES;
    g.stages = c.force_stages ? c.force_stages : std::max(2, std::min(8, budget / stage_bytes));
    const size_t psmem = gemmp_smem_bytes(g.stages, g.BN, 2);
    RFB_CHECK(psmem <= 227 * 1024, "GEMM smem over budget");
    const int m_groups = (int)grid.x / mcast_cs, n_tiles = (int)grid.y;
    const int units = m_groups * n_tiles * (int)grid.z;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(GEMMC_THREADS);
    cfg.dynamicSmemBytes = psmem;
    cfg.stream = c.stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)mcast_cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see launch_pdl (engine.h)
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
#define RFB_MCAST(MODE_, CS_)                                                                                          \
  do {                                                                                                                 \
    auto kfn = gemm_mcast_kernel<MODE_, CS_>;                                                                          \
    const std::string key = std::string("gemm_mcast_") + #MODE_ + "_" + #CS_;                                          \
    if (c.first_use(key.c_str())) {                                                                                    \
      CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                     \
      cfg.gridDim = dim3((unsigned)(c.num_sms / CS_ * CS_));                                                           \
      int nc = 0;                                                                                                      \
      if (cudaOccupancyMaxActiveClusters(&nc, kfn, &cfg) != cudaSuccess || nc < 1) nc = 1, cudaGetLastError();        \
      c.once_flags[key + "_max"] = nc;                                                                                 \
    }                                                                                                                  \
    const int clusters = std::min(units, c.once_flags[key + "_max"]);                                                  \
    cfg.gridDim = dim3((unsigned)(clusters * CS_));                                                                    \
    cfg.numAttrs = c.pdl_now() ? 2 : 1;                                                                                \
    CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, tmA, *tmBp, g, m_groups, n_tiles, units));                                   \
  } while (0)
    if (g.geglu) {
      RFB_CHECK(mcast_cs == 2, "multicast GEGLU GEMM: cluster of 2 only");
      RFB_MCAST(EPI_GEGLU, 2);
    } else if (fast) {
      if (mcast_cs == 2 && epi_lean_ok(c, g)) RFB_MCAST(EPI_LEAN, 2);
      else if (mcast_cs == 8) RFB_MCAST(EPI_FAST, 8);
      else if (mcast_cs == 4) RFB_MCAST(EPI_FAST, 4);
      else RFB_MCAST(EPI_FAST, 2);
    } else {
      if (mcast_cs == 8) RFB_MCAST(EPI_GENERIC, 8);
      else if (mcast_cs == 4) RFB_MCAST(EPI_GENERIC, 4);
      else RFB_MCAST(EPI_GENERIC, 2);
    }
#undef RFB_MCAST
  } else if (pair_ok) {
    // 2-CTA pairs (cta_group::2): each CTA loads its 128 A rows and half of the B tile
    if (c.first_use("gemm_pair")) {
      CUDA_OK(cudaFuncSetAttribute(gemm_pair_kernel<EPI_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CUDA_OK(cudaFuncSetAttribute(gemm_pair_kernel<EPI_GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CUDA_OK(cudaFuncSetAttribute(gemm_pair_kernel<EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    g.kmerge = (c.gemm_kmerge >= 2 && g.nk >= 4) ? 2 : 1;
    const int sb2 = g.kmerge * (GEMM_A_STAGE_BYTES + (g.BN / 2) * 128);
    g.stages = c.force_stages ? c.force_stages : std::max(2, std::min(8, (186 * 1024) / sb2));  // ~37 KB: epilogue staging
    const size_t psmem = gemm2_smem_bytes(g.stages, g.BN, g.kmerge);
    RFB_CHECK(psmem <= 227 * 1024, "GEMM smem over budget");
    const uint64_t db[2] = {(uint64_t)kp, (uint64_t)nrows_w};
    const uint64_t sb[1] = {(uint64_t)kp * 2};
    const uint32_t bb[2] = {64, (uint32_t)(g.BN / 2)};
    CUtensorMap tmB2 = make_tmap(c, Bplain, 2, db, sb, bb);
    const int m_pairs = ((int)grid.x + 1) / 2, n_tiles = (int)grid.y;
    const int total_pairs = m_pairs * n_tiles;
    const int pairs = std::min(total_pairs, c.num_sms / 2);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(GEMMP_THREADS);
    cfg.dynamicSmemBytes = psmem;
    cfg.stream = c.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    if (g.geglu)
      CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI_GEGLU>, tmA, tmB2, g, m_pairs, n_tiles, total_pairs));
    else if (g.act == 0 && !g.relu_after_res && g.out32 == nullptr && g.alpha == 1.0f)
      CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI_FAST>, tmA, tmB2, g, m_pairs, n_tiles, total_pairs));
    else
      CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI_GENERIC>, tmA, tmB2, g, m_pairs, n_tiles, total_pairs));
  } else {
    // persistent, double-buffered-accumulator kernel: one CTA per SM, deep smem ring
    if (c.first_use("gemm_persist")) {
      const int mx = 227 * 1024;
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_FAST, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_GEGLU, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_GENERIC, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_FAST, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_GEGLU, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_LEAN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
      CUDA_OK(cudaFuncSetAttribute(gemm_persist_kernel<EPI_LEAN, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    }
    // short-K GEMMs are epilogue-bound: 3 epilogue warps per lane quadrant (448 threads) instead of 2
    const int np = (g.nk <= c.gemm_epi3_max_nk && g.BN > 64 && g.act == 0 && !g.relu_after_res && g.out32 == nullptr && g.alpha == 1.0f) ? 3 : 2;
    const int budget = (227 - 3) * 1024 - 4 * np * EPI_WARP_BYTES;
    int ps = c.force_stages ? c.force_stages : std::max(2, std::min(8, budget / stage_bytes));
    g.stages = ps;
    const size_t psmem = gemmp_smem_bytes(ps, g.BN, np);
    RFB_CHECK(psmem <= 227 * 1024, "GEMM smem over budget");
    const int m_tiles = (int)grid.x, n_tiles = (int)grid.y;
    const int total = m_tiles * n_tiles * (int)grid.z;
    const int ctas = std::min(total, c.num_sms);
    const int thr = 64 + np * 128;
    if (g.geglu) {
      if (np == 3) launch_pdl(c, gemm_persist_kernel<EPI_GEGLU, 3>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
      else launch_pdl(c, gemm_persist_kernel<EPI_GEGLU, 2>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
    } else if (g.act == 0 && !g.relu_after_res && g.out32 == nullptr && g.alpha == 1.0f) {
      // every tile full and vectorisable -> the kernel with the lean drain only (gemm_epilogue.cuh)
      if (epi_lean_ok(c, g)) {
        if (np == 3) launch_pdl(c, gemm_persist_kernel<EPI_LEAN, 3>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
        else launch_pdl(c, gemm_persist_kernel<EPI_LEAN, 2>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
      } else if (np == 3) launch_pdl(c, gemm_persist_kernel<EPI_FAST, 3>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
      else launch_pdl(c, gemm_persist_kernel<EPI_FAST, 2>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
    } else {
      launch_pdl(c, gemm_persist_kernel<EPI_GENERIC, 2>, dim3(ctas), dim3(thr), psmem, tmA, tmB, tmA2, g, m_tiles, n_tiles, total);
    }
  }
  LAUNCH_CHECK(c);
  if (c.profile) {
    CUDA_OK(cudaEventRecord(rec.b, c.stream));
    c.prof.push_back(rec);
  }
}

// The epilogue can leave GroupNorm partial statistics only on its vectorised fp16 path (gemm_epilogue.cuh, EPI_FAST).
static bool epi_stats_ok(const Ctx& c, const Epi& e, int N, long long ldo, long long rows_per_sample) {
  return c.gn_epi_stats && e.want_stats && !e.geglu && e.act == 0 && !e.relu_after_res && e.out32 == nullptr &&
         e.alpha == 1.0f && (N & 7) == 0 && ldo == N && (!e.res || (e.ldr & 7) == 0) && rows_per_sample % 32 == 0;
}
static void fill_epi(GemmArgs& g, const Epi& e, __half* out, long long ldo) {
  g.alpha = e.alpha, g.bias = e.bias, g.rowvec = e.rowvec, g.rows_per_vec = e.rows_per_vec, g.ldv = e.ldv;
  g.act_param = e.act_param, g.act = e.act, g.geglu = e.geglu, g.res = e.res, g.ldr = e.ldr;
  g.relu_after_res = e.relu_after_res;
  g.res_mod = e.res_mod;
  g.stats = e.stats_out;
  g.out = out, g.ldo = ldo, g.out32 = e.out32, g.o32_sn = e.o32_sn, g.o32_sp = e.o32_sp, g.o32_sc = e.o32_sc;
  g.o32_rpn = e.o32_rpn;
}

// Split-K for the long-K 3x3 convolutions of the smallest maps (<= 8x8 output pixels per sample: the deepest UNet level,
// M = 64 rows per sample).  At batch 8 that is M = 1024, N = 1280, K = 11520-23040: 40 tiles of 128x256 leave most SMs idle,
// narrower tiles fill them but are L2-bandwidth bound on operand re-reads (555 TFLOP/s,
// profiles/r01s2_gemm_shapes_per_launch.txt).  K is cut into 3 equal slices of 128x256 tiles; every unit writes its raw fp32
// accumulators and splitk_reduce_kernel folds the slices in a fixed order and applies the epilogue.  The slice count is
// a function of the PER-SAMPLE shape only (never of the batch or the SM count), so the K order of every output element's
// accumulation -- and therefore every bit of the result -- is independent of the batch size, like everything else in
// the path (test_batch_independence).  Option gemm_splitk = 0 switches it off.
static int pick_ksplit(Ctx& c, int rows_per_sample, int N, int nk, const Epi& e, long long ldo) {
  if (!c.gemm_splitk || c.force_bn || nk < 90 || nk % 3 || rows_per_sample > 64 || e.geglu || e.act || e.out32 ||
      e.alpha != 1.0f || e.relu_after_res || (N & 7) || (ldo & 7) || (e.res && (e.ldr & 3)))
    return 1;
  return 3;
}
static void splitk_finish(Ctx& c, const float* part, int ks, long long M, int N, const Epi& e, __half* out, long long ldo,
                          float* stats) {
  if (M % 32 == 0 && N % 128 == 0 && (!e.res || (e.ldr & 3) == 0)) {
    dim3 grid((unsigned)(M / 32), (unsigned)(N / 128));
    launch_pdl(c, splitk_reduce_stats_kernel, grid, dim3(128), 0, part, ks, M, N, e.bias, e.rowvec,
                                                          e.rows_per_vec > 0 ? e.rows_per_vec : 1, e.ldv, e.res, e.ldr, out, ldo,
                                                          stats);
  } else {
    RFB_CHECK(stats == nullptr, "split-K statistics need M % 32 == 0 and N % 128 == 0");
    launch_pdl(c, splitk_reduce_kernel, dim3(grid_for(M * (N / 4))), dim3(256), 0, part, ks, M, N, e.bias, e.rowvec,
                                                                       e.rows_per_vec > 0 ? e.rows_per_vec : 1, e.ldv, e.res,
                                                                       e.ldr, out, ldo);
  }
  LAUNCH_CHECK(c);
}

void gemm(Ctx& c, const __half* A, long long lda, long long M, int K, const __half* W, int kp, int N, __half* out,
          long long ldo, const Epi& e, int force_bn, int kalg) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = (int)M, g.N = N, g.nk = (K + 63) / 64;
  const bool vec_epi = e.out32 == nullptr && (N & 7) == 0 && (ldo & 7) == 0 && (!e.res || (e.ldr & 7) == 0);
  g.BN = force_bn ? force_bn : pick_bn(c, M, N, e.geglu != 0, K, vec_epi);
  g.a_mode = A_PLAIN, g.b_mode = B_PLAIN;
  fill_epi(g, e, out, ldo);
  const uint64_t da[2] = {(uint64_t)K, (uint64_t)M};
  const uint64_t sa[1] = {(uint64_t)lda * 2};
  const uint32_t ba[2] = {64, 128};
  const int nrows_w = round_up(N, 32);
  const uint64_t db[2] = {(uint64_t)kp, (uint64_t)nrows_w};
  const uint64_t sb[1] = {(uint64_t)kp * 2};
  const uint32_t bb[2] = {64, (uint32_t)g.BN};
  CUtensorMap tmA = make_tmap(c, A, 2, da, sa, ba);
  CUtensorMap tmB = make_tmap(c, W, 2, db, sb, bb);
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((N + g.BN - 1) / g.BN), 1);
  launch_gemm(c, tmA, tmB, g, grid, (double)(kalg > 0 ? kalg : K), W, kp, nrows_w);
}

// out = [A1 | A2] W^T (+ epilogue): the channel concatenation of two row-aligned activations as ONE K loop over two TMA
// descriptors (torch.cat of the UNet skip connections, openaimodel.py:897-899, without materialising the copy).
// K1 % 64 == 0; a2_rows > 0: A2 has only a2_rows rows and row m reads A2[m mod a2_rows] (a2_rows % 128 == 0).
void gemm2(Ctx& c, const __half* A1, long long lda1, int K1, const __half* A2, long long lda2, int K2, long long a2_rows,
           long long M, const __half* W, int kp, int N, __half* out, long long ldo, const Epi& e) {
  RFB_CHECK(K1 % 64 == 0 && K2 > 0, "two-source GEMM: the first source must be a multiple of 64 channels wide");
  RFB_CHECK(a2_rows == 0 || (a2_rows % GEMM_BM == 0 && M % a2_rows == 0), "two-source GEMM: bad row period of source 2");
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  const int K = K1 + K2;
  g.M = (int)M, g.N = N, g.nk = (K + 63) / 64, g.nk1 = K1 / 64, g.a2_mod = (int)a2_rows;
  const bool vec_epi = e.out32 == nullptr && (N & 7) == 0 && (ldo & 7) == 0 && (!e.res || (e.ldr & 7) == 0);
  g.BN = pick_bn(c, M, N, e.geglu != 0, K, vec_epi);
  g.a_mode = A_PLAIN, g.b_mode = B_PLAIN;
  fill_epi(g, e, out, ldo);
  const uint64_t d1[2] = {(uint64_t)K1, (uint64_t)M};
  const uint64_t s1[1] = {(uint64_t)lda1 * 2};
  const uint64_t d2[2] = {(uint64_t)K2, (uint64_t)(a2_rows ? a2_rows : M)};
  const uint64_t s2[1] = {(uint64_t)lda2 * 2};
  const uint32_t ba[2] = {64, 128};
  const int nrows_w = round_up(N, 32);
  const uint64_t db[2] = {(uint64_t)kp, (uint64_t)nrows_w};
  const uint64_t sb[1] = {(uint64_t)kp * 2};
  const uint32_t bb[2] = {64, (uint32_t)g.BN};
  CUtensorMap tmA = make_tmap(c, A1, 2, d1, s1, ba);
  CUtensorMap tmA2 = make_tmap(c, A2, 2, d2, s2, ba);
  CUtensorMap tmB = make_tmap(c, W, 2, db, sb, bb);
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((N + g.BN - 1) / g.BN), 1);
  launch_gemm(c, tmA, tmB, g, grid, (double)K, nullptr, 0, 0, &tmA2);
}

Tens linear_t(Ctx& c, const Tens& x, const LinW& w, Epi e) {
  RFB_CHECK(x.c == w.in, "linear: input width mismatch");
  const int out_c = e.geglu ? w.out / 2 : w.out;
  Tens y = c.new_tens(x.n, x.h, x.w, out_c);
  if (!e.bias) e.bias = w.b;
  e.stats_out = nullptr;
  if (epi_stats_ok(c, e, w.out, out_c, (long long)x.h * x.w))
    e.stats_out = y.stats = c.alloc_t<float>((size_t)(y.rows() / 32) * out_c * 2);
  gemm(c, x.p, x.c, x.rows(), x.c, w.w, w.kp, w.out, y.p, out_c, e);
  return y;
}

// Implicit GEMM straight from the NHWC activation: stride 1 with padding 1, or (option conv_tma_stride2) stride 2 with
// any top/left padding in {0, 1} -- the A tile of tap (ky, kx) is then a 4-D TMA box with element strides {1,2,2,1}
// starting at (2*ox0 + kx - pad_l, 2*oy0 + ky - pad_t); out-of-range pixels are zero-filled by TMA on every side.
static bool conv_tma_ok(Ctx& c, const Tens& x, const ConvW& w, int stride, int pt, int pl, int pb, int pr, int Ho, int Wo) {
  if (w.ksz != 3 || w.cin % 64 != 0) return false;
  if (stride == 1) {
    if (pt != 1 || pl != 1 || pb != 1 || pr != 1) return false;
  } else if (stride == 2) {
    if (!c.conv_tma_stride2 || pt < 0 || pt > 1 || pl < 0 || pl > 1) return false;
  } else {
    return false;
  }
  const int W = Wo, H = Ho;  // tiles are cut from the OUTPUT map
  if (W >= 128) return W % 128 == 0;
  if (128 % W != 0) return false;
  const int rows = 128 / W;  // image rows per tile
  if (H >= rows) return H % rows == 0;
  return rows % H == 0;
}

Tens conv3x3_t(Ctx& c, const Tens& x, const ConvW& w, Epi e, int stride, int pad_t, int pad_l, int pad_b, int pad_r) {
  RFB_CHECK(x.c == w.cin, "conv: channel mismatch");
  const int Ho = (x.h + pad_t + pad_b - w.ksz) / stride + 1;
  const int Wo = (x.w + pad_l + pad_r - w.ksz) / stride + 1;
  Tens y;
  y.n = x.n, y.h = Ho, y.w = Wo, y.c = w.cout;
  const bool f16_out = (e.out32 == nullptr);
  if (f16_out) y.p = c.alloc_t<__half>((size_t)y.rows() * y.c);
  if (!e.bias) e.bias = w.b;
  if (e.rowvec && e.rows_per_vec <= 1) e.rows_per_vec = Ho * Wo;
  const long long M = y.rows();
  const bool tma_path = !(w.ksz == 1 && stride == 1) && conv_tma_ok(c, x, w, stride, pad_t, pad_l, pad_b, pad_r, Ho, Wo);
  const bool split = tma_path && pick_ksplit(c, Ho * Wo, w.cout, 9 * (w.cin / 64), e, y.c) > 1;
  e.stats_out = nullptr;
  // (split-K convs get their statistics from the reduction kernel instead of the GEMM epilogue)
  if (f16_out && epi_stats_ok(c, e, w.cout, y.c, (long long)Ho * Wo) && (!split || (M % 32 == 0 && w.cout % 128 == 0)))
    e.stats_out = y.stats = c.alloc_t<float>((size_t)(M / 32) * y.c * 2);
  if (w.ksz == 1 && stride == 1) {
    gemm(c, x.p, x.c, M, x.c, w.w, w.kp, w.cout, y.p, y.c, e);
    return y;
  }
  if (tma_path) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M, g.N = w.cout, g.cblocks = w.cin / 64, g.nk = 9 * g.cblocks;
    const int ks = pick_ksplit(c, Ho * Wo, w.cout, g.nk, e, y.c);
    float* part = nullptr;
    size_t mk_split = 0;
    Epi e_full = e;
    if (ks > 1) {  // raw fp32 partial tiles per K slice, epilogue applied by the reduction (see pick_ksplit)
      mk_split = c.mark();
      part = c.alloc_t<float>((size_t)ks * M * w.cout);
      g.nk /= ks, g.ksplit = 1;
      e = Epi();
      e.out32 = part;
    }
    g.BN = ks > 1 ? 256
                  : pick_bn(c, M, w.cout, false, 9 * w.cin,
                            e.out32 == nullptr && (w.cout & 7) == 0 && (!e.res || (e.ldr & 7) == 0));
    g.a_mode = A_CONV3, g.b_mode = B_PLAIN;
    g.bw = std::min(Wo, 128);
    g.bh = std::min(Ho, 128 / g.bw);
    g.bimg = 128 / (g.bw * g.bh);
    g.tiles_w = Wo / g.bw, g.tiles_h = Ho / g.bh;
    g.cstride = stride, g.cpad_l = pad_l, g.cpad_t = pad_t;
    fill_epi(g, e, ks > 1 ? nullptr : y.p, y.c);
    const uint64_t da[4] = {(uint64_t)x.c, (uint64_t)x.w, (uint64_t)x.h, (uint64_t)x.n};
    const uint64_t sa[3] = {(uint64_t)x.c * 2, (uint64_t)x.w * x.c * 2, (uint64_t)x.h * x.w * x.c * 2};
    // with element strides the box spans stride * (elements loaded) positions of the traversed dimension
    const uint32_t ba[4] = {64, (uint32_t)(g.bw * stride), (uint32_t)(g.bh * stride), (uint32_t)g.bimg};
    const uint32_t ea[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    // Weight-tile multicast across a cluster of CTAs on consecutive M tiles (gemm_mcast.cuh) where the weight operand
    // dominates the operand traffic: the split-K convolutions of the <= 8x8 maps.  The cluster size follows from the M
    // tile count (a pure work-distribution choice: the arithmetic is the 1-CTA kernel's).
    int cs = 0;
    if (c.gemm_mcast && ks > 1 && stride == 1 && M % 128 == 0) {
      const long long mt = M / 128;
      cs = mt % 8 == 0 ? 8 : (mt % 4 == 0 ? 4 : (mt % 2 == 0 ? 2 : 0));
    }
    const uint64_t db[2] = {(uint64_t)w.kp, (uint64_t)round_up(w.cout, 32)};
    const uint64_t sb[1] = {(uint64_t)w.kp * 2};
    const uint32_t bb[2] = {64, (uint32_t)(cs > 1 ? g.BN / cs : g.BN)};
    CUtensorMap tmA = make_tmap(c, x.p, 4, da, sa, ba, stride > 1 ? ea : nullptr);
    CUtensorMap tmB = make_tmap(c, w.w, 2, db, sb, bb);
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((w.cout + g.BN - 1) / g.BN), (unsigned)ks);
    launch_gemm(c, tmA, tmB, g, grid, 9.0 * w.cin / ks, w.w, w.kp, round_up(w.cout, 32), nullptr, cs);
    if (ks > 1) {
      splitk_finish(c, part, ks, M, w.cout, e_full, y.p, y.c, e_full.stats_out);
      c.release(mk_split);
    }
    return y;
  }
  // generic path: explicit im2col then a plain GEMM
  const size_t mk = c.mark();
  __half* col = nullptr;
  {
    // allocate after the output so that releasing the mark keeps y alive
    col = c.alloc_t<__half>((size_t)M * w.kp);
  }
  im2col_kernel<<<grid_for(M * (w.kp / 8)), 256, 0, c.stream>>>(x.p, col, x.n, x.h, x.w, x.c, w.ksz, w.ksz, stride, pad_t,
                                                              pad_l, Ho, Wo, w.kp);
  LAUNCH_CHECK(c);
  gemm(c, col, w.kp, M, w.kp, w.w, w.kp, w.cout, y.p, y.c, e, 0, w.taps * w.cin);
  c.release(mk);
  return y;
}

// ------------------------------------------------------------------------------------------ folded upsample conv
// Upsample.forward = nearest-2x interpolation followed by a 3x3 convolution (openaimodel.py:109-119, model.py:53-66).
// Output pixel (2y+py, 2x+px) only ever sees the 2x2 input neighbourhood {y-1+py, y+py} x {x-1+px, x+px}: the three kernel
// rows collapse onto two input rows (py = 0: {W0} on y-1, {W1+W2} on y; py = 1: {W0+W1} on y, {W2} on y+1), likewise the
// columns.  The layer is therefore FOUR 2x2 convolutions over the low-resolution input (one per output phase) with
// pre-summed weights: 4/9 of the multiply-adds, no up-sampled tensor in HBM, zero fill at the border identical to the
// padding of the original formulation.  Weights: [4 phases][cout_p][4 taps * cin] fp16, summed in fp32 before rounding.
__global__ void pack_upconv_w_kernel(const float* __restrict__ src, __half* __restrict__ dst, int O, int Op, int I, int Kp) {
  const long long total = 4ll * Op * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const int o = (int)((idx / Kp) % Op);
    const int ph = (int)(idx / ((long long)Kp * Op));
    const int tap = k / I, i = k % I;
    float v = 0.f;
    if (o < O && tap < 4) {
      const int py = ph >> 1, px = ph & 1, a = tap >> 1, b = tap & 1;
      // kernel rows / columns that land on input row a (column b) for this phase
      const int ky0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), ky1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
      const int kx0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), kx1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
      const float* w = src + ((long long)o * I + i) * 9;
      for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx) v += w[ky * 3 + kx];
    }
    dst[idx] = __float2half_rn(v);
  }
}
ConvW pack_upconv(Ctx& c, const std::string& wname, const std::string& bname) {
  const Param& p = c.param(wname);
  RFB_CHECK(p.shape.size() == 4 && p.shape[2] == 3 && p.shape[3] == 3, "upsample conv weight must be [O,I,3,3]");
  ConvW w;
  w.cout = (int)p.shape[0], w.cin = (int)p.shape[1], w.ksz = 3, w.taps = 4, w.cin_p = w.cin, w.up = 1;
  RFB_CHECK(w.cin % 64 == 0, "folded upsample conv needs Cin % 64 == 0");
  w.kp = 4 * w.cin;
  const int cout_p = round_up(w.cout, 32);
  w.w = (__half*)c.dmalloc((size_t)4 * cout_p * w.kp * sizeof(__half));
  pack_upconv_w_kernel<<<grid_for(4ll * cout_p * w.kp), 256, 0, c.stream>>>(p.f32, w.w, w.cout, cout_p, w.cin, w.kp);
  LAUNCH_CHECK(c);
  w.b = bname.empty() ? nullptr : c.pf(bname);
  return w;
}
static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}
// conv3x3(nearest_2x(x)) with the folded weights of pack_upconv: x [n,h,w,cin] -> [n,2h,2w,cout]
Tens upconv3x3_t(Ctx& c, const Tens& x, const ConvW& w, Epi e) {
  RFB_CHECK(w.up && x.c == w.cin, "upconv: weights were not packed by pack_upconv / channel mismatch");
  const int wl = ilog2_exact(x.w), hwl = ilog2_exact(x.h * x.w);
  RFB_CHECK(wl >= 0 && hwl >= 0 && conv_tma_ok(c, x, w, 1, 1, 1, 1, 1, x.h, x.w), "upconv: map size not supported");
  Tens y = c.new_tens(x.n, 2 * x.h, 2 * x.w, w.cout);
  if (!e.bias) e.bias = w.b;
  const long long M = x.rows();
  e.stats_out = nullptr;
  if (epi_stats_ok(c, e, w.cout, y.c, (long long)x.h * x.w))
    e.stats_out = y.stats = c.alloc_t<float>((size_t)(y.rows() / 32) * y.c * 2);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = (int)M, g.N = w.cout, g.cblocks = w.cin / 64, g.nk = 4 * g.cblocks;
  g.BN = pick_bn(c, M, w.cout, false, 4 * w.cin, (w.cout & 7) == 0);
  g.a_mode = A_CONV3, g.b_mode = B_BATCH3;
  g.bw = std::min(x.w, 128);
  g.bh = std::min(x.h, 128 / g.bw);
  g.bimg = 128 / (g.bw * g.bh);
  g.tiles_w = x.w / g.bw, g.tiles_h = x.h / g.bh;
  g.cstride = 1, g.cpad_l = 1, g.cpad_t = 1, g.ctw = 2, g.up = 1, g.up_wlog2 = wl, g.up_hwlog2 = hwl;
  fill_epi(g, e, y.p, y.c);
  g.zdiv = 1, g.zs_outer = 0, g.zs_inner = 0;
  const int cout_p = round_up(w.cout, 32);
  const uint64_t da[4] = {(uint64_t)x.c, (uint64_t)x.w, (uint64_t)x.h, (uint64_t)x.n};
  const uint64_t sa[3] = {(uint64_t)x.c * 2, (uint64_t)x.w * x.c * 2, (uint64_t)x.h * x.w * x.c * 2};
  const uint32_t ba[4] = {64, (uint32_t)g.bw, (uint32_t)g.bh, (uint32_t)g.bimg};
  const uint64_t db[3] = {(uint64_t)w.kp, (uint64_t)cout_p, 4};
  const uint64_t sb[2] = {(uint64_t)w.kp * 2, (uint64_t)cout_p * w.kp * 2};
  const uint32_t bb[3] = {64, (uint32_t)g.BN, 1};
  CUtensorMap tmA = make_tmap(c, x.p, 4, da, sa, ba);
  CUtensorMap tmB = make_tmap(c, w.w, 3, db, sb, bb);
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((w.cout + g.BN - 1) / g.BN), 4);
  launch_gemm(c, tmA, tmB, g, grid, 4.0 * w.cin, nullptr, 0, 0, nullptr, 0, 9.0 * w.cin);  // executes 4 of the 9 taps' MACs
  return y;
}

// ------------------------------------------------------------------------------------------ attention (materialised)
// qkv: fused projection rows [N*L, ldq]; q/k/v start at column offsets q_off/k_off/v_off, head h at +h*d.
// S = scale * Q K^T (fp16, [N*heads, L, Lp]) -> row softmax -> O = P V written to out[N*L, ldo] at column h*d.
// [ref: ldm/modules/attention.py:204-220]
void attention(Ctx& c, const __half* qkv, long long ldq, int N, int L, int heads, int d, __half* out, long long ldo,
               float scale, int q_off, int k_off, int v_off, int hs) {
  if (hs <= 0) hs = d;
  if (c.attn_flash && attention_flash(c, qkv, ldq, N, L, heads, d, out, ldo, scale, q_off, k_off, v_off, hs)) return;
  const size_t mk = c.mark();
  const int Z = N * heads;
  const int Lp = round_up(L, 8);
  __half* S = c.alloc_t<__half>((size_t)Z * L * Lp);
  __half* Vt = c.alloc_t<__half>((size_t)Z * d * Lp);
  {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = L, g.N = L, g.nk = (d + 63) / 64;
    g.BN = pick_bn(c, (long long)L * Z, L, false);
    g.a_mode = A_HEADS4, g.b_mode = B_HEADS4, g.heads = heads;
    Epi e;
    e.alpha = scale;
    fill_epi(g, e, S, Lp);
    g.zdiv = 1, g.zs_outer = (long long)L * Lp, g.zs_inner = 0;
    const uint64_t dq[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)L, (uint64_t)N};
    const uint64_t sq[3] = {(uint64_t)hs * 2, (uint64_t)ldq * 2, (uint64_t)L * ldq * 2};
    const uint32_t bq[4] = {64, 1, 128, 1};
    const uint32_t bk[4] = {64, 1, (uint32_t)g.BN, 1};
    CUtensorMap tmA = make_tmap(c, qkv + q_off, 4, dq, sq, bq);
    CUtensorMap tmB = make_tmap(c, qkv + k_off, 4, dq, sq, bk);
    dim3 grid((unsigned)((L + 127) / 128), (unsigned)((L + g.BN - 1) / g.BN), (unsigned)Z);
    launch_gemm(c, tmA, tmB, g, grid, (double)d);
  }
  {
    const long long rows = (long long)Z * L;
    if (Lp <= 256)
      softmax_rows_warp_kernel<8><<<(unsigned)((rows + 7) / 8), 256, 0, c.stream>>>(S, rows, L, Lp);
    else if (Lp <= 1024)
      softmax_rows_warp_kernel<32><<<(unsigned)((rows + 7) / 8), 256, 0, c.stream>>>(S, rows, L, Lp);
    else
      softmax_rows_kernel<<<(unsigned)rows, 256, 0, c.stream>>>(S, rows, L, Lp);
    LAUNCH_CHECK(c);
  }
  {
    dim3 grid((unsigned)((Lp + 31) / 32), (unsigned)((d + 31) / 32), (unsigned)Z), block(32, 8);
    transpose_v_kernel<<<grid, block, 0, c.stream>>>(qkv + v_off, Vt, N, L, heads, d, ldq, Lp, hs);
    LAUNCH_CHECK(c);
  }
  {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = L, g.N = d, g.nk = (Lp + 63) / 64;
    g.BN = std::min(256, round_up(d, 32));
    g.a_mode = A_BATCH3, g.b_mode = B_BATCH3;
    Epi e;
    fill_epi(g, e, out, ldo);
    g.zdiv = heads, g.zs_outer = (long long)L * ldo, g.zs_inner = d;
    const uint64_t dp[3] = {(uint64_t)Lp, (uint64_t)L, (uint64_t)Z};
    const uint64_t sp[2] = {(uint64_t)Lp * 2, (uint64_t)L * Lp * 2};
    const uint32_t bp[3] = {64, 128, 1};
    const uint64_t dv[3] = {(uint64_t)Lp, (uint64_t)d, (uint64_t)Z};
    const uint64_t sv[2] = {(uint64_t)Lp * 2, (uint64_t)d * Lp * 2};
    const uint32_t bv[3] = {64, (uint32_t)g.BN, 1};
    CUtensorMap tmA = make_tmap(c, S, 3, dp, sp, bp);
    CUtensorMap tmB = make_tmap(c, Vt, 3, dv, sv, bv);
    dim3 grid((unsigned)((L + 127) / 128), (unsigned)((d + g.BN - 1) / g.BN), (unsigned)Z);
    launch_gemm(c, tmA, tmB, g, grid, (double)L);
  }
  c.release(mk);
}

// ------------------------------------------------------------------------------------------ normalisation etc.
// GroupNorm(32) of [x1 | x2] as a per-(sample, channel) affine map y = x * a + b, folded from the partial statistics the
// producing epilogues left with the tensors: ab[n][c] = (rstd * gamma, beta - mean * rstd * gamma) (arena, [N][C])
float2* gn_affine_from_stats(Ctx& c, const Tens& x1, const Tens& x2, const float* gamma, const float* beta, float eps) {
  const int C = x1.c + (x2.p ? x2.c : 0), N = x1.n, HW = x1.h * x1.w;
  RFB_CHECK(x1.stats && (!x2.p || x2.stats) && HW % 32 == 0 && C % 32 == 0 && C / 32 <= 256, "GroupNorm: no producer statistics");
  float2* ab = c.alloc_t<float2>((size_t)N * C);
  GnStatSrc ss{x1.stats, x2.p ? x2.stats : nullptr, x1.c, (x2.p && x2.n != x1.n) ? x2.n : 0};
  launch_pdl(c, gn_finalize3_kernel, dim3(32, (unsigned)N), dim3(256), 0, ss, gamma, beta, ab, HW / 32, HW, C, 32, eps);
  LAUNCH_CHECK(c);
  return ab;
}
Tens groupnorm(Ctx& c, const Tens& x, const float* gamma, const float* beta, float eps, bool silu) {
  return groupnorm2(c, x, Tens(), gamma, beta, eps, silu);
}
// GroupNorm(32) of the channel concatenation [x1 | x2] (x2.p == nullptr: of x1 alone) -> one contiguous NHWC tensor.
// x2 may hold fewer samples than x1 (x1.n % x2.n == 0): sample n reads x2[n mod x2.n].
Tens groupnorm2(Ctx& c, const Tens& x1, const Tens& x2, const float* gamma, const float* beta, float eps, bool silu) {
  const int C = x1.c + (x2.p ? x2.c : 0), N = x1.n, HW = x1.h * x1.w;
  RFB_CHECK(C % 32 == 0 && x1.c % 8 == 0 && C % 8 == 0, "GroupNorm(32) needs C % 32 == 0 and 8-channel aligned sources");
  RFB_CHECK(!x2.p || (x2.h == x1.h && x2.w == x1.w && x2.n > 0 && x1.n % x2.n == 0), "GroupNorm: sources do not line up");
  GnSrc src{x1.p, x2.p, x1.c, (x2.p && x2.n != x1.n) ? x2.n : 0};
  Tens y = c.new_tens(N, x1.h, x1.w, C);
  const size_t mk = c.mark();
  const int cv = C / 8;
  RFB_CHECK(cv <= 512, "GroupNorm: too many channels");
  int R = std::max(1, 512 / cv);
  if (c.gn_epi_stats && x1.stats && (!x2.p || x2.stats) && HW % 32 == 0) {
    // statistics came with the tensor(s) from the producing epilogue: fold them per (sample, group), then ONE streaming
    // pass (the tensor is read once instead of twice)
    float2* ab = gn_affine_from_stats(c, x1, x2, gamma, beta, eps);
    // gn_apply_bps blocks per SM over the whole batch; a function of the batch only through the work split, never the values
    const int want = std::max(1, (c.gn_apply_bps * c.num_sms) / std::max(1, N));
    const int slab = std::max(R, (HW + want - 1) / want);
    dim3 g3((unsigned)((HW + slab - 1) / slab), (unsigned)N);
    launch_pdl(c, gn_apply3_kernel, g3, dim3(cv * R), 0, src, ab, y.p, HW, C, silu ? 1 : 0, slab);
    LAUNCH_CHECK(c);
    c.release(mk);
    return y;
  }
  // One cluster launch for small and medium maps; the whole-grid two-launch path (statistics, apply with the finalize
  // folded in) for maps of >= gn_fused_max_elems elements per sample (64^2 x 640 and up, the VAE's 256^2 / 512^2 levels),
  // where 16 CTAs per sample are too few to stream from HBM (profiles/r01s2_micro_bench.txt).  The choice depends on the
  // per-sample shape only, so results stay independent of the batch size.
  if (c.gn_fused && (long long)HW * C < c.gn_fused_max_elems) {
    R = std::max(1, std::min(c.gn_threads, 512) / cv);
    // one launch: a cluster of 16 (or 8) CTAs per sample (statistics exchanged through DSMEM), see elem.cuh
    int& max_cluster = c.gn_max_cluster;
    if (c.first_use("gn_fused")) {
      CUDA_OK(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      CUDA_OK(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      // 16-CTA clusters are a non-portable size: fall back to the portable 8 where the device (e.g. a partitioned
      // GPU) cannot co-schedule them.  The choice is fixed per context, so results stay reproducible.
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(16, 1), q.blockDim = dim3(512), q.dynamicSmemBytes = 48 * 1024;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = 1;
      q.attrs = qa, q.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, gn_fused_kernel, &q) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        max_cluster = 8;
      }
    }
    const int GC = std::min(c.gn_cluster, max_cluster);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)GC, (unsigned)N);
    cfg.blockDim = dim3((unsigned)(cv * R));
    cfg.dynamicSmemBytes = ((size_t)std::max(R * 2 * C, 64 * GC) + 2 * C + 4 * 32) * sizeof(float);
    cfg.stream = c.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = GC, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    RFB_CHECK(cfg.dynamicSmemBytes <= 96 * 1024, "GroupNorm: smem over budget");
    CUDA_OK(cudaLaunchKernelEx(&cfg, gn_fused_kernel, src, gamma, beta, y.p, HW, C, 32, eps, silu ? 1 : 0));
    LAUNCH_CHECK(c);
    return y;
  }
  // two launches: whole-grid statistics (8 loads in flight per thread) + apply with the finalize folded in
  const int slab = std::max(R, (HW + 31) / 32);  // <= 32 slabs per sample, a function of the shape only
  const int nslab = (HW + slab - 1) / slab;
  float* partial = c.alloc_t<float>((size_t)N * nslab * 32 * 2);
  if (c.first_use("gn_stats2"))
    CUDA_OK(cudaFuncSetAttribute(gn_stats2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  dim3 g1((unsigned)nslab, (unsigned)N);
  gn_stats2_kernel<<<g1, cv * R, (size_t)(R + 1) * 2 * C * sizeof(float), c.stream>>>(src, partial, HW, C, slab, 32);
  LAUNCH_CHECK(c);
  const int want2 = std::max(1, (8 * c.num_sms) / std::max(1, N));
  const int slab2 = std::max(R, (HW + want2 - 1) / want2);
  dim3 g2((unsigned)((HW + slab2 - 1) / slab2), (unsigned)N);
  RFB_CHECK(cv * R >= 256, "GroupNorm: block too small for the in-kernel finalize");
  gn_apply2_kernel<<<g2, cv * R, 0, c.stream>>>(src, partial, nslab, gamma, beta, y.p, HW, C, 32, eps, silu ? 1 : 0, slab2);
  LAUNCH_CHECK(c);
  c.release(mk);
  return y;
}

// conv1x1(GroupNorm(x)) with the activation-free GroupNorm folded into per-sample weights (elem.cuh:
// gn_fold_weights_kernel); x must carry producer statistics.  w32: fp32 [Cout, Cin] (the 1x1 conv weight), bias [Cout].
Tens conv1x1_gn_folded(Ctx& c, const Tens& x, const float* gn_gamma, const float* gn_beta, float eps, const float* w32,
                       const float* bias, int Cout) {
  const int Cin = x.c, N = x.n, HW = x.h * x.w;
  RFB_CHECK(Cin % 64 == 0 && Cout % 8 == 0 && HW % 128 == 0, "folded GroupNorm + 1x1 conv: shape not supported");
  Tens y = c.new_tens(N, x.h, x.w, Cout);
  const size_t mk = c.mark();
  float2* ab = gn_affine_from_stats(c, x, Tens(), gn_gamma, gn_beta, eps);
  const int kp = Cin, cout_p = round_up(Cout, 32);
  __half* Wn = c.alloc_t<__half>((size_t)N * cout_p * kp);
  float* biasn = c.alloc_t<float>((size_t)N * Cout);
  launch_pdl(c, gn_fold_weights_kernel, dim3((unsigned)((cout_p + 7) / 8), (unsigned)N), dim3(256), 0, w32, bias, ab, Wn, biasn, Cin,
                                                                                             Cout, kp, cout_p);
  LAUNCH_CHECK(c);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = HW, g.N = Cout, g.nk = Cin / 64;
  g.BN = pick_bn(c, (long long)HW * N, Cout, false, Cin, true);
  g.a_mode = A_BATCH3, g.b_mode = B_BATCH3;
  Epi e;
  e.rowvec = biasn, e.ldv = 0, e.rows_per_vec = HW;
  fill_epi(g, e, y.p, Cout);
  g.rowvec_zs = Cout;
  g.zdiv = 1, g.zs_outer = (long long)HW * Cout, g.zs_inner = 0;
  const uint64_t da[3] = {(uint64_t)Cin, (uint64_t)HW, (uint64_t)N};
  const uint64_t sa[2] = {(uint64_t)Cin * 2, (uint64_t)HW * Cin * 2};
  const uint32_t ba[3] = {64, 128, 1};
  const uint64_t db[3] = {(uint64_t)kp, (uint64_t)cout_p, (uint64_t)N};
  const uint64_t sb[2] = {(uint64_t)kp * 2, (uint64_t)cout_p * kp * 2};
  const uint32_t bb[3] = {64, (uint32_t)g.BN, 1};
  CUtensorMap tmA = make_tmap(c, x.p, 3, da, sa, ba);
  CUtensorMap tmB = make_tmap(c, Wn, 3, db, sb, bb);
  dim3 grid((unsigned)(HW / 128), (unsigned)((Cout + g.BN - 1) / g.BN), (unsigned)N);
  launch_gemm(c, tmA, tmB, g, grid, (double)Cin);
  c.release(mk);
  return y;
}

// 16-byte-vectorised LayerNorm where C = 8 * LPR * VPL fits (LPR lanes per row, VPL vectors per lane)
template <int LPR, int VPL>
static void launch_ln_vec(Ctx& c, const __half* x, const float* gamma, const float* beta, __half* y, long long rows,
                          long long ldx, long long ldy, float eps) {
  const long long rows_per_block = 8 * (32 / LPR);
  launch_pdl(c, layernorm_vec_kernel<LPR, VPL>, dim3((unsigned)((rows + rows_per_block - 1) / rows_per_block)), dim3(256), 0, 
      x, gamma, beta, y, rows, ldx, ldy, eps);
  LAUNCH_CHECK(c);
}
bool layernorm_vec(Ctx& c, const __half* x, const float* gamma, const float* beta, __half* y, long long rows, int C,
                   long long ldx, long long ldy, float eps) {
  if ((C & 7) || (ldx & 7) || (ldy & 7)) return false;
  const int nv = C >> 3;
#define RFB_LN_CASE(LPR, VPL)                                                 \
  if (nv == LPR * VPL) {                                                      \
    launch_ln_vec<LPR, VPL>(c, x, gamma, beta, y, rows, ldx, ldy, eps);       \
    return true;                                                              \
  }
  RFB_LN_CASE(32, 5) RFB_LN_CASE(32, 4) RFB_LN_CASE(32, 3) RFB_LN_CASE(32, 2) RFB_LN_CASE(32, 1)
  RFB_LN_CASE(16, 5) RFB_LN_CASE(16, 3) RFB_LN_CASE(16, 1)
  RFB_LN_CASE(8, 5) RFB_LN_CASE(8, 3) RFB_LN_CASE(8, 1)
#undef RFB_LN_CASE
  return false;
}

Tens layernorm(Ctx& c, const Tens& x, const float* gamma, const float* beta, float eps) {
  RFB_CHECK(x.c % 64 == 0 && x.c <= 2048, "LayerNorm: C must be a multiple of 64, <= 2048");
  Tens y = c.new_tens(x.n, x.h, x.w, x.c);
  const long long rows = x.rows();
  if (c.ln_vec && layernorm_vec(c, x.p, gamma, beta, y.p, rows, x.c, x.c, x.c, eps)) return y;
  layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, c.stream>>>(x.p, gamma, beta, y.p, rows, x.c, x.c, x.c, eps);
  LAUNCH_CHECK(c);
  return y;
}

void cross_attn_small(Ctx& c, const __half* q, const float* kc, const float* vc, __half* out, int N, int L, int T, int C,
                      int heads) {
  RFB_CHECK(T >= 1 && T <= 16, "cross-attention: context length must be in [1, 16]");
  RFB_CHECK(C % heads == 0 && (C / heads) % 8 == 0, "cross-attention: head dim must be a multiple of 8");
  const size_t smem = (size_t)2 * T * C * sizeof(float);
  if (c.first_use("cross_attn_small"))
    CUDA_OK(cudaFuncSetAttribute(cross_attn_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  RFB_CHECK(smem <= 200 * 1024, "cross-attention: context does not fit shared memory");
  dim3 grid((unsigned)((L * heads + 255) / 256), (unsigned)N);
  launch_pdl(c, cross_attn_small_kernel<16>, grid, dim3(256), smem, q, kc, vc, out, L, T, C, heads,
                                                             1.0f / sqrtf((float)(C / heads)));
  LAUNCH_CHECK(c);
}

Tens upsample2x(Ctx& c, const Tens& x) {
  Tens y = c.new_tens(x.n, x.h * 2, x.w * 2, x.c);
  upsample2x_kernel<<<grid_for(y.rows() * (x.c / 8)), 256, 0, c.stream>>>(x.p, y.p, x.n, x.h, x.w, x.c);
  LAUNCH_CHECK(c);
  return y;
}
Tens concat_c(Ctx& c, const Tens& a, const Tens& b) {
  RFB_CHECK(a.rows() == b.rows() && a.c % 8 == 0 && b.c % 8 == 0, "concat: shape mismatch");
  Tens y = c.new_tens(a.n, a.h, a.w, a.c + b.c);
  concat_c_kernel<<<grid_for(y.rows() * (y.c / 8)), 256, 0, c.stream>>>(a.p, b.p, y.p, a.rows(), a.c, b.c);
  LAUNCH_CHECK(c);
  return y;
}
Tens from_nchw_f32(Ctx& c, const float* src, int N, int C, int H, int W, int Cp) {
  Tens y = c.new_tens(N, H, W, Cp);
  nchw_f32_to_nhwc_f16_kernel<<<grid_for(y.rows() * Cp), 256, 0, c.stream>>>(src, y.p, N, C, H * W, Cp);
  LAUNCH_CHECK(c);
  return y;
}
void to_nchw_f32(Ctx& c, const Tens& x, float* dst) {
  nhwc_f16_to_nchw_f32_kernel<<<grid_for(x.rows() * x.c), 256, 0, c.stream>>>(x.p, dst, x.n, x.c, x.h * x.w, x.c);
  LAUNCH_CHECK(c);
}
void linear_small(Ctx& c, const float* x, long long ldx, int R, const Lin32& w, float* out, long long ldo, int act_in,
                  int act_out, const float* res) {
  dim3 grid((unsigned)((w.out + 7) / 8));
  for (int r0 = 0; r0 < R; r0 += 16) {
    const int r = std::min(16, R - r0);
    const float* xs = x + (long long)r0 * ldx;
    float* os = out + (long long)r0 * ldo;
    const float* rs = res ? res + (long long)r0 * ldo : nullptr;
    if (r <= 2)
      launch_pdl(c, linear_small_kernel<2>, grid, dim3(256), 0, xs, w.w, w.b, os, r, w.in, w.out, ldx, ldo, act_in, act_out, rs);
    else if (r <= 4)
      launch_pdl(c, linear_small_kernel<4>, grid, dim3(256), 0, xs, w.w, w.b, os, r, w.in, w.out, ldx, ldo, act_in, act_out, rs);
    else
      launch_pdl(c, linear_small_kernel<16>, grid, dim3(256), 0, xs, w.w, w.b, os, r, w.in, w.out, ldx, ldo, act_in, act_out, rs);
    LAUNCH_CHECK(c);
  }
}
void timestep_embedding(Ctx& c, const long long* t, float* out, int N, int dim) {
  timestep_embedding_kernel<<<(N * dim / 2 + 127) / 128, 128, 0, c.stream>>>(t, out, N, dim);
  LAUNCH_CHECK(c);
}
void concat9(Ctx& c, const float* x, const float* z, const float* mask, float* out, int B, int HW, int dup) {
  launch_pdl(c, concat9_kernel, dim3(grid_for((long long)dup * B * 9 * HW)), dim3(256), 0, x, z, mask, out, B, HW, dup);
  LAUNCH_CHECK(c);
}
void cfg_ddim_update(Ctx& c, const float* x, const float* eps2, const float* noise, float* x_prev, float* pred_x0,
                     long long count, float scale, float a_t, float a_prev, float sigma, float sqrt_one_minus_at,
                     int has_uncond) {
  cfg_ddim_update_kernel<<<grid_for(count), 256, 0, c.stream>>>(x, eps2, noise, x_prev, pred_x0, count, scale, a_t,
                                                                a_prev, sigma, sqrt_one_minus_at, has_uncond);
  LAUNCH_CHECK(c);
}

void eps_from_taps(Ctx& c, const float* taps, const float* bias, float* eps, int N, int L) {
  eps_from_taps_kernel<<<grid_for((long long)N * L * L), 256, 0, c.stream>>>(taps, bias, eps, N, L);
  LAUNCH_CHECK(c);
}
void taps_cfg_ddim_update(Ctx& c, const float* x, const float* taps, const float* bias, const float* noise, float* x_prev,
                          float* pred_x0, int B, int L, float scale, float a_t, float a_prev, float sigma,
                          float sqrt_one_minus_at, int has_uncond) {
  launch_pdl(c, taps_cfg_ddim_update_kernel, dim3(grid_for((long long)B * L * L, 128)), dim3(128), 0, 
      x, taps, bias, noise, x_prev, pred_x0, B, L, scale, a_t, a_prev, sigma, sqrt_one_minus_at, has_uncond);
  LAUNCH_CHECK(c);
}
LinW pack_out_taps(Ctx& c, const std::string& wname) {
  const Param& p = c.param(wname);
  RFB_CHECK(p.shape.size() == 4 && p.shape[0] == 4 && p.shape[2] == 3 && p.shape[3] == 3, "output conv must be [4,C,3,3]");
  LinW w;
  w.in = (int)p.shape[1], w.out = 36, w.kp = round_up(w.in, 64);
  w.w = (__half*)c.dmalloc((size_t)64 * w.kp * sizeof(__half));
  pack_out_taps_kernel<<<grid_for(64ll * w.kp), 256, 0, c.stream>>>(p.f32, w.w, w.in, w.kp);
  LAUNCH_CHECK(c);
  return w;
}
void cfg_combine(Ctx& c, const float* eps2, float* out, long long count, float scale) {
  cfg_combine_kernel<<<grid_for(count), 256, 0, c.stream>>>(eps2, out, count, scale);
  LAUNCH_CHECK(c);
}
void plms_combine(Ctx& c, const float* e_t, const float* o1, const float* o2, const float* o3, const float* e_next,
                  float* out, long long count, int order) {
  plms_combine_kernel<<<grid_for(count), 256, 0, c.stream>>>(e_t, o1, o2, o3, e_next, out, count, order);
  LAUNCH_CHECK(c);
}
void q_sample(Ctx& c, const float* x0, const float* noise, const float* coef_dev, float* out, long long per_sample, int B) {
  q_sample_kernel<<<grid_for(per_sample * B), 256, 0, c.stream>>>(x0, noise, coef_dev, out, per_sample, B);
  LAUNCH_CHECK(c);
}

}  // namespace rfb
