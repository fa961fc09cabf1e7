// Argument block, operand modes and activation helpers shared by the tcgen05 GEMM kernels (gemm_persist.cuh,
// gemm_pair.cuh).  One kernel family serves every dense contraction on REFace's inference path: nn.Linear / 1x1 conv
// (plain 2-D A), 3x3 convolutions as implicit GEMM (A tiles are shifted NHWC boxes fetched by TMA, zero fill = padding),
// and the batched Q.K^T / P.V products of the materialised attention path.
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] (fp16, K-major, smem SW128)  x  B[BN x K]^T (fp16, K-major)
// (The first-generation non-persistent kernel that lived here was removed in round 2.)
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
  // taps per kernel row: 3 for a 3x3 convolution; 2 for the FOLDED nearest-2x-upsample + 3x3 convolution (`up` = 1): the
  // grid's z is the output phase (py, px) = (z >> 1, z & 1), tap (a, b) reads input pixel (y + a - 1 + py, x + b - 1 + px)
  // and the tile's rows are written to output pixels (2y + py, 2x + px) (input map: 2^up_wlog2 wide, 2^up_hwlog2 pixels)
  int ctw, up, up_wlog2, up_hwlog2;
  // GroupNorm statistics from the epilogue: per 32 output rows (one epilogue warp's accumulator rows) and channel the sum
  // and the sum of squares of the fp16 results, stats[(prow * N + col) * 2 + {0,1}], prow = m / 32 (folded upsample conv:
  // (m / 32) * 4 + phase, which keeps every sample's partial rows contiguous).  nullptr: off.
  float* stats;
  // A_PLAIN over TWO row-aligned sources (channel concatenation without the copy): k-blocks [0, nk1) come from tmA, the
  // rest from tmA2; a2_mod > 0: source 2 has only a2_mod rows and is read at row (m mod a2_mod) (CFG halves sharing a tensor)
  int nk1, a2_mod;
  int heads;  // *_HEADS4: z = n*heads + head
  // epilogue
  float alpha;
  const float* bias;    // [N] (packed order for GEGLU) or nullptr
  const float* rowvec;  // per-sample vector added per column: rowvec[(m / rows_per_vec) * ldv + col]
  int rows_per_vec, ldv;
  long long rowvec_zs;  // batched launches: rowvec advances by rowvec_zs floats per batch index z
  const float* act_param;  // PReLU slopes [N]
  int act, geglu;
  int relu_after_res;  // ReLU applied AFTER the residual add (ResNet BasicBlock: relu(shortcut + bn(conv)))
  const __half* res;  // residual [M, ldr] added after activation
  long long ldr;
  long long res_mod;  // > 0: the residual tensor has only res_mod rows, row m reads res[m mod res_mod] (multiple of 128)
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

static constexpr int GEMM_BM = 128;
static constexpr int GEMM_BK = 64;
static constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KiB


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

}  // namespace rfb
