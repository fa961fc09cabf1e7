// Host-side engine: context (weights, activation arena, stream), op launchers over the CUDA kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace rfb {

struct Param {
  float* f32 = nullptr;  // device copy, fp32, original layout (nullptr after rfb_release_packed_originals)
  std::vector<int64_t> shape;
  size_t numel = 0;
  // how the builders used it: `packed` = converted into an fp16 GEMM operand (the fp32 original is then dead weight),
  // `pinned` = a built model reads the fp32 buffer itself (biases, norm affine vectors, small-M GEMV weights)
  mutable bool packed = false, pinned = false;
};

struct Tens {  // NHWC fp16 activation
  __half* p = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  // GroupNorm partial statistics left by the producing GEMM / conv epilogue: [rows / 32][c][2] (sum, sum of squares of
  // the fp16 values per 32-row chunk and channel); nullptr when the producer did not write them
  float* stats = nullptr;
  long long rows() const { return (long long)n * h * w; }
};

struct ConvW {  // packed [cout_p, kp] fp16, k = tap*cin_p + c
  __half* w = nullptr;
  const float* b = nullptr;
  int cin = 0, cout = 0, cin_p = 0, taps = 0, kp = 0, ksz = 0;
  int up = 0;  // packed by pack_upconv: [4 phases][cout_p][4 taps * cin] (nearest-2x upsample folded into the 3x3 conv)
};
struct LinW {  // packed [out_p, kp] fp16
  __half* w = nullptr;
  const float* b = nullptr;
  int in = 0, out = 0, kp = 0;
};
struct Lin32 {  // small-M path keeps fp32 weights
  const float* w = nullptr;
  const float* b = nullptr;
  int in = 0, out = 0;
};

struct Epi {
  const float* bias = nullptr;
  const float* rowvec = nullptr;
  int rows_per_vec = 1, ldv = 0;
  int act = 0;
  const float* act_param = nullptr;
  int geglu = 0;
  int relu_after_res = 0;  // ReLU after the residual add (needs the generic epilogue)
  const __half* res = nullptr;
  long long ldr = 0;
  bool want_stats = false;  // ask the conv / linear launcher to leave GroupNorm partial statistics with the output (Tens::stats)
  float* stats_out = nullptr;  // (set by the launcher)
  long long res_mod = 0;  // residual has only res_mod rows and is read at row (m mod res_mod) (CFG halves sharing a tensor)
  float alpha = 1.0f;
  float* out32 = nullptr;
  long long o32_sn = 0, o32_sp = 0, o32_sc = 0;
  int o32_rpn = 1;
};

struct Ctx;
struct UNet;
struct VAE;
struct ClipVision;
struct ArcFace;
struct FaceParser;

struct Ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::unordered_map<std::string, Param> params;
  std::vector<void*> owned;  // packed weights etc. (cudaMalloc'd); moved into the model that packed them (capi.cu)
  std::vector<void*> retired;  // fp32 buffers of re-registered parameters whose size changed (freed at destroy)
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0, arena_peak = 0;
  std::string err;
  void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
  long long launches = 0;     // kernels launched by this library (reported by bench)
  int gemm_smem_budget = 110 * 1024;  // per-CTA smem target (2 CTAs/SM)
  int force_bn = 0, force_stages = 0, attn_flash = 4, gemm_kmerge = 1;
  // 2-CTA (cta_group::2) kernel: correct and faster on isolated long-K GEMMs (1180 vs 1114 TFLOP/s) but measured
  // ~4 % slower over the whole UNet step than 1-CTA tiles (profiles/r01_gemm_sweep_v2.txt) -> opt-in
  int gemm_pair = 0;
  int attn_stagger = 0;  // attention v4: cycles by which the second query tile's softmax starts late
  int attn_pad = 0;      // test hook: rfb_op_attention repacks q/k/v with 64-element head slices
  int attn_pingpong = 0; // (opt-in: +2 % on one box, -3 % on another) attention v4: the two query tiles' softmax warps take turns on the MUFU unit (see attn_flash.cu)
  int attn_poly = 0;     // attention v3/v4: exponentials per 8 evaluated on the FMA pipe (measured slower: off)
  int cfg_share = 1;     // samplers: compute the context-independent head of the UNet once per CFG pair
  int conv_tma_stride2 = 1;  // stride-2 3x3 convs as implicit GEMM through strided TMA boxes (0: explicit im2col)
  // programmatic dependent launch between consecutive kernels (launch_pdl below).  1: eager launches only -- measured
  // on one box, interleaved: eager UNet forward 18.17 -> 17.41 ms, but the graph-replayed bench 9.32 -> 9.15 faces/s
  // (the graph's own kernel-to-kernel edges are cheaper than the release of griddepcontrol.wait); 2: captured graphs too
  int pdl = 1;
  bool pdl_now() const { return pdl == 2 || (pdl == 1 && (gstream == nullptr || stream != gstream)); }
  int gemm_lean = 1;     // lean-drain kernel for launches whose every tile is full and vectorisable
  // weight-tile multicast for every long-K GEMM / conv that fills the GPU: cluster size (0: off) and the K blocks from
  // which a launch counts as long-K (short-K launches are epilogue-bound)
  int gemm_mcast_big = 2, gemm_mcast_min_nk = 16;
  int gemm_mcast = 1;    // weight-tile TMA multicast across a cluster of M tiles for the split-K convs (gemm_mcast.cuh)
  int gemm_splitk = 1;   // 3-way split-K for the long-K 3x3 convs of <= 8x8 maps (slice count from the per-sample shape only)
  int gemm_wave_bn = 1;  // long-K GEMMs: wave-quantisation-aware tile width (multiples of 16)
  int gn_epi_stats = 1;  // GroupNorm from the producer epilogue's partial statistics where available (finalize + streaming apply)
  int gn_fold = 1;  // SpatialTransformer: fold the activation-free GroupNorm into per-sample proj_in weights (maps >= 32x32)
  int gn_apply_bps = 4;  // streaming GroupNorm apply: blocks per SM
  int gn_fused = 1;  // single-launch cluster GroupNorm (0: stats / finalize / apply kernels)
  long long gn_fused_max_elems = 2621440;  // = 64*64*640: per-sample H*W*C from which GroupNorm takes the whole-grid path
  int gn_cluster = 16, gn_threads = 512;  // fused GroupNorm: CTAs per sample (cluster size), threads per CTA
  int ln_vec = 1;    // 16-byte-vectorised LayerNorm (0: one warp per row, 4-byte loads)
  int gemm_epi3_max_nk = 10;  // K <= 640: 3 epilogue warps per TMEM lane quadrant
  int gemm_pair_min_nk = 12;  // CTA pairs only for K >= 768: short-K GEMMs are epilogue-bound and measured faster on 1-CTA tiles
  // optional per-launch CUDA-event timing of the tensor-core kernels (bench.py roofline)
  int profile = 0;       // 1: time every tensor-core launch; 2: also print one line per launch in rfb_profile_read
  int gemm_debug = 0;                       // per-CTA clock64 counters of the last 2-CTA GEMM launch
  unsigned long long* dbg_buf = nullptr;    // [num_sms * 8]
  // flops: ALGORITHMIC (reference-equivalent: the folded upsample conv counts the 9-tap conv at output resolution it
  // replaces); flops_exec: what the tensor cores executed
  struct ProfRec { cudaEvent_t a, b; double flops; int kind; int M = 0, N = 0, K = 0, BN = 0, mode = 0, z = 1; double flops_exec = 0; };
  std::vector<ProfRec> prof;
  // CUDA graphs of whole sampling loops (unet.cu: ddim_sample), keyed by shape / schedule / arena mark / graph_epoch
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int seen = 0; bool failed = false; long long launches = 0; };
  std::unordered_map<std::string, GraphEntry> graphs;
  cudaStream_t gstream = nullptr;  // capture stream (the caller's stream may be the legacy default stream)
  int use_graph = 1;               // option "use_graph"
  long long graph_epoch = 0;       // bumped by every option / weight change: invalidates the cached graphs
  long long graph_replays = 0;
  void drop_graphs() {
    for (auto& kv : graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
    ++graph_epoch;
  }
  UNet* unet = nullptr;
  VAE* vae = nullptr;
  ClipVision* clip = nullptr;
  ArcFace* arc = nullptr;
  FaceParser* parser = nullptr;

  // Kernel attributes (dynamic smem limit, non-portable cluster size) are PER-DEVICE state: every context sets them
  // once for its own device (a process-wide `static bool` would leave a second Engine(device=1) without them).
  std::unordered_map<std::string, int> once_flags;
  bool first_use(const char* key) { return once_flags.emplace(key, 1).second; }
  int gn_max_cluster = 16;
  void* alloc(size_t bytes);  // arena bump allocation (256 B aligned)
  size_t mark() const { return arena_off; }
  void release(size_t m) { arena_off = m; }
  template <class T>
  T* alloc_t(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
  Tens new_tens(int n, int h, int w, int c) {
    Tens t;
    t.n = n, t.h = h, t.w = w, t.c = c;
    t.p = alloc_t<__half>((size_t)t.rows() * c);
    return t;
  }
  void* dmalloc(size_t bytes);  // persistent device allocation (weights)
  const Param& param(const std::string& name) const;  // for packing (marks it `packed`)
  const float* pf(const std::string& name) const;      // fp32 buffer a model keeps reading (marks it `pinned`)
  bool has(const std::string& name) const { return params.count(name) != 0; }
};

#define RFB_CHECK(cond, msg)                                                                  \
  do {                                                                                        \
    if (!(cond)) throw std::runtime_error(std::string(msg) + " [" #cond "] at " __FILE__ ":" + \
                                          std::to_string(__LINE__));                          \
  } while (0)
#define CUDA_OK(expr)                                                                                     \
  do {                                                                                                    \
    cudaError_t e_ = (expr);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #expr " at " \
                               __FILE__ ":" + std::to_string(__LINE__));                                  \
  } while (0)

// Launch on the context's stream with the programmatic-stream-serialization attribute (option "pdl"): the kernel may be
// scheduled while its predecessor drains and must call pdl_wait() (ptx.cuh) before it touches anything the predecessor
// wrote or still reads.  ONLY for kernels that do so; every other launch keeps full stream serialization.  Stream
// capture turns the attribute into programmatic edges of the graph.
template <typename... KArgs, typename... Args>
inline void launch_pdl(Ctx& c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = c.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = c.pdl_now() ? 1 : 0;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

// ---- weight packing (device side, at build time)
// [O,I,KH,KW] -> implicit-GEMM layout; oscale (device, [O]) optionally folds a per-output-channel scale (BatchNorm)
ConvW pack_conv(Ctx& c, const std::string& wname, const std::string& bname, const float* oscale = nullptr);
LinW pack_linear(Ctx& c, const std::string& wname, const std::string& bname);
LinW pack_linear_rows(Ctx& c, const std::vector<std::string>& wnames);  // concatenated along out (fused QKV)
LinW pack_geglu(Ctx& c, const std::string& wname, const std::string& bname, int BN);
Lin32 lin32(Ctx& c, const std::string& wname, const std::string& bname);
int pick_bn(Ctx& c, long long M, int N, bool geglu, int K = 0, bool allow16 = false);

// TMA descriptor (fp16, 128-byte swizzle, zero fill out of bounds); dims/box innermost first, strides in bytes
CUtensorMap make_tmap(Ctx& c, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box, const uint32_t* elem_strides = nullptr);
// fused flash-style attention (tcgen05, S and O tiles in TMEM); returns false when the shape is not covered
// hs = distance in elements between consecutive heads inside a q / k / v section (0: packed, hs = d)
bool attention_flash(Ctx& c, const __half* qkv, long long ldq, int N, int L, int heads, int d, __half* out,
                     long long ldo, float scale, int q_off, int k_off, int v_off, int hs = 0);

// ---- op launchers (all asynchronous on c.stream)
void gemm(Ctx& c, const __half* A, long long lda, long long M, int K, const __half* W, int kp, int N, __half* out,
          long long ldo, const Epi& e, int force_bn = 0, int kalg = 0);
void conv3x3(Ctx& c, const Tens& x, const ConvW& w, __half* out, long long ldo, Epi e, int stride = 1, int pad_t = 1,
             int pad_l = 1, int Ho = -1, int Wo = -1);
Tens conv3x3_t(Ctx& c, const Tens& x, const ConvW& w, Epi e, int stride = 1, int pad_t = 1, int pad_l = 1,
               int pad_b = 1, int pad_r = 1);
void gemm2(Ctx& c, const __half* A1, long long lda1, int K1, const __half* A2, long long lda2, int K2, long long a2_rows,
           long long M, const __half* W, int kp, int N, __half* out, long long ldo, const Epi& e);
ConvW pack_upconv(Ctx& c, const std::string& wname, const std::string& bname);
Tens upconv3x3_t(Ctx& c, const Tens& x, const ConvW& w, Epi e);
Tens linear_t(Ctx& c, const Tens& x, const LinW& w, Epi e);
void attention(Ctx& c, const __half* qkv, long long ldq, int N, int L, int heads, int d, __half* out, long long ldo,
               float scale, int q_off, int k_off, int v_off, int hs = 0);
// softmax_T(q k^T / sqrt(d)) v for a short context (T <= 16 tokens per sample); kc / vc fp32 [N*T, C]
void cross_attn_small(Ctx& c, const __half* q, const float* kc, const float* vc, __half* out, int N, int L, int T, int C,
                      int heads);
Tens groupnorm(Ctx& c, const Tens& x, const float* gamma, const float* beta, float eps, bool silu);
float2* gn_affine_from_stats(Ctx& c, const Tens& x1, const Tens& x2, const float* gamma, const float* beta, float eps);
Tens conv1x1_gn_folded(Ctx& c, const Tens& x, const float* gn_gamma, const float* gn_beta, float eps, const float* w32,
                       const float* bias, int Cout);
Tens groupnorm2(Ctx& c, const Tens& x1, const Tens& x2, const float* gamma, const float* beta, float eps, bool silu);
Tens layernorm(Ctx& c, const Tens& x, const float* gamma, const float* beta, float eps);
Tens upsample2x(Ctx& c, const Tens& x);
Tens concat_c(Ctx& c, const Tens& a, const Tens& b);
Tens from_nchw_f32(Ctx& c, const float* src, int N, int C, int H, int W, int Cp);
void to_nchw_f32(Ctx& c, const Tens& x, float* dst);
// out[r,:] = act_out(W act_in(x[r,:]) + b) (+ res[r,:]); any R (chunked by 16 rows internally)
void linear_small(Ctx& c, const float* x, long long ldx, int R, const Lin32& w, float* out, long long ldo, int act_in,
                  int act_out, const float* res = nullptr);
void timestep_embedding(Ctx& c, const long long* t, float* out, int N, int dim);
void concat9(Ctx& c, const float* x, const float* z, const float* mask, float* out, int B, int HW, int dup);
void cfg_ddim_update(Ctx& c, const float* x, const float* eps2, const float* noise, float* x_prev, float* pred_x0,
                     long long count, float scale, float a_t, float a_prev, float sigma, float sqrt_one_minus_at,
                     int has_uncond);

void eps_from_taps(Ctx& c, const float* taps, const float* bias, float* eps, int N, int L);
void taps_cfg_ddim_update(Ctx& c, const float* x, const float* taps, const float* bias, const float* noise, float* x_prev,
                          float* pred_x0, int B, int L, float scale, float a_t, float a_prev, float sigma,
                          float sqrt_one_minus_at, int has_uncond);
LinW pack_out_taps(Ctx& c, const std::string& wname);
void cfg_combine(Ctx& c, const float* eps2, float* out, long long count, float scale);
void plms_combine(Ctx& c, const float* e_t, const float* o1, const float* o2, const float* o3, const float* e_next,
                  float* out, long long count, int order);
void q_sample(Ctx& c, const float* x0, const float* noise, const float* coef_dev, float* out, long long per_sample, int B);

inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

}  // namespace rfb
