// extern "C" boundary (include/reface_b200.h).  Exceptions never cross it: every entry point returns an
// error code and stores the message in the context.
#include <cstdio>
#include "../../include/reface_b200.h"

#include <cudaTypedefs.h>

#include <cstring>

#include "models.h"

using namespace rfb;

struct rfb_ctx {
  Ctx c;
};

#define API_BEGIN(ctx_)                                 \
  if (!(ctx_)) return -1;                               \
  Ctx& c = (ctx_)->c;                                   \
  try {                                                 \
    CUDA_OK(cudaSetDevice(c.device));
#define API_END                                         \
    return 0;                                           \
  } catch (const std::exception& e) {                   \
    c.err = e.what();                                   \
    c.arena_off = 0;                                    \
    return -2;                                          \
  }

// A model owns the device buffers its builder packed (Ctx::dmalloc): they are moved from Ctx::owned into the model and
// freed with it, so rebuilding a model does not leak its previous ~GBs of packed fp16 weights.
template <class M>
static void drop_model(Ctx& c, M*& slot) {
  if (!slot) return;
  cudaDeviceSynchronize();
  c.drop_graphs();  // captured launches hold pointers into the packed weights
  for (void* p : slot->owned) cudaFree(p);
  delete slot;
  slot = nullptr;
}
template <class M, class F>
static void rebuild_model(Ctx& c, M*& slot, F build) {
  drop_model(c, slot);
  const size_t o0 = c.owned.size();
  M* m = build();
  m->owned.assign(c.owned.begin() + o0, c.owned.end());
  c.owned.resize(o0);
  slot = m;
}

extern "C" {

int rfb_init(int device, size_t arena_bytes, rfb_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  rfb_ctx* h = new rfb_ctx();
  Ctx& c = h->c;
  try {
    c.device = device;
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev <= 0)
      throw std::runtime_error("reface_b200 requires a CUDA device (sm_100a); none is visible -- there is no CPU fallback");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw std::runtime_error(std::string("reface_b200 kernels are built for sm_100a only; device is sm_") +
                               std::to_string(prop.major) + std::to_string(prop.minor));
    c.num_sms = prop.multiProcessorCount;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    c.encode_fn = fn;
    c.arena_cap = arena_bytes ? arena_bytes : (size_t)40 << 30;
    CUDA_OK(cudaMalloc((void**)&c.arena, c.arena_cap));
  } catch (const std::exception& e) {
    // keep the handle alive so that the caller can read the message
    c.err = e.what();
    *out = h;
    return -2;
  }
  *out = h;
  return 0;
}

void rfb_destroy(rfb_ctx* h) {
  if (!h) return;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  c.drop_graphs();
  if (c.gstream) cudaStreamDestroy(c.gstream);
  for (auto& kv : c.params) cudaFree(kv.second.f32);
  for (void* p : c.retired) cudaFree(p);
  for (void* p : c.owned) cudaFree(p);
  drop_model(c, c.unet);
  drop_model(c, c.vae);
  drop_model(c, c.clip);
  drop_model(c, c.arc);
  drop_model(c, c.parser);
  if (c.arena) cudaFree(c.arena);
  delete h;
}

const char* rfb_last_error(rfb_ctx* h) { return h ? h->c.err.c_str() : "null context"; }

int rfb_set_param(rfb_ctx* h, const char* name, const float* data, int ndim, const int64_t* shape) {
  API_BEGIN(h)
  Param p;
  p.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    p.shape.push_back(shape[i]);
    p.numel *= (size_t)shape[i];
  }
  // Built models keep raw pointers into the fp32 parameter storage (biases, norm affine vectors, the GEMV weights):
  // a re-registered key is therefore overwritten IN PLACE when its size is unchanged, and otherwise the old buffer is
  // retired (kept alive until rfb_destroy) instead of freed.  Packed fp16 copies are refreshed by the next rfb_build_*.
  c.drop_graphs();
  auto it = c.params.find(name);
  if (it != c.params.end() && it->second.numel == p.numel && it->second.f32) {
    CUDA_OK(cudaMemcpy(it->second.f32, data, p.numel * sizeof(float), cudaMemcpyDefault));
    it->second.shape = p.shape;
    return 0;
  }
  if (it != c.params.end()) {
    if (it->second.f32) c.retired.push_back(it->second.f32);
    c.params.erase(it);
  }
  CUDA_OK(cudaMalloc((void**)&p.f32, std::max<size_t>(p.numel * sizeof(float), 256)));
  CUDA_OK(cudaMemcpy(p.f32, data, p.numel * sizeof(float), cudaMemcpyDefault));
  c.params[name] = p;
  API_END
}
int rfb_has_param(rfb_ctx* h, const char* name) { return (h && h->c.has(name)) ? 1 : 0; }
long long rfb_release_packed_originals(rfb_ctx* h) {
  if (!h) return -1;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();  // pack kernels read the fp32 originals asynchronously
  long long freed = 0;
  for (auto& kv : c.params) {
    Param& p = kv.second;
    if (p.f32 && p.packed && !p.pinned && kv.first.rfind("__op.", 0) != 0) {
      cudaFree(p.f32);
      p.f32 = nullptr;
      freed += (long long)(p.numel * sizeof(float));
    }
  }
  return freed;
}

int rfb_build_unet(rfb_ctx* h, const char* prefix) {
  API_BEGIN(h)
  rebuild_model(c, c.unet, [&] { return build_unet(c, prefix, UNetCfg()); });
  API_END
}
int rfb_build_vae(rfb_ctx* h, const char* prefix) {
  API_BEGIN(h)
  rebuild_model(c, c.vae, [&] { return build_vae(c, prefix); });
  API_END
}
int rfb_build_clip(rfb_ctx* h, const char* prefix) {
  API_BEGIN(h)
  rebuild_model(c, c.clip, [&] { return build_clip(c, prefix); });
  API_END
}
int rfb_build_arcface(rfb_ctx* h, const char* prefix) {
  API_BEGIN(h)
  rebuild_model(c, c.arc, [&] { return build_arcface(c, prefix); });
  API_END
}

int rfb_build_face_parser(rfb_ctx* h, const char* prefix) {
  API_BEGIN(h)
  rebuild_model(c, c.parser, [&] { return build_face_parser(c, prefix); });
  API_END
}

int rfb_set_option(rfb_ctx* h, const char* key, long long value) {
  if (!h) return -1;
  Ctx& c = h->c;
  const std::string k = key;
  if (k != "profile") c.drop_graphs();  // captured launches bake the options in
  if (k == "use_graph") c.use_graph = (int)value;
  else if (k == "gemm_bn") c.force_bn = (int)value;
  else if (k == "gemm_stages") c.force_stages = (int)value;
  else if (k == "gemm_smem_budget") c.gemm_smem_budget = (int)value;
  else if (k == "attn_flash") c.attn_flash = (int)value;
  else if (k == "profile") c.profile = (int)value;
  else if (k == "gn_fused") c.gn_fused = (int)value;
  else if (k == "gn_epi_stats") c.gn_epi_stats = (int)value;
  else if (k == "gn_apply_bps") c.gn_apply_bps = (int)value;
  else if (k == "gn_fold") c.gn_fold = (int)value;
  else if (k == "cfg_share") c.cfg_share = (int)value;
  else if (k == "gn_cluster") c.gn_cluster = (int)value;
  else if (k == "gn_fused_max_elems") c.gn_fused_max_elems = value;
  else if (k == "gn_threads") c.gn_threads = (int)value;
  else if (k == "attn_poly") c.attn_poly = (int)value;
  else if (k == "attn_pingpong") c.attn_pingpong = (int)value;
  else if (k == "attn_pad") c.attn_pad = (int)value;
  else if (k == "attn_stagger") c.attn_stagger = (int)value;
  else if (k == "gemm_wave_bn") c.gemm_wave_bn = (int)value;
  else if (k == "gemm_splitk") c.gemm_splitk = (int)value;
  else if (k == "gemm_mcast") c.gemm_mcast = (int)value;
  else if (k == "gemm_lean") c.gemm_lean = (int)value;
  else if (k == "pdl") c.pdl = (int)value;
  else if (k == "gemm_mcast_big") c.gemm_mcast_big = (int)value;
  else if (k == "gemm_mcast_min_nk") c.gemm_mcast_min_nk = (int)value;
  else if (k == "conv_tma_stride2") c.conv_tma_stride2 = (int)value;
  else if (k == "ln_vec") c.ln_vec = (int)value;
  else if (k == "gemm_pair") c.gemm_pair = (int)value;
  else if (k == "gemm_kmerge") c.gemm_kmerge = (int)value;
  else if (k == "gemm_debug") c.gemm_debug = (int)value;
  else if (k == "gemm_pair_min_nk") c.gemm_pair_min_nk = (int)value;
  else if (k == "gemm_epi3_max_nk") c.gemm_epi3_max_nk = (int)value;
  else return -1;
  return 0;
}
long long rfb_launch_count(rfb_ctx* h) { return h ? h->c.launches : 0; }
long long rfb_graph_replays(rfb_ctx* h) { return h ? h->c.graph_replays : 0; }
int rfb_debug_read(rfb_ctx* h, unsigned long long* out, int n) {
  API_BEGIN(h)
  RFB_CHECK(c.dbg_buf, "option gemm_debug was never enabled");
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(out, c.dbg_buf, (size_t)std::min(n, c.num_sms * 8) * sizeof(unsigned long long),
                     cudaMemcpyDeviceToHost));
  API_END
}
int rfb_profile_read(rfb_ctx* h, double* ms, double* flops, double* flops_exec, long long* n) {
  API_BEGIN(h)
  CUDA_OK(cudaDeviceSynchronize());
  double tms = 0, tf = 0, te = 0;
  for (auto& r : c.prof) {
    float e = 0;
    CUDA_OK(cudaEventElapsedTime(&e, r.a, r.b));
    tms += e, tf += r.flops, te += r.flops_exec;
    if (c.profile >= 2)
      printf("PROF kind=%d M=%d N=%d K=%d BN=%d z=%d mode=%d us=%.2f tflops=%.1f\n", r.kind, r.M, r.N, r.K, r.BN, r.z, r.mode,
             e * 1e3, r.flops / (e * 1e-3) / 1e12);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (ms) *ms = tms;
  if (flops) *flops = tf;
  if (flops_exec) *flops_exec = te;
  if (n) *n = (long long)c.prof.size();
  c.prof.clear();
  API_END
}
size_t rfb_arena_peak(rfb_ctx* h) { return h ? h->c.arena_peak : 0; }

// ------------------------------------------------------------------------------------------ hot path
int rfb_unet_forward(rfb_ctx* h, const float* x9, const int64_t* t, const float* context, int N, int L, int T,
                     float* eps, void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.unet, "rfb_build_unet has not been called");
  c.stream = (cudaStream_t)stream;
  unet_forward(c, *c.unet, x9, (const long long*)t, context, N, L, T, eps);
  API_END
}
int rfb_concat9(rfb_ctx* h, const float* x, const float* z, const float* mask, int B, int L, int dup, float* out,
                void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  concat9(c, x, z, mask, out, B, L * L, dup);
  API_END
}
int rfb_cfg_ddim_update(rfb_ctx* h, const float* x, const float* eps2, const float* noise, long long count, float scale,
                        float a_t, float a_prev, float sigma, float sqrt_one_minus_at, int has_uncond, float* x_prev,
                        float* pred_x0, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  cfg_ddim_update(c, x, eps2, noise, x_prev, pred_x0, count, scale, a_t, a_prev, sigma, sqrt_one_minus_at, has_uncond);
  API_END
}
int rfb_ddim_sample(rfb_ctx* h, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                    const float* uncond, int B, int L, int T, const int64_t* timesteps, const float* a_t,
                    const float* a_prev, const float* sigma, const float* sqrt_one_minus_a, int n_steps, float cfg_scale,
                    const float* noise, int log_every_t, float* x0_out, float* inter_x, float* inter_pred_x0,
                    void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.unet, "rfb_build_unet has not been called");
  c.stream = (cudaStream_t)stream;
  DdimSchedule s;
  s.n = n_steps;
  s.timesteps.assign(timesteps, timesteps + n_steps);
  s.a_t.assign(a_t, a_t + n_steps);
  s.a_prev.assign(a_prev, a_prev + n_steps);
  s.sigma.assign(sigma, sigma + n_steps);
  s.sqrt_one_minus_a.assign(sqrt_one_minus_a, sqrt_one_minus_a + n_steps);
  ddim_sample(c, *c.unet, x_T, z_inpaint, mask, cond, uncond, B, L, T, s, cfg_scale, noise, x0_out, inter_x,
              inter_pred_x0, log_every_t);
  API_END
}
int rfb_plms_sample(rfb_ctx* h, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                    const float* uncond, int B, int L, int T, const int64_t* timesteps, const float* a_t,
                    const float* a_prev, const float* sigma, const float* sqrt_one_minus_a, int n_steps, float cfg_scale,
                    int log_every_t, float* x0_out, float* inter_x, float* inter_pred_x0, void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.unet, "rfb_build_unet has not been called");
  c.stream = (cudaStream_t)stream;
  DdimSchedule s;
  s.n = n_steps;
  s.timesteps.assign(timesteps, timesteps + n_steps);
  s.a_t.assign(a_t, a_t + n_steps);
  s.a_prev.assign(a_prev, a_prev + n_steps);
  s.sigma.assign(sigma, sigma + n_steps);
  s.sqrt_one_minus_a.assign(sqrt_one_minus_a, sqrt_one_minus_a + n_steps);
  for (int i = 0; i < n_steps; ++i) RFB_CHECK(sigma[i] == 0.0f, "ddim_eta must be 0 for PLMS");  // plms.py:25-26
  plms_sample(c, *c.unet, x_T, z_inpaint, mask, cond, uncond, B, L, T, s, cfg_scale, x0_out, inter_x, inter_pred_x0,
              log_every_t);
  API_END
}
int rfb_q_sample(rfb_ctx* h, const float* x_start, const float* noise, const float* coef, int B, long long per_sample,
                 float* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  float* cd = c.alloc_t<float>((size_t)2 * B);
  CUDA_OK(cudaMemcpyAsync(cd, coef, (size_t)2 * B * sizeof(float), cudaMemcpyHostToDevice, c.stream));
  q_sample(c, x_start, noise, cd, out, per_sample, B);
  CUDA_OK(cudaStreamSynchronize(c.stream));  // `coef` is a host array owned by the caller
  c.release(mk);
  API_END
}
int rfb_face_parse(rfb_ctx* h, const float* img01, int B, int H, int W, float* logits8, uint8_t* seg19, uint8_t* seg12,
                   void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.parser, "rfb_build_face_parser has not been called");
  c.stream = (cudaStream_t)stream;
  face_parse(c, *c.parser, img01, B, H, W, logits8, seg19, seg12);
  API_END
}
int rfb_inpaint_from_parsing(rfb_ctx* h, const float* img, const uint8_t* seg12, const int* remove, int n_remove, int B,
                             int H, int W, float* mask, float* inpaint, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  unsigned bits = 0;
  for (int i = 0; i < n_remove; ++i) {
    RFB_CHECK(remove[i] >= 0 && remove[i] < 32, "label out of range");
    bits |= 1u << remove[i];
  }
  inpaint_from_parsing(c, img, seg12, bits, B, H, W, mask, inpaint);
  API_END
}
int rfb_paste_back(rfb_ctx* h, const float* x01, const uint8_t* orig, const double* coeffs, int B, int hh, int ww, int up,
                   int H, int W, uint8_t* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  paste_back(c, x01, orig, coeffs, B, hh, ww, up, H, W, out);
  API_END
}
int rfb_vae_encode(rfb_ctx* h, const float* img, const float* noise, int B, int H, int W, double scale_factor, float* z,
                   float* mean, float* logvar, void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.vae, "rfb_build_vae has not been called");
  c.stream = (cudaStream_t)stream;
  vae_encode(c, *c.vae, img, noise, B, H, W, (float)scale_factor, z, mean, logvar);
  API_END
}
int rfb_vae_decode(rfb_ctx* h, const float* z, int B, int hh, int ww, double scale_factor, float* img, void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.vae, "rfb_build_vae has not been called");
  RFB_CHECK(scale_factor != 0.0, "scale_factor must be non-zero");
  c.stream = (cudaStream_t)stream;
  // ddpm.py:1284 `z = 1. / self.scale_factor * z`: the Python double 1/s is rounded to fp32 by the tensor multiply
  vae_decode(c, *c.vae, z, B, hh, ww, (float)(1.0 / scale_factor), img);
  API_END
}
int rfb_clip_encode(rfb_ctx* h, const float* img224, int B, float* out768, void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.clip, "rfb_build_clip has not been called");
  c.stream = (cudaStream_t)stream;
  clip_embed(c, *c.clip, img224, B, out768);
  API_END
}
int rfb_arcface_embed(rfb_ctx* h, const float* img, int B, float* out512, void* stream) {
  API_BEGIN(h)
  RFB_CHECK(c.arc, "rfb_build_arcface has not been called");
  c.stream = (cudaStream_t)stream;
  arcface_embed(c, *c.arc, img, B, out512);
  API_END
}

// c = (w_clip*(Ps(clip_src)+Pt(clip_tgt)) + w_id*Pid(id) + w_lm*Plm(lm)) / (w_clip+w_id+w_lm)   [ddpm.py:904-915,1010-1039]
__global__ void fuse_cond_kernel(const float* a, const float* b, const float* id, const float* lm, float* out, int n,
                                 float wc, float wi, float wl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ((a[i] + b[i]) * wc + id[i] * wi + lm[i] * wl) / (wc + wi + wl);
}
int rfb_landmark_project(rfb_ctx* h, const float* lm136, int B, float* out768, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  Lin32 pl = lin32(c, "landmark_proj_out.weight", "landmark_proj_out.bias");
  linear_small(c, lm136, 136, B, pl, out768, 768, 0, 0);
  API_END
}
int rfb_condition_fuse(rfb_ctx* h, const float* clip_src, const float* clip_tgt, const float* id_feat,
                       const float* lm136, const float* lm_proj768, int B, float w_clip, float w_id, float w_lm,
                       float* c_out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  RFB_CHECK((lm136 != nullptr) != (lm_proj768 != nullptr), "pass exactly one of lm136 (raw points) / lm_proj768 (projected)");
  const size_t mk = c.mark();
  float* t = c.alloc_t<float>((size_t)4 * B * 768);
  Lin32 ps = lin32(c, "proj_out_source.weight", "proj_out_source.bias");
  Lin32 pt = lin32(c, "proj_out_target.weight", "proj_out_target.bias");
  Lin32 pi = lin32(c, "ID_proj_out.weight", "ID_proj_out.bias");
  Lin32 pl = lin32(c, "landmark_proj_out.weight", "landmark_proj_out.bias");
  for (int r0 = 0; r0 < B; r0 += 16) {
    const int R = std::min(16, B - r0);
    linear_small(c, clip_src + (size_t)r0 * 768, 768, R, ps, t + (size_t)r0 * 768, 768, 0, 0);
    linear_small(c, clip_tgt + (size_t)r0 * 768, 768, R, pt, t + (size_t)(B + r0) * 768, 768, 0, 0);
    linear_small(c, id_feat + (size_t)r0 * 512, 512, R, pi, t + (size_t)(2 * B + r0) * 768, 768, 0, 0);
    if (lm136) linear_small(c, lm136 + (size_t)r0 * 136, 136, R, pl, t + (size_t)(3 * B + r0) * 768, 768, 0, 0);
  }
  const int n = B * 768;
  fuse_cond_kernel<<<(n + 255) / 256, 256, 0, c.stream>>>(t, t + n, t + 2 * n, lm136 ? t + 3 * n : lm_proj768, c_out, n, w_clip,
                                                          w_id, w_lm);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  c.release(mk);
  API_END
}

// ((tar+1)/2 - mean)/std, then bilinear resize to 224 without antialiasing, align_corners=False [ddpm.py:907-912]
__global__ void target_clip_input_kernel(const float* __restrict__ tar, float* __restrict__ out, int B, int H, int W) {
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  const long long total = (long long)B * 3 * 224 * 224;
  const float sy = (float)H / 224.0f, sx = (float)W / 224.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % 224), oy = (int)((i / 224) % 224);
    const int ch = (int)((i / (224 * 224)) % 3);
    const long long b = i / (3 * 224 * 224);
    float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* p = tar + (b * 3 + ch) * (long long)H * W;
    auto nrm = [&](float v) { return ((v + 1.0f) * 0.5f - mean[ch]) / stdv[ch]; };
    const float v00 = nrm(p[(long long)y0 * W + x0]), v01 = nrm(p[(long long)y0 * W + x1]);
    const float v10 = nrm(p[(long long)y1 * W + x0]), v11 = nrm(p[(long long)y1 * W + x1]);
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}
int rfb_target_clip_input(rfb_ctx* h, const float* tar, int B, int H, int W, float* out224, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  target_clip_input_kernel<<<grid_for((long long)B * 3 * 224 * 224), 256, 0, c.stream>>>(tar, out224, B, H, W);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  API_END
}

// ------------------------------------------------------------------------------------------ single ops (tests)
struct TempParams {  // registers temporaries under "__op." and frees everything packed from them on exit
  Ctx& c;
  size_t owned0;
  std::vector<std::string> names;
  explicit TempParams(Ctx& c_) : c(c_), owned0(c_.owned.size()) {}
  void add(const std::string& n, const float* dev, std::vector<int64_t> shape) {
    Param p;
    p.shape = shape;
    p.numel = 1;
    for (auto s : shape) p.numel *= (size_t)s;
    p.f32 = const_cast<float*>(dev);
    c.params[n] = p;
    names.push_back(n);
  }
  ~TempParams() {
    cudaStreamSynchronize(c.stream);
    for (auto& n : names) c.params.erase(n);
    for (size_t i = owned0; i < c.owned.size(); ++i) cudaFree(c.owned[i]);
    c.owned.resize(owned0);
  }
};

// [rows, 3*heads*d] -> [rows, 3*heads*hs] with zero-filled head slices (test hook for the padded attention layout)
__global__ void pad_heads_kernel(const __half* __restrict__ src, __half* __restrict__ dst, long long rows, int nh, int d,
                                 int hs) {
  const long long total = rows * nh * hs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % hs);
    const long long t = i / hs;
    const int h = (int)(t % nh);
    const long long r = t / nh;
    dst[i] = e < d ? src[(r * nh + h) * d + e] : __float2half_rn(0.f);
  }
}

__global__ void f32_to_f16_kernel(const float* s, __half* d, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = __float2half_rn(s[i]);
}
__global__ void f16_to_f32_kernel(const __half* s, float* d, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = __half2float(s[i]);
}

int rfb_op_linear(rfb_ctx* h, const float* x, const float* w, const float* bias, const float* residual, long long M,
                  int K, int N, int act, int geglu, const float* x2, int K2, long long M2, float* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  {
    TempParams tp(c);
    tp.add("__op.w", w, {N, K + (x2 ? K2 : 0)});
    if (bias) tp.add("__op.b", bias, {N});
    const int NO = geglu ? N / 2 : N;
    Tens xt = c.new_tens(1, 1, (int)M, K);
    f32_to_f16_kernel<<<grid_for(M * K), 256, 0, c.stream>>>(x, xt.p, M * K);
    Epi e;
    e.act = act, e.geglu = geglu;
    __half* r16 = nullptr;
    if (residual) {
      r16 = c.alloc_t<__half>((size_t)M * NO);
      f32_to_f16_kernel<<<grid_for(M * NO), 256, 0, c.stream>>>(residual, r16, M * NO);
      e.res = r16, e.ldr = NO;
    }
    LinW lw;
    if (geglu) {
      RFB_CHECK(bias && !x2, "geglu needs a bias and a single source");
      int bn = 256;
      while (N % bn) bn /= 2;
      lw = pack_geglu(c, "__op.w", "__op.b", bn);
    } else {
      lw = pack_linear(c, "__op.w", bias ? "__op.b" : "");
    }
    Tens y;
    if (x2) {  // [x | x2] W^T through two TMA descriptors (x2 has M2 rows, read modulo)
      Tens x2t = c.new_tens(1, 1, (int)M2, K2);
      f32_to_f16_kernel<<<grid_for(M2 * K2), 256, 0, c.stream>>>(x2, x2t.p, M2 * K2);
      y = c.new_tens(1, 1, (int)M, N);
      e.bias = lw.b;
      gemm2(c, xt.p, K, K, x2t.p, K2, K2, M2 == M ? 0 : M2, M, lw.w, lw.kp, N, y.p, N, e);
    } else {
      y = linear_t(c, xt, lw, e);
    }
    f16_to_f32_kernel<<<grid_for(M * NO), 256, 0, c.stream>>>(y.p, out, M * NO);
    CUDA_OK(cudaGetLastError());
  }
  c.release(mk);
  API_END
}

int rfb_op_upconv(rfb_ctx* h, const float* x, const float* w, const float* bias, int N, int C, int H, int W, int O,
                  float* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  {
    TempParams tp(c);
    tp.add("__op.w", w, {O, C, 3, 3});
    if (bias) tp.add("__op.b", bias, {O});
    ConvW cw = pack_upconv(c, "__op.w", bias ? "__op.b" : "");
    Tens xt = from_nchw_f32(c, x, N, C, H, W, C);
    Tens y = upconv3x3_t(c, xt, cw, Epi());
    to_nchw_f32(c, y, out);
  }
  c.release(mk);
  API_END
}

int rfb_op_conv2d(rfb_ctx* h, const float* x, const float* w, const float* bias, int N, int C, int H, int W, int O,
                  int ksz, int stride, int pad_t, int pad_l, int pad_b, int pad_r, const float* gn_gamma,
                  const float* gn_beta, float* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  {
    TempParams tp(c);
    tp.add("__op.w", w, {O, C, ksz, ksz});
    if (bias) tp.add("__op.b", bias, {O});
    ConvW cw = pack_conv(c, "__op.w", bias ? "__op.b" : "");
    Tens xt = from_nchw_f32(c, x, N, C, H, W, C);
    Epi e;
    e.want_stats = gn_gamma != nullptr;  // conv -> GroupNorm(32) + SiLU chain: statistics from the conv epilogue
    Tens y = conv3x3_t(c, xt, cw, e, stride, pad_t, pad_l, pad_b, pad_r);
    if (gn_gamma) y = groupnorm(c, y, gn_gamma, gn_beta, 1e-5f, true);
    to_nchw_f32(c, y, out);
  }
  c.release(mk);
  API_END
}

int rfb_op_groupnorm(rfb_ctx* h, const float* x, const float* gamma, const float* beta, int N, int C, int H, int W,
                     float eps, int silu, const float* x2, int C2, int N2, float* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  Tens xt = from_nchw_f32(c, x, N, C, H, W, C);
  Tens x2t;
  if (x2) x2t = from_nchw_f32(c, x2, N2, C2, H, W, C2);   // GroupNorm over the channel concatenation [x | x2]
  Tens y = groupnorm2(c, xt, x2t, gamma, beta, eps, silu != 0);
  to_nchw_f32(c, y, out);
  c.release(mk);
  API_END
}

int rfb_op_layernorm(rfb_ctx* h, const float* x, const float* gamma, const float* beta, long long rows, int C, float eps,
                     float* out, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  Tens xt = c.new_tens(1, 1, (int)rows, C);
  f32_to_f16_kernel<<<grid_for(rows * C), 256, 0, c.stream>>>(x, xt.p, rows * C);
  Tens y = layernorm(c, xt, gamma, beta, eps);
  f16_to_f32_kernel<<<grid_for(rows * C), 256, 0, c.stream>>>(y.p, out, rows * C);
  CUDA_OK(cudaGetLastError());
  c.release(mk);
  API_END
}

int rfb_op_attention(rfb_ctx* h, const float* qkv, int N, int L, int heads, int d, float scale, float* out,
                     void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  const int C = heads * d;
  const long long n_in = (long long)N * L * 3 * C, n_out = (long long)N * L * C;
  __half* q16 = c.alloc_t<__half>(n_in);
  __half* o16 = c.alloc_t<__half>(n_out);
  f32_to_f16_kernel<<<grid_for(n_in), 256, 0, c.stream>>>(qkv, q16, n_in);
  if (c.attn_pad && d < 64) {
    const int hs = 64;
    __half* qp = c.alloc_t<__half>((size_t)N * L * 3 * heads * hs);
    pad_heads_kernel<<<grid_for((long long)N * L * 3 * heads * hs), 256, 0, c.stream>>>(q16, qp, (long long)N * L, 3 * heads, d,
                                                                                       hs);
    attention(c, qp, 3 * heads * hs, N, L, heads, d, o16, C, scale, 0, heads * hs, 2 * heads * hs, hs);
  } else {
    attention(c, q16, 3 * C, N, L, heads, d, o16, C, scale, 0, C, 2 * C);
  }
  f16_to_f32_kernel<<<grid_for(n_out), 256, 0, c.stream>>>(o16, out, n_out);
  CUDA_OK(cudaGetLastError());
  c.release(mk);
  API_END
}

/* Kernel-only timing of the HBM-bound normalisation ops on device-resident fp16 tensors (CUDA events on the
 * launching stream, `iters` back-to-back launches after one warm-up); kind 0 = GroupNorm(32)+SiLU, 1 = LayerNorm,
 * 2 = GroupNorm(32)+SiLU fed by producer-epilogue statistics (finalize + one streaming pass). */
int rfb_bench_norm(rfb_ctx* h, int kind, int N, int C, int H, int W, int iters, double* ms_per_launch, void* stream) {
  API_BEGIN(h)
  c.stream = (cudaStream_t)stream;
  const size_t mk = c.mark();
  Tens x = c.new_tens(N, H, W, C);
  CUDA_OK(cudaMemsetAsync(x.p, 0x3c, (size_t)x.rows() * C * sizeof(__half), c.stream));  // 0x3c3c = 1.0586
  float* gb = c.alloc_t<float>((size_t)2 * C);
  CUDA_OK(cudaMemsetAsync(gb, 0, (size_t)2 * C * sizeof(float), c.stream));
  if (kind == 2) {  // GroupNorm fed by producer statistics (finalize + streaming apply): any partial sums will do for timing
    x.stats = c.alloc_t<float>((size_t)(x.rows() / 32) * C * 2);
    CUDA_OK(cudaMemsetAsync(x.stats, 0, (size_t)(x.rows() / 32) * C * 2 * sizeof(float), c.stream));
  }
  cudaEvent_t a, b;
  CUDA_OK(cudaEventCreate(&a));
  CUDA_OK(cudaEventCreate(&b));
  for (int it = -1; it < iters; ++it) {
    if (it == 0) CUDA_OK(cudaEventRecord(a, c.stream));
    const size_t m2 = c.mark();
    if (kind == 0 || kind == 2) groupnorm(c, x, gb, gb + C, 1e-5f, true);
    else layernorm(c, x, gb, gb + C, 1e-5f);
    c.release(m2);
  }
  CUDA_OK(cudaEventRecord(b, c.stream));
  CUDA_OK(cudaEventSynchronize(b));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, a, b));
  CUDA_OK(cudaEventDestroy(a));
  CUDA_OK(cudaEventDestroy(b));
  *ms_per_launch = (double)ms / iters;
  c.release(mk);
  API_END
}

}  // extern "C"
