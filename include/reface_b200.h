/* reface_b200 -- C ABI of the B200-native REFace inference hot path.
 *
 * The reference (Sanoojan/REFace) is pure Python/PyTorch and has no FFI boundary of its own; its plugin
 * mechanism is ldm/util.py:78-93 `instantiate_from_config` (YAML `target:` -> constructor).  The Python
 * classes in reface_b200/ are the drop-in targets for those slots and call this library through ctypes.
 * Every entry point below names the reference interface it replaces (file:line relative to the reference).
 *
 * Conventions: plain pointers and sizes only, no torch types.  All tensor pointers are DEVICE pointers
 * (fp32, contiguous, NCHW for images/latents, row-major otherwise) unless marked [host].  The caller owns
 * every buffer; the library borrows them for the duration of the call.  Work is enqueued on `stream`
 * (a cudaStream_t passed as void*) and is asynchronous w.r.t. the host.  Return value: 0 = OK, negative =
 * error (message via rfb_last_error).  One context per GPU; a context is not re-entrant.
 */
#ifndef REFACE_B200_H
#define REFACE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rfb_ctx rfb_ctx;

/* ---- context / weights ------------------------------------------------------------------------ */
/* One context per device: owns weights, the activation arena (arena_bytes; 0 = default) and workspaces. */
int rfb_init(int device, size_t arena_bytes, rfb_ctx** out);
void rfb_destroy(rfb_ctx* ctx);
const char* rfb_last_error(rfb_ctx* ctx);
/* Replaces model.load_state_dict(sd) (scripts/inference_test_bench.py:98-103): one call per state-dict
 * entry, keyed by the reference's own key (e.g. "model.diffusion_model.input_blocks.0.0.weight").
 * `data` may be a host or a device pointer (fp32). */
int rfb_set_param(rfb_ctx* ctx, const char* name, const float* data, int ndim, const int64_t* shape);
int rfb_has_param(rfb_ctx* ctx, const char* name);
/* After the rfb_build_* calls: frees the fp32 device copies of every parameter that only served as the source of a
 * packed fp16 GEMM operand (conv / linear weights: ~5 GB of the ~8 GB a full REFace checkpoint occupies); biases, norm
 * vectors and the small fp32 GEMV weights the networks keep reading stay.  Returns the bytes freed.  Rebuilding a
 * network afterwards requires registering its weights again (rfb_set_param). */
long long rfb_release_packed_originals(rfb_ctx* ctx);
/* Build the packed (fp16, implicit-GEMM layout) networks from the registered parameters.  `prefix` is the
 * state-dict prefix ("model.diffusion_model.", "first_stage_model.", "cond_stage_model.",
 * "face_ID_model.facenet.").  Replace the constructors reached through instantiate_from_config:
 * UNetModel (openaimodel.py:558), AutoencoderKL (autoencoder.py:285), FrozenCLIPEmbedder
 * (encoders/modules.py:211), Backbone (model_irse.py:10). */
int rfb_build_unet(rfb_ctx* ctx, const char* prefix);
int rfb_build_vae(rfb_ctx* ctx, const char* prefix);
int rfb_build_clip(rfb_ctx* ctx, const char* prefix);
int rfb_build_arcface(rfb_ctx* ctx, const char* prefix);
/* BiSeNet(n_classes=19) of pretrained/face_parsing/model.py:214-239 (keys below `prefix`, e.g. "face_parser.seg."). */
int rfb_build_face_parser(rfb_ctx* ctx, const char* prefix);
/* Tunables ("gemm_bn", "gemm_stages", "gemm_smem_budget", "attn_flash", "profile"); returns 0 if known. */
int rfb_set_option(rfb_ctx* ctx, const char* key, long long value);
long long rfb_launch_count(rfb_ctx* ctx); /* kernels launched by this library so far (graph replays count their kernels) */
/* rfb_ddim_sample captures its whole S-step loop into a CUDA graph the second time it sees a (shape, schedule, scale)
 * and replays it afterwards (option "use_graph", default 1); this counts the replays. */
long long rfb_graph_replays(rfb_ctx* ctx);
/* With option "profile"=1 every tensor-core (tcgen05 GEMM / implicit-GEMM conv / attention) launch is bracketed
 * by CUDA events on the launching stream; this drains them: summed device time [ms], summed ALGORITHMIC FLOPs
 * (2*M*N*K of the reference-equivalent contraction, no padding: a folded upsample convolution counts the 9-tap
 * convolution at output resolution that it replaces), summed EXECUTED FLOPs and the number of launches. */
int rfb_profile_read(rfb_ctx* ctx, double* ms, double* flops, double* flops_executed, long long* n_launches);
/* With option "gemm_debug"=1 the 2-CTA GEMM records per-CTA clock64 totals (8 x u64 per CTA: MMA-thread total, wait
 * on operands, wait on accumulator drain, producer wait, epilogue wait, epilogue total, k-stages, tiles). [host] out. */
int rfb_debug_read(rfb_ctx* ctx, unsigned long long* out, int n_u64);
size_t rfb_arena_peak(rfb_ctx* ctx);

/* ---- hot path ---------------------------------------------------------------------------------- */
/* UNetModel.forward (openaimodel.py:860-907) behind LatentDiffusion.apply_model (ddpm.py:1519-1617):
 * x9 [N,9,L,L], t [N] int64, context [N,T,768] -> eps [N,4,L,L]. */
int rfb_unet_forward(rfb_ctx* ctx, const float* x9, const int64_t* t, const float* context, int N, int L, int T,
                     float* eps, void* stream);
/* torch.cat([x, inpaint_image, inpaint_mask], 1) and the CFG duplication (ddim.py:330,338): out [dup*B,9,L,L]. */
int rfb_concat9(rfb_ctx* ctx, const float* x, const float* z_inpaint, const float* mask, int B, int L, int dup,
                float* out, void* stream);
/* CFG combine + DDIM x_{t-1} (ddim.py:346,363-374).  eps2 = [e_uncond ; e_cond] ([2B,4,L,L]) when has_uncond,
 * else [B,4,L,L].  noise may be NULL (eta = 0).  pred_x0 may be NULL. */
int rfb_cfg_ddim_update(rfb_ctx* ctx, const float* x, const float* eps2, const float* noise, long long count,
                        float scale, float a_t, float a_prev, float sigma, float sqrt_one_minus_at, int has_uncond,
                        float* x_prev, float* pred_x0, void* stream);
/* DDIMSampler.ddim_sampling with test_model_kwargs (ddim.py:200-251 driving p_sample_ddim :323-375).
 * Schedule tables are [host] arrays of length n_steps indexed by DDIM index (computed by the caller exactly
 * as make_schedule does, ddim.py:110-139).  cond/uncond [B,T,768]; uncond NULL or scale==1 disables CFG.
 * noise: NULL or [n_steps,B,4,L,L].  inter_x/inter_pred_x0: NULL or room for every logged step
 * (index % log_every_t == 0 or index == n_steps-1, ddim.py:247-249), in loop order. */
int rfb_ddim_sample(rfb_ctx* ctx, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                    const float* uncond, int B, int L, int T, const int64_t* timesteps, const float* a_t,
                    const float* a_prev, const float* sigma, const float* sqrt_one_minus_a, int n_steps, float cfg_scale,
                    const float* noise, int log_every_t, float* x0_out, float* inter_x, float* inter_pred_x0,
                    void* stream);
/* PLMSSampler.plms_sampling with test_model_kwargs (ldm/models/diffusion/plms.py:116-242): pseudo improved Euler on
 * the first step (two UNet evaluations), then Adams-Bashforth of order 2-4 over the last three eps.  Same argument
 * meaning as rfb_ddim_sample; every sigma must be 0 (plms.py:25-26 raises for eta != 0). */
int rfb_plms_sample(rfb_ctx* ctx, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                    const float* uncond, int B, int L, int T, const int64_t* timesteps, const float* a_t,
                    const float* a_prev, const float* sigma, const float* sqrt_one_minus_a, int n_steps, float cfg_scale,
                    int log_every_t, float* x0_out, float* inter_x, float* inter_pred_x0, void* stream);
/* DDPM.q_sample (ldm/models/diffusion/ddpm.py:412-415), the --Start_from_target initialisation of
 * scripts/inference_test_bench.py:414-435: out[b] = coef[2b] * x_start[b] + coef[2b+1] * noise[b] with
 * coef = (sqrt_alphas_cumprod[t_b], sqrt_one_minus_alphas_cumprod[t_b]) ([host], 2*B floats). */
int rfb_q_sample(rfb_ctx* ctx, const float* x_start, const float* noise, const float* coef, int B, long long per_sample,
                 float* out, void* stream);
/* Face parsing, the step before the swap path.  FaceParser.forward for an already 512-sized input
 * (pretrained/face_parsing/face_parsing_demo.py:266-281: clamp, ImageNet normalisation, BiSeNet, argmax) and
 * __ffhq_masks_to_faceParser_mask_detailed (:74-122).  img01 [B,3,H,W] fp32 in [0,1], H and W multiples of 32.
 * Outputs (each may be NULL): logits8 fp32 [B,19,H/8,W/8] (before the bilinear align_corners upsampling),
 * seg19 / seg12 uint8 [B,H,W]. */
int rfb_face_parse(rfb_ctx* ctx, const float* img01, int B, int H, int W, float* logits8, uint8_t* seg19, uint8_t* seg12,
                   void* stream);
/* mask = 1 - isin(seg12, remove), inpaint = img * mask (ldm/data/video_swap_dataset.py:150-222).  img [B,3,H,W] fp32,
 * remove: [host] label list (project_ffhq.yaml remove_mask_tar_FFHQ), mask [B,1,H,W], inpaint [B,3,H,W] (may be NULL). */
int rfb_inpaint_from_parsing(rfb_ctx* ctx, const float* img, const uint8_t* seg12, const int* remove, int n_remove, int B,
                             int H, int W, float* mask, float* inpaint, void* stream);
/* Paste-back of scripts/inference_swap_video.py:702-721 without the PIL round trip, bit exact with Pillow:
 * uint8 conversion of the decoded face x01 [B,3,h,w] (fp32 in [0,1], device), Image.resize((up, up), BILINEAR) (up = 0: no
 * resize), Image.transform(frame size, PERSPECTIVE, coeffs, BILINEAR) and alpha_composite over orig [B,H,W,3] (uint8,
 * device).  coeffs: [host] [B,8] doubles, the inverse perspective coefficients the reference stores per frame
 * (inference_swap_video.py:495-499,713).  out [B,H,W,3] uint8 (device). */
int rfb_paste_back(rfb_ctx* ctx, const float* x01, const uint8_t* orig, const double* coeffs, int B, int h, int w, int up,
                   int H, int W, uint8_t* out, void* stream);
/* get_first_stage_encoding(encode_first_stage(x)) (ddpm.py:1402-1439, 850-857; autoencoder.py:324-328;
 * distributions.py:24-37): img [B,3,H,W] -> z = scale_factor*(mean + std*noise) [B,4,H/8,W/8] (scale_factor =
 * LatentDiffusion.scale_factor, 0.18215 in project_ffhq.yaml:20; 1.0 returns the bare posterior.sample()).
 * noise NULL => mode() (z = scaled mean).  mean/logvar outputs optional (NULL). */
int rfb_vae_encode(rfb_ctx* ctx, const float* img, const float* noise, int B, int H, int W, double scale_factor, float* z,
                   float* mean, float* logvar, void* stream);
/* decode_first_stage (ddpm.py:1277-1337; autoencoder.py:330-333): z [B,4,h,w] -> `1/scale_factor * z` (the Python
 * double rounded to fp32, ddpm.py:1284) -> post_quant_conv -> Decoder -> img [B,3,8h,8w]; scale_factor 1.0 is
 * AutoencoderKL.decode itself. */
int rfb_vae_decode(rfb_ctx* ctx, const float* z, int B, int h, int w, double scale_factor, float* img, void* stream);
/* FrozenCLIPEmbedder.encode (encoders/modules.py:253-264): img [B,3,224,224] -> [B,1,768]. */
int rfb_clip_encode(rfb_ctx* ctx, const float* img224, int B, float* out768, void* stream);
/* IDLoss.extract_feats(x)[0] (ddpm.py:112-124 + model_irse.py:44-69): CLIP-normalised [B,3,224,224] -> [B,512]. */
int rfb_arcface_embed(rfb_ctx* ctx, const float* img224_clipnorm, int B, float* out512, void* stream);
/* LatentDiffusion.conditioning_with_feat (ddpm.py:872-1045) for the shipped config:
 * c = (w_clip*(proj_src(clip_src)+proj_tgt(clip_tgt)) + w_id*ID_proj(id) + w_lm*LM) / sum(w).
 * clip_src/clip_tgt [B,768], id_feat [B,512] -> c [B,1,768].  The landmark term LM is given EITHER as raw dlib points
 * lm136 [B,136] (projected here by landmark_proj_out, ddpm.py:1096) OR already projected, lm_proj768 [B,768] -- what
 * the reference passes: conditioning_with_feat(ref, landmarks=model.get_landmarks(x), tar=x)
 * (scripts/inference_test_bench.py:447-448, ddpm.py:1061-1062).  Exactly one of the two must be non-NULL.  Uses the
 * registered proj_out_source / proj_out_target / ID_proj_out / landmark_proj_out parameters. */
int rfb_condition_fuse(rfb_ctx* ctx, const float* clip_src, const float* clip_tgt, const float* id_feat,
                       const float* lm136, const float* lm_proj768, int B, float w_clip, float w_id, float w_lm,
                       float* c_out, void* stream);
/* The device half of LatentDiffusion.get_landmarks (ddpm.py:1068-1099): landmark_proj_out(lm136) [B,136] -> [B,768];
 * rows of zeros are the reference's no-face branch (ddpm.py:1081-1083).  Detection itself (dlib, CPU) stays with the
 * caller. */
int rfb_landmark_project(rfb_ctx* ctx, const float* lm136, int B, float* out768, void* stream);
/* tar -> un_norm -> CLIP normalise -> bilinear 224 (ddpm.py:907-912): [B,3,H,W] in [-1,1] -> [B,3,224,224]. */
int rfb_target_clip_input(rfb_ctx* ctx, const float* tar, int B, int H, int W, float* out224, void* stream);

/* ---- single-op entry points (kernel-level parity tests; fp32 in/out, converted internally) ------ */
/* x2 != NULL: out = [x | x2] w^T with x2 [M2,K2] read at row (m mod M2) -- the channel concatenation of the UNet skip
 * connections (openaimodel.py:897-899) as one K loop over two TMA descriptors; w is [N, K+K2]. */
int rfb_op_linear(rfb_ctx* ctx, const float* x, const float* w, const float* bias, const float* residual, long long M,
                  int K, int N, int act, int geglu, const float* x2, int K2, long long M2, float* out, void* stream);
/* Upsample.forward (openaimodel.py:109-119): conv3x3(nearest_2x(x)), x [N,C,H,W], w [O,C,3,3] -> [N,O,2H,2W]. */
int rfb_op_upconv(rfb_ctx* ctx, const float* x, const float* w, const float* bias, int N, int C, int H, int W, int O,
                  float* out, void* stream);
/* gn_gamma/gn_beta != NULL: followed by GroupNorm(32, eps 1e-5) + SiLU (ResBlock in_layers / out_layers,
 * openaimodel.py:202-206,226-230) whose statistics come out of the convolution's epilogue. */
int rfb_op_conv2d(rfb_ctx* ctx, const float* x, const float* w, const float* bias, int N, int C, int H, int W, int O,
                  int ksz, int stride, int pad_t, int pad_l, int pad_b, int pad_r, const float* gn_gamma,
                  const float* gn_beta, float* out, void* stream);
/* x2 != NULL: GroupNorm(32) over the channel concatenation [x | x2], x2 [N2,C2,H,W] read at sample (n mod N2). */
int rfb_op_groupnorm(rfb_ctx* ctx, const float* x, const float* gamma, const float* beta, int N, int C, int H, int W,
                     float eps, int silu, const float* x2, int C2, int N2, float* out, void* stream);
int rfb_op_layernorm(rfb_ctx* ctx, const float* x, const float* gamma, const float* beta, long long rows, int C,
                     float eps, float* out, void* stream);
int rfb_op_attention(rfb_ctx* ctx, const float* qkv, int N, int L, int heads, int d, float scale, float* out,
                     void* stream);

/* Kernel-only timing of the HBM-bound normalisation kernels (kind 0: GroupNorm(32)+SiLU as in ResBlock.in_layers,
 * openaimodel.py:203-206; kind 1: LayerNorm as in BasicTransformerBlock, attention.py:232-234) on device-resident
 * NHWC fp16 tensors [N,H,W,C]; CUDA events on `stream`. */
int rfb_bench_norm(rfb_ctx* ctx, int kind, int N, int C, int H, int W, int iters, double* ms_per_launch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REFACE_B200_H */
