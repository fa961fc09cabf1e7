// Model graphs on top of engine.h: UNet + DDIM loop, AutoencoderKL, CLIP ViT-L/14 vision + mapper, ArcFace.
#pragma once
#include <cmath>

#include "engine.h"

namespace rfb {

struct UNetCfg {
  int in_channels = 9, out_channels = 4, model_channels = 320, num_res_blocks = 2, num_heads = 8, context_dim = 768;
  std::vector<int> attention_resolutions = {4, 2, 1};
  std::vector<int> channel_mult = {1, 2, 4, 4};
};

struct ResW {
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  ConvW c1, c2, skipw;
  Lin32 emb;
  bool skip = false;
  int cin = 0, cout = 0, emb_off = 0;  // emb_off: column of this block in the concatenated emb_layers output
};
struct STW {
  const float *gn_g = nullptr, *gn_b = nullptr, *ln1g = nullptr, *ln1b = nullptr, *ln2g = nullptr, *ln2b = nullptr,
              *ln3g = nullptr, *ln3b = nullptr;
  ConvW proj_in, proj_out;
  const float *proj_in_w32 = nullptr, *proj_in_b = nullptr;  // fp32 originals for the folded GroupNorm (engine.cu)
  LinW qkv, o1, ff1, ff2, q2, kv2, o2h;
  Lin32 v2, o2, k2;
  int c = 0, heads = 0, d = 0, ctx_dim = 0, ff_bn = 256;
};
enum UOpKind { OP_CONV_IN, OP_RES, OP_ATTN, OP_DOWN, OP_UP };
struct UOp {
  int kind = OP_RES;
  ResW res;
  STW st;
  ConvW conv;
};
struct UNet {
  std::vector<void*> owned;  // device buffers packed by the builder (freed with the model)
  UNetCfg cfg;
  std::string pfx;
  std::vector<std::vector<UOp>> inp, out;
  std::vector<UOp> mid;
  Lin32 te0, te2, emb_cat;
  const float *out_g = nullptr, *out_b = nullptr;
  LinW out_taps;  // out.2 (3x3, 320 -> 4) as a [36, 320] tap GEMM: row = tap * 4 + co (elem.cuh: eps_gather)
  const float* out_bias = nullptr;
};

struct DdimSchedule {
  int n = 0;
  std::vector<long long> timesteps;
  std::vector<float> a_t, a_prev, sigma, sqrt_one_minus_a;
};

UNet* build_unet(Ctx& c, const std::string& pfx, const UNetCfg& cfg);
struct UNetAux {                        // optional step-invariant inputs of a forward pass
  int uniform_t = 0;                    // all N samples share t[0] (DDIM loop)
  // classifier-free guidance batch: x9[n] == x9[n + N/2] and t[n] == t[n + N/2] (ddim.py:338-344), only the context
  // differs.  Everything before the first context-dependent operation (conv_in, the first ResBlock, attn1 of the first
  // SpatialTransformer) is then computed once for both halves.
  int cfg_dup = 0;
  std::vector<const float*> crossvec;   // per SpatialTransformer (execution order): to_out(to_v(ctx)) [N, C]
  // uniform_t only: this step's row of every ResBlock's emb_layers output (UNet::emb_cat), precomputed for all the
  // steps of a sampling run (unet_time_embeddings); nullptr: computed inside the forward pass
  const float* emb_all = nullptr;
  // != nullptr: leave the output convolution's tap partials [N*L*L, 36] fp32 here instead of writing eps (the sampler's
  // update kernel finishes the convolution itself)
  float* taps_out = nullptr;
};
std::vector<const float*> unet_cross_vectors(Ctx& c, UNet& u, const float* ctx, int N, int T);
// emb_layers(SiLU(time_embed(timestep_embedding(t)))) of all ResBlocks for R timesteps: [R, u.emb_cat.out] (arena)
float* unet_time_embeddings(Ctx& c, UNet& u, const long long* t, int R);
void unet_forward(Ctx& c, UNet& u, const float* x9, const long long* t, const float* ctx, int N, int L, int T,
                  float* eps, const UNetAux* aux = nullptr);
Tens cross_attention_general(Ctx& c, const STW& s, const Tens& x, const float* ctx, int T, int N);
void ddim_sample(Ctx& c, UNet& u, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                 const float* uncond, int B, int L, int T, const DdimSchedule& s, float scale, const float* noise,
                 float* x0_out, float* inter_x, float* inter_p0, int log_every_t);
void plms_sample(Ctx& c, UNet& u, const float* x_T, const float* z_inpaint, const float* mask, const float* cond,
                 const float* uncond, int B, int L, int T, const DdimSchedule& s, float scale, float* x0_out,
                 float* inter_x, float* inter_p0, int log_every_t);

// ---- AutoencoderKL
struct VResW {
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  ConvW c1, c2, nin;
  bool skip = false;
};
struct VAttnW {
  const float *g = nullptr, *b = nullptr;
  LinW qkv, proj;
  float* qkv_bias = nullptr;
};
struct VAE {
  std::vector<void*> owned;  // device buffers packed by the builder (freed with the model)
  std::string pfx;
  int ch = 128, z = 4;
  std::vector<int> mult = {1, 2, 4, 4};
  // encoder
  ConvW e_in, e_out;
  std::vector<std::vector<VResW>> e_down;
  std::vector<ConvW> e_ds;
  VResW e_mid1, e_mid2;
  VAttnW e_attn;
  const float *e_ng = nullptr, *e_nb = nullptr;
  const float *quant_w = nullptr, *quant_b = nullptr, *pquant_w = nullptr, *pquant_b = nullptr;
  // decoder
  ConvW d_in, d_out;
  VResW d_mid1, d_mid2;
  VAttnW d_attn;
  std::vector<std::vector<VResW>> d_up;  // indexed by level
  std::vector<ConvW> d_us;
  const float *d_ng = nullptr, *d_nb = nullptr;
};
VAE* build_vae(Ctx& c, const std::string& pfx);
// scale: LatentDiffusion.scale_factor (z = scale * sample; 1.0 gives the bare posterior sample); inv_scale: fp32(1/scale)
void vae_encode(Ctx& c, VAE& v, const float* img, const float* noise, int B, int H, int W, float scale, float* z,
                float* mean, float* logvar);
void vae_decode(Ctx& c, VAE& v, const float* z, int B, int h, int w, float inv_scale, float* img);

// ---- conditioning encoders
struct ClipLayerW {
  const float *ln1g, *ln1b, *ln2g, *ln2b;
  LinW qkv, o, fc1, fc2;
  float* qkv_bias;
};
struct MapperLayerW {
  const float *ln1g, *ln1b, *ln2g, *ln2b;
  Lin32 qkv, proj, fc, fc2;
};
struct ClipVision {
  std::vector<void*> owned;  // device buffers packed by the builder (freed with the model)
  std::string pfx;
  int width = 1024, heads = 16, patch = 14, image = 224, ntok = 257, layers = 24, proj = 768;
  LinW patch_w;
  const float *cls = nullptr, *pos = nullptr, *pre_g = nullptr, *pre_b = nullptr, *post_g = nullptr, *post_b = nullptr;
  std::vector<ClipLayerW> L;
  Lin32 vproj;
  std::vector<MapperLayerW> M;
  const float *fln_g = nullptr, *fln_b = nullptr;
};
ClipVision* build_clip(Ctx& c, const std::string& pfx);
void clip_embed(Ctx& c, ClipVision& m, const float* img, int B, float* out768);

struct ArcUnitW {
  int cin, depth, stride;
  bool sc_conv;
  ConvW sc, c1, c2;               // BN folded
  float *bn0_s, *bn0_t;           // res_layer.0 (BN before conv1): applied as per-channel affine on the input
  float *c1_bias, *c2_bias, *sc_bias;
  const float* prelu;
  Lin32 se1, se2;
};
struct ArcFace {
  std::vector<void*> owned;  // device buffers packed by the builder (freed with the model)
  std::string pfx;
  ConvW stem;
  float* stem_bias;
  const float* stem_prelu;
  std::vector<ArcUnitW> units;
  float *out_s, *out_t;  // output_layer.0 BN as affine
  Lin32 fc;              // Linear 25088->512 with BN1d folded
  float* fc_w = nullptr; float* fc_b = nullptr;
};
ArcFace* build_arcface(Ctx& c, const std::string& pfx);
// eval-mode BatchNorm `p` as a per-channel affine: s = gamma / sqrt(var + eps), t = beta - mean * s (device arrays)
void bn_affine(Ctx& c, const std::string& p, int C, float** s, float** t);
void channel_mean(Ctx& c, const Tens& x, float* out);  // mean over H*W per (n, c): fp32 [N, C]

// ---- face parsing (BiSeNet, pretrained/face_parsing/model.py)
struct ParseCBR {  // bias-free conv with the following BatchNorm folded in (+ ReLU in the epilogue)
  ConvW w;
  float* bias = nullptr;
};
struct ParseBlock {  // resnet.py BasicBlock
  ParseCBR c1, c2, ds;
  bool down = false;
  int stride = 1;
};
struct FaceParser {
  std::vector<void*> owned;  // device buffers packed by the builder (freed with the model)
  std::string pfx;
  ParseCBR stem, arm32, arm16, head32, head16, ffm_blk, out_cbr;
  std::vector<ParseBlock> blocks;
  Lin32 conv_avg, arm32_att, arm16_att, ffm1, ffm2;  // 1x1 convs on per-sample channel vectors (fp32 GEMV path)
  ConvW out_conv;
  int n_classes = 19;
};
FaceParser* build_face_parser(Ctx& c, const std::string& pfx);
void face_parse(Ctx& c, FaceParser& m, const float* img01, int B, int H, int W, float* logits8, uint8_t* seg19,
                uint8_t* seg12);
// paste-back (scripts/inference_swap_video.py:702-721), Pillow-exact: see paste.cu
void paste_back(Ctx& c, const float* x01, const uint8_t* orig, const double* coeffs, int B, int h, int w, int up, int H, int W,
                uint8_t* out);
void inpaint_from_parsing(Ctx& c, const float* img, const uint8_t* seg12, unsigned remove_bits, int B, int H, int W,
                          float* mask, float* inpaint);
void arcface_embed(Ctx& c, ArcFace& m, const float* img224, int B, float* out512);

}  // namespace rfb
