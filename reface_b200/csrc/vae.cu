// AutoencoderKL encode / decode on the engine's kernels.
//   Encoder.forward / Decoder.forward / ResnetBlock / AttnBlock / Up-/Downsample
//                                   ldm/modules/diffusionmodules/model.py:33-202, 368-568
//   AutoencoderKL.encode/decode     ldm/models/autoencoder.py:324-333
//   DiagonalGaussianDistribution    ldm/modules/distributions/distributions.py:24-37
//   scale_factor handling           ldm/models/diffusion/ddpm.py:850-857, 1277-1337
#include "models.h"

namespace rfb {

// raw encoder moments h [M,8] fp32 -> quant_conv (1x1, 8->8) -> mean/logvar(clamped) -> z = s*(mean + std*noise)
__global__ void vae_quant_sample_kernel(const float* __restrict__ h, const float* __restrict__ Wq,
                                        const float* __restrict__ bq, const float* __restrict__ noise,
                                        float* __restrict__ z, float* __restrict__ mean, float* __restrict__ logvar, int B,
                                        int HW, float scale) {
  const long long total = (long long)B * HW;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
    float in[8], o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) in[j] = h[m * 8 + j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = bq[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) a += Wq[i * 8 + j] * in[j];
      o[i] = a;
    }
    const int b = (int)(m / HW), px = (int)(m % HW);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const long long idx = ((long long)b * 4 + ch) * HW + px;
      const float mu = o[ch];
      const float lv = fminf(fmaxf(o[4 + ch], -30.0f), 20.0f);
      if (mean) mean[idx] = mu;
      if (logvar) logvar[idx] = lv;
      if (z) z[idx] = scale * (mu + expf(0.5f * lv) * (noise ? noise[idx] : 0.0f));
    }
  }
}

// z NCHW fp32 (first 4 channels) -> (1/scale) -> post_quant_conv (1x1, 4->4) -> NHWC fp16 [B,h,w,4]
__global__ void vae_post_quant_kernel(const float* __restrict__ z, const float* __restrict__ Wp,
                                      const float* __restrict__ bp, __half* __restrict__ out, int B, int HW, int zc,
                                      float inv_scale) {
  const long long total = (long long)B * HW;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(m / HW), px = (int)(m % HW);
    float in[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) in[j] = inv_scale * z[((long long)b * zc + j) * HW + px];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = bp[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) a += Wp[i * 4 + j] * in[j];
      out[m * 4 + i] = __float2half_rn(a);
    }
  }
}

static VResW build_vres(Ctx& c, const std::string& p) {
  VResW r;
  r.g1 = c.pf(p + "norm1.weight"), r.b1 = c.pf(p + "norm1.bias");
  r.g2 = c.pf(p + "norm2.weight"), r.b2 = c.pf(p + "norm2.bias");
  r.c1 = pack_conv(c, p + "conv1.weight", p + "conv1.bias");
  r.c2 = pack_conv(c, p + "conv2.weight", p + "conv2.bias");
  r.skip = c.has(p + "nin_shortcut.weight");
  if (r.skip) r.nin = pack_conv(c, p + "nin_shortcut.weight", p + "nin_shortcut.bias");
  return r;
}
static VAttnW build_vattn(Ctx& c, const std::string& p) {
  VAttnW a;
  a.g = c.pf(p + "norm.weight"), a.b = c.pf(p + "norm.bias");
  a.qkv = pack_linear_rows(c, {p + "q.weight", p + "k.weight", p + "v.weight"});
  const int C = a.qkv.in;
  a.qkv_bias = (float*)c.dmalloc((size_t)3 * C * sizeof(float));
  const char* nm[3] = {"q.bias", "k.bias", "v.bias"};
  for (int i = 0; i < 3; ++i)
    CUDA_OK(cudaMemcpyAsync(a.qkv_bias + (size_t)i * C, c.pf(p + nm[i]), C * sizeof(float), cudaMemcpyDeviceToDevice,
                            c.stream));
  a.qkv.b = a.qkv_bias;
  a.proj = pack_linear(c, p + "proj_out.weight", p + "proj_out.bias");
  return a;
}

VAE* build_vae(Ctx& c, const std::string& pfx) {
  VAE* v = new VAE();
  v->pfx = pfx;
  const std::string e = pfx + "encoder.", d = pfx + "decoder.";
  const int nlev = (int)v->mult.size();
  v->e_in = pack_conv(c, e + "conv_in.weight", e + "conv_in.bias");
  for (int l = 0; l < nlev; ++l) {
    std::vector<VResW> blocks;
    for (int j = 0; j < 2; ++j) blocks.push_back(build_vres(c, e + "down." + std::to_string(l) + ".block." + std::to_string(j) + "."));
    v->e_down.push_back(blocks);
    if (l != nlev - 1) {
      const std::string p = e + "down." + std::to_string(l) + ".downsample.conv.";
      v->e_ds.push_back(pack_conv(c, p + "weight", p + "bias"));
    }
  }
  v->e_mid1 = build_vres(c, e + "mid.block_1.");
  v->e_attn = build_vattn(c, e + "mid.attn_1.");
  v->e_mid2 = build_vres(c, e + "mid.block_2.");
  v->e_ng = c.pf(e + "norm_out.weight"), v->e_nb = c.pf(e + "norm_out.bias");
  v->e_out = pack_conv(c, e + "conv_out.weight", e + "conv_out.bias");
  v->quant_w = c.pf(pfx + "quant_conv.weight"), v->quant_b = c.pf(pfx + "quant_conv.bias");
  v->pquant_w = c.pf(pfx + "post_quant_conv.weight"), v->pquant_b = c.pf(pfx + "post_quant_conv.bias");
  v->d_in = pack_conv(c, d + "conv_in.weight", d + "conv_in.bias");
  v->d_mid1 = build_vres(c, d + "mid.block_1.");
  v->d_attn = build_vattn(c, d + "mid.attn_1.");
  v->d_mid2 = build_vres(c, d + "mid.block_2.");
  v->d_up.resize(nlev);
  v->d_us.resize(nlev);
  for (int l = 0; l < nlev; ++l) {
    for (int j = 0; j < 3; ++j) v->d_up[l].push_back(build_vres(c, d + "up." + std::to_string(l) + ".block." + std::to_string(j) + "."));
    if (l != 0) {
      const std::string p = d + "up." + std::to_string(l) + ".upsample.conv.";
      v->d_us[l] = pack_upconv(c, p + "weight", p + "bias");  // model.py:53-66: nearest-2x folded into the conv
    }
  }
  v->d_ng = c.pf(d + "norm_out.weight"), v->d_nb = c.pf(d + "norm_out.bias");
  v->d_out = pack_conv(c, d + "conv_out.weight", d + "conv_out.bias");
  CUDA_OK(cudaStreamSynchronize(c.stream));
  return v;
}

static Epi stats_epi() {  // plain epilogue that also leaves GroupNorm partial statistics with the output (engine.cu)
  Epi e;
  e.want_stats = true;
  return e;
}
static Tens run_vres(Ctx& c, const VResW& r, const Tens& x) {  // model.py:121-141
  Tens h = groupnorm(c, x, r.g1, r.b1, 1e-6f, true);
  Tens h1 = conv3x3_t(c, h, r.c1, stats_epi());
  Tens h2 = groupnorm(c, h1, r.g2, r.b2, 1e-6f, true);
  Tens skip = x;
  if (r.skip) skip = conv3x3_t(c, x, r.nin, Epi(), 1, 0, 0, 0, 0);
  Epi e;
  e.res = skip.p, e.ldr = skip.c;
  e.want_stats = true;  // every block output is normalised next (the following block, the attention block or norm_out)
  return conv3x3_t(c, h2, r.c2, e);
}
static Tens run_vattn(Ctx& c, const VAttnW& a, const Tens& x) {  // model.py:178-202
  const int C = x.c;
  Tens hn = groupnorm(c, x, a.g, a.b, 1e-6f, false);
  Tens qkv = linear_t(c, hn, a.qkv, Epi());
  Tens o = c.new_tens(x.n, x.h, x.w, C);
  attention(c, qkv.p, 3 * C, x.n, x.h * x.w, 1, C, o.p, C, 1.0f / sqrtf((float)C), 0, C, 2 * C);
  Epi e;
  e.res = x.p, e.ldr = C;
  e.want_stats = true;
  return linear_t(c, o, a.proj, e);
}

void vae_encode(Ctx& c, VAE& v, const float* img, const float* noise, int B, int H, int W, float scale, float* z,
                float* mean, float* logvar) {
  const size_t mk = c.mark();
  Tens h = from_nchw_f32(c, img, B, 3, H, W, 3);
  h = conv3x3_t(c, h, v.e_in, stats_epi());
  const int nlev = (int)v.mult.size();
  for (int l = 0; l < nlev; ++l) {
    for (auto& r : v.e_down[l]) h = run_vres(c, r, h);
    if (l != nlev - 1) h = conv3x3_t(c, h, v.e_ds[l], stats_epi(), 2, 0, 0, 1, 1);  // model.py:72-79: pad (0,1,0,1), stride 2
  }
  h = run_vres(c, v.e_mid1, h);
  h = run_vattn(c, v.e_attn, h);
  h = run_vres(c, v.e_mid2, h);
  h = groupnorm(c, h, v.e_ng, v.e_nb, 1e-6f, true);
  const long long M = h.rows();
  float* raw = c.alloc_t<float>((size_t)M * 8);
  Epi e;
  e.out32 = raw, e.o32_sn = 8, e.o32_sp = 0, e.o32_sc = 1, e.o32_rpn = 1;
  conv3x3_t(c, h, v.e_out, e);
  vae_quant_sample_kernel<<<grid_for(M), 256, 0, c.stream>>>(raw, v.quant_w, v.quant_b, noise, z, mean, logvar, B,
                                                           h.h * h.w, scale);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  c.release(mk);
}

void vae_decode(Ctx& c, VAE& v, const float* z, int B, int hh, int ww, float inv_scale, float* img) {
  const size_t mk = c.mark();
  Tens h = c.new_tens(B, hh, ww, 4);
  vae_post_quant_kernel<<<grid_for((long long)B * hh * ww), 256, 0, c.stream>>>(z, v.pquant_w, v.pquant_b, h.p, B,
                                                                               hh * ww, 4, inv_scale);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  h = conv3x3_t(c, h, v.d_in, stats_epi());
  h = run_vres(c, v.d_mid1, h);
  h = run_vattn(c, v.d_attn, h);
  h = run_vres(c, v.d_mid2, h);
  const int nlev = (int)v.mult.size();
  for (int l = nlev - 1; l >= 0; --l) {
    for (auto& r : v.d_up[l]) h = run_vres(c, r, h);
    if (l != 0) h = upconv3x3_t(c, h, v.d_us[l], stats_epi());
  }
  h = groupnorm(c, h, v.d_ng, v.d_nb, 1e-6f, true);
  const long long HW = (long long)h.h * h.w;
  Epi e;
  e.out32 = img, e.o32_sn = 3 * HW, e.o32_sp = 1, e.o32_sc = HW, e.o32_rpn = (int)HW;
  conv3x3_t(c, h, v.d_out, e);
  c.release(mk);
}

}  // namespace rfb
