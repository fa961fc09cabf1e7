// CLIP ViT-L/14 vision tower + REFace's mapper2/final_ln2 head on the engine's kernels.
//   FrozenCLIPEmbedder.forward      ldm/modules/encoders/modules.py:253-261
//   vision tower                    third-party `transformers` CLIPVisionTransformer (pre-LN ViT, quick-GELU,
//                                   CLS pooling + post-LN, bias-free visual_projection), called at modules.py:254-256
//   mapper2 (1-token transformer)   ldm/modules/encoders/xf.py:22-130
#include "models.h"

namespace rfb {

// x[b, 0] = cls + pos[0];  x[b, 1+p] = patch[b, p] + pos[1+p]      (fp16 out)
__global__ void clip_assemble_kernel(const __half* __restrict__ patches, const float* __restrict__ cls,
                                     const float* __restrict__ pos, __half* __restrict__ x, int B, int ntok, int W) {
  const long long total = (long long)B * ntok * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W);
    const int tok = (int)((i / W) % ntok);
    const long long b = i / ((long long)W * ntok);
    const float v = (tok == 0) ? cls[c] : __half2float(patches[(b * (ntok - 1) + (tok - 1)) * W + c]);
    x[i] = __float2half_rn(v + pos[(long long)tok * W + c]);
  }
}
// LayerNorm of selected fp16 rows (row r at x + r*ldx) -> fp32 [rows, C]; one block per row
__global__ void ln_rows_f16_to_f32_kernel(const __half* __restrict__ x, long long ldx, const float* __restrict__ g,
                                          const float* __restrict__ b, float* __restrict__ out, int C, float eps) {
  __shared__ float red[32];
  const __half* p = x + (long long)blockIdx.x * ldx;
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s += __half2float(p[i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const float mean = tot / (float)C;
  __syncthreads();
  float ss = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float d = __half2float(p[i]) - mean;
    ss += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float vt = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) vt += red[w];
  const float rstd = rsqrtf(vt / (float)C + eps);
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    out[(long long)blockIdx.x * C + i] = (__half2float(p[i]) - mean) * rstd * g[i] + b[i];
}
// fp32 LayerNorm, one block per row
__global__ void ln_f32_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                              float* __restrict__ out, int C, float eps) {
  __shared__ float red[32];
  const float* p = x + (long long)blockIdx.x * C;
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s += p[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const float mean = tot / (float)C;
  __syncthreads();
  float ss = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float d = p[i] - mean;
    ss += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float vt = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) vt += red[w];
  const float rstd = rsqrtf(vt / (float)C + eps);
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    out[(long long)blockIdx.x * C + i] = (p[i] - mean) * rstd * g[i] + b[i];
}

static float* cat_bias(Ctx& c, const std::vector<std::string>& names, int each) {
  float* b = (float*)c.dmalloc(names.size() * each * sizeof(float));
  for (size_t i = 0; i < names.size(); ++i)
    CUDA_OK(cudaMemcpyAsync(b + i * each, c.pf(names[i]), each * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  return b;
}

ClipVision* build_clip(Ctx& c, const std::string& pfx) {
  ClipVision* m = new ClipVision();
  m->pfx = pfx;
  const std::string v = pfx + "model.vision_model.";
  {
    ConvW pw = pack_conv(c, v + "embeddings.patch_embedding.weight", "");
    m->patch_w.w = pw.w, m->patch_w.in = pw.taps * pw.cin, m->patch_w.out = pw.cout, m->patch_w.kp = pw.kp;
    m->width = pw.cout, m->patch = pw.ksz;
  }
  m->cls = c.pf(v + "embeddings.class_embedding");
  m->pos = c.pf(v + "embeddings.position_embedding.weight");
  m->ntok = (int)c.param(v + "embeddings.position_embedding.weight").shape[0];
  m->pre_g = c.pf(v + "pre_layrnorm.weight"), m->pre_b = c.pf(v + "pre_layrnorm.bias");
  m->post_g = c.pf(v + "post_layernorm.weight"), m->post_b = c.pf(v + "post_layernorm.bias");
  const int W = m->width;
  m->layers = 0;
  while (c.has(v + "encoder.layers." + std::to_string(m->layers) + ".layer_norm1.weight")) ++m->layers;
  for (int i = 0; i < m->layers; ++i) {
    const std::string l = v + "encoder.layers." + std::to_string(i) + ".";
    ClipLayerW w;
    w.ln1g = c.pf(l + "layer_norm1.weight"), w.ln1b = c.pf(l + "layer_norm1.bias");
    w.ln2g = c.pf(l + "layer_norm2.weight"), w.ln2b = c.pf(l + "layer_norm2.bias");
    w.qkv = pack_linear_rows(c, {l + "self_attn.q_proj.weight", l + "self_attn.k_proj.weight", l + "self_attn.v_proj.weight"});
    w.qkv_bias = cat_bias(c, {l + "self_attn.q_proj.bias", l + "self_attn.k_proj.bias", l + "self_attn.v_proj.bias"}, W);
    w.qkv.b = w.qkv_bias;
    w.o = pack_linear(c, l + "self_attn.out_proj.weight", l + "self_attn.out_proj.bias");
    w.fc1 = pack_linear(c, l + "mlp.fc1.weight", l + "mlp.fc1.bias");
    w.fc2 = pack_linear(c, l + "mlp.fc2.weight", l + "mlp.fc2.bias");
    m->L.push_back(w);
  }
  m->vproj = lin32(c, pfx + "model.visual_projection.weight", "");
  m->proj = m->vproj.out;
  int nm = 0;
  while (c.has(pfx + "mapper2.resblocks." + std::to_string(nm) + ".ln_1.weight")) ++nm;
  for (int i = 0; i < nm; ++i) {
    const std::string l = pfx + "mapper2.resblocks." + std::to_string(i) + ".";
    MapperLayerW w;
    w.ln1g = c.pf(l + "ln_1.weight"), w.ln1b = c.pf(l + "ln_1.bias");
    w.ln2g = c.pf(l + "ln_2.weight"), w.ln2b = c.pf(l + "ln_2.bias");
    w.qkv = lin32(c, l + "attn.c_qkv.weight", l + "attn.c_qkv.bias");
    w.proj = lin32(c, l + "attn.c_proj.weight", l + "attn.c_proj.bias");
    w.fc = lin32(c, l + "mlp.c_fc.weight", l + "mlp.c_fc.bias");
    w.fc2 = lin32(c, l + "mlp.c_proj.weight", l + "mlp.c_proj.bias");
    m->M.push_back(w);
  }
  m->fln_g = c.pf(pfx + "final_ln2.weight"), m->fln_b = c.pf(pfx + "final_ln2.bias");
  m->heads = 16;
  CUDA_OK(cudaStreamSynchronize(c.stream));
  return m;
}

void clip_embed(Ctx& c, ClipVision& m, const float* img, int B, float* out) {
  const size_t mk = c.mark();
  const int W = m.width, nt = m.ntok, np = nt - 1, P = m.proj;
  const int side = m.image / m.patch;
  RFB_CHECK(side * side == np, "CLIP: position embedding does not match image/patch size");
  // patch embedding: 14x14 stride-14 conv == GEMM over non-overlapping patches (im2col path)
  Tens x0 = from_nchw_f32(c, img, B, 3, m.image, m.image, 3);
  ConvW pw;
  pw.w = m.patch_w.w, pw.cin = 3, pw.cout = W, pw.cin_p = 3, pw.ksz = m.patch, pw.taps = m.patch * m.patch, pw.kp = m.patch_w.kp;
  Tens patches = conv3x3_t(c, x0, pw, Epi(), m.patch, 0, 0, 0, 0);  // [B, side, side, W]
  Tens x = c.new_tens(B, 1, nt, W);
  clip_assemble_kernel<<<grid_for(x.rows() * W), 256, 0, c.stream>>>(patches.p, m.cls, m.pos, x.p, B, nt, W);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  x = layernorm(c, x, m.pre_g, m.pre_b, 1e-5f);
  const int d = W / m.heads;
  for (auto& l : m.L) {
    const size_t lm = c.mark();
    Tens h = layernorm(c, x, l.ln1g, l.ln1b, 1e-5f);
    Tens qkv = linear_t(c, h, l.qkv, Epi());
    Tens a = c.new_tens(B, 1, nt, W);
    attention(c, qkv.p, 3 * W, B, nt, m.heads, d, a.p, W, 1.0f / sqrtf((float)d), 0, W, 2 * W);
    Epi e1;
    e1.res = x.p, e1.ldr = W;
    Tens x1 = linear_t(c, a, l.o, e1);
    Tens h2 = layernorm(c, x1, l.ln2g, l.ln2b, 1e-5f);
    Epi ef;
    ef.act = 3;  // quick-GELU
    Tens f = linear_t(c, h2, l.fc1, ef);
    // write the block output over x (same shape) so that the per-layer temporaries can be released
    Epi e2;
    e2.res = x1.p, e2.ldr = W;
    gemm(c, f.p, f.c, f.rows(), f.c, l.fc2.w, l.fc2.kp, l.fc2.out, x.p, W, [&] { Epi t = e2; t.bias = l.fc2.b; return t; }());
    c.release(lm);
  }
  // pooled = post_layernorm(CLS); visual_projection; mapper2 (1 token: attention == identity on v); final_ln2
  float* pooled = c.alloc_t<float>((size_t)B * W);
  ln_rows_f16_to_f32_kernel<<<B, 256, 0, c.stream>>>(x.p, (long long)nt * W, m.post_g, m.post_b, pooled, W, 1e-5f);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  float* z = c.alloc_t<float>((size_t)B * P);
  float* t0 = c.alloc_t<float>((size_t)B * P);
  float* qkv = c.alloc_t<float>((size_t)B * 3 * P);
  float* ff = c.alloc_t<float>((size_t)B * 4 * P);
  linear_small(c, pooled, W, B, m.vproj, z, P, 0, 0);
  for (auto& l : m.M) {
    ln_f32_kernel<<<B, 256, 0, c.stream>>>(z, l.ln1g, l.ln1b, t0, P, 1e-5f);
    c.launches++;
    linear_small(c, t0, P, B, l.qkv, qkv, 3 * P, 0, 0);
    // n_ctx = 1, heads = 1: softmax over a single key is 1 -> attention output = v (xf.py:66-77)
    linear_small(c, qkv + 2 * P, 3 * P, B, l.proj, z, P, 0, 0, z);
    ln_f32_kernel<<<B, 256, 0, c.stream>>>(z, l.ln2g, l.ln2b, t0, P, 1e-5f);
    c.launches++;
    linear_small(c, t0, P, B, l.fc, ff, 4 * P, 0, /*gelu*/ 2);
    linear_small(c, ff, 4 * P, B, l.fc2, z, P, 0, 0, z);
  }
  ln_f32_kernel<<<B, 256, 0, c.stream>>>(z, m.fln_g, m.fln_b, out, P, 1e-5f);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  c.release(mk);
}

}  // namespace rfb
