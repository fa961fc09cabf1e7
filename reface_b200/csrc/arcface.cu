// ArcFace IR-SE50 identity encoder + its pre-processing on the engine's kernels.
//   IDLoss.extract_feats            ldm/models/diffusion/ddpm.py:112-124
//   Backbone.forward                src/Face_models/encoders/model_irse.py:44-69
//   bottleneck_IR_SE / SEModule     src/Face_models/encoders/helpers.py:56-72, 97-119
// Eval-mode BatchNorm that FOLLOWS a conv is folded into the conv weights (exact); BatchNorm that PRECEDES a
// zero-padded conv is applied as an explicit per-channel affine (folding it would change the border pixels).
#include "models.h"

namespace rfb {

// s = gamma / sqrt(var + eps), t = beta - mean * s
__global__ void bn_affine_kernel(const float* g, const float* b, const float* mean, const float* var, float* s, float* t,
                                 int C, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float sc = g[i] * rsqrtf(var[i] + eps);
    s[i] = sc;
    t[i] = b[i] - mean[i] * sc;
  }
}
// y = x * s[c] + t[c]   (NHWC fp16)
__global__ void affine_c_kernel(const __half* __restrict__ x, const float* __restrict__ s, const float* __restrict__ t,
                                __half* __restrict__ y, long long total, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    y[i] = __float2half_rn(__half2float(x[i]) * s[c] + t[c]);
  }
}
__global__ void affine_f32_kernel(const float* x, const float* s, const float* t, float* y, int total, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) y[i] = x[i] * s[i % C] + t[i % C];
}
// mean over HW per (n, c): NHWC fp16 -> fp32 [N, C]
__global__ void channel_mean_kernel(const __half* __restrict__ x, float* __restrict__ out, int HW, int C) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += __half2float(x[((long long)n * HW + p) * C + c]);
  out[(long long)n * C + c] = s / (float)HW;
}
// out = r * se[n,c] + shortcut(n, oy*stride, ox*stride, c)
__global__ void se_scale_add_kernel(const __half* __restrict__ r, const float* __restrict__ se,
                                    const __half* __restrict__ sc, __half* __restrict__ out, int N, int Ho, int Wo, int C,
                                    int sH, int sW, int stride) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    const float s = __half2float(sc[(((long long)n * sH + oy * stride) * sW + ox * stride) * C + c]);
    out[i] = __float2half_rn(__half2float(r[i]) * se[(long long)n * C + c] + s);
  }
}
__global__ void l2norm_kernel(const float* x, float* out, int C) {
  __shared__ float red[32];
  const float* p = x + (long long)blockIdx.x * C;
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s += p[i] * p[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const float inv = 1.0f / sqrtf(tot);
  for (int i = threadIdx.x; i < C; i += blockDim.x) out[(long long)blockIdx.x * C + i] = p[i] * inv;
}
// un-CLIP-normalise -> [-1,1] -> AdaptiveAvgPool(256) -> crop [35:223, 32:220] -> AdaptiveAvgPool(112); NHWC fp16 out
__global__ void arcface_preproc_kernel(const float* __restrict__ img, __half* __restrict__ out, int B, int S) {
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  const long long total = (long long)B * 112 * 112 * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 3);
    const int ox = (int)((i / 3) % 112), oy = (int)((i / (3 * 112)) % 112);
    const long long b = i / (3 * 112 * 112);
    const float* p = img + (b * 3 + ch) * (long long)S * S;
    // second pool: 188 -> 112
    const int ys = (oy * 188) / 112, ye = ((oy + 1) * 188 + 111) / 112;
    const int xs = (ox * 188) / 112, xe = ((ox + 1) * 188 + 111) / 112;
    float acc = 0.f;
    for (int y = ys; y < ye; ++y)
      for (int x = xs; x < xe; ++x) {
        float v;
        if (S == 256) {
          v = (p[(long long)(y + 35) * S + (x + 32)] * stdv[ch] + mean[ch] - 0.5f) / 0.5f;
        } else {
          // first pool: S -> 256 at (y+35, x+32)
          const int Y = y + 35, X = x + 32;
          const int y0 = (Y * S) / 256, y1 = ((Y + 1) * S + 255) / 256;
          const int x0 = (X * S) / 256, x1 = ((X + 1) * S + 255) / 256;
          float a = 0.f;
          for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) a += (p[(long long)yy * S + xx] * stdv[ch] + mean[ch] - 0.5f) / 0.5f;
          v = a / (float)((y1 - y0) * (x1 - x0));
        }
        acc += v;
      }
    out[i] = __float2half_rn(acc / (float)((ye - ys) * (xe - xs)));
  }
}

void bn_affine(Ctx& c, const std::string& p, int C, float** s, float** t) {
  *s = (float*)c.dmalloc(C * sizeof(float));
  *t = (float*)c.dmalloc(C * sizeof(float));
  bn_affine_kernel<<<(C + 127) / 128, 128, 0, c.stream>>>(c.pf(p + ".weight"), c.pf(p + ".bias"), c.pf(p + ".running_mean"),
                                                         c.pf(p + ".running_var"), *s, *t, C, 1e-5f);
  CUDA_OK(cudaGetLastError());
  c.launches++;
}

void channel_mean(Ctx& c, const Tens& x, float* out) {
  dim3 g((unsigned)((x.c + 127) / 128), (unsigned)x.n);
  channel_mean_kernel<<<g, 128, 0, c.stream>>>(x.p, out, x.h * x.w, x.c);
  CUDA_OK(cudaGetLastError());
  c.launches++;
}

ArcFace* build_arcface(Ctx& c, const std::string& pfx) {
  ArcFace* m = new ArcFace();
  m->pfx = pfx;
  float *s, *t;
  bn_affine(c, pfx + "input_layer.1", 64, &s, &t);
  m->stem = pack_conv(c, pfx + "input_layer.0.weight", "", s);
  m->stem_bias = t;
  m->stem_prelu = c.pf(pfx + "input_layer.2.weight");
  static const int blocks[4][3] = {{64, 64, 3}, {64, 128, 4}, {128, 256, 14}, {256, 512, 3}};  // helpers.py:29-36
  int idx = 0;
  for (auto& b : blocks) {
    for (int u = 0; u < b[2]; ++u) {
      ArcUnitW w;
      w.cin = (u == 0) ? b[0] : b[1], w.depth = b[1], w.stride = (u == 0) ? 2 : 1;
      const std::string p = pfx + "body." + std::to_string(idx++) + ".";
      w.sc_conv = w.cin != w.depth;
      w.sc_bias = nullptr;
      if (w.sc_conv) {
        bn_affine(c, p + "shortcut_layer.1", w.depth, &s, &t);
        w.sc = pack_conv(c, p + "shortcut_layer.0.weight", "", s);
        w.sc_bias = t;
      }
      bn_affine(c, p + "res_layer.0", w.cin, &w.bn0_s, &w.bn0_t);
      w.c1 = pack_conv(c, p + "res_layer.1.weight", "");
      w.prelu = c.pf(p + "res_layer.2.weight");
      bn_affine(c, p + "res_layer.4", w.depth, &s, &t);
      w.c2 = pack_conv(c, p + "res_layer.3.weight", "", s);
      w.c2_bias = t;
      w.c1_bias = nullptr;
      w.se1 = lin32(c, p + "res_layer.5.fc1.weight", "");
      w.se2 = lin32(c, p + "res_layer.5.fc2.weight", "");
      m->units.push_back(w);
    }
  }
  bn_affine(c, pfx + "output_layer.0", 512, &m->out_s, &m->out_t);
  m->fc = lin32(c, pfx + "output_layer.3.weight", pfx + "output_layer.3.bias");
  bn_affine(c, pfx + "output_layer.4", 512, &m->fc_w, &m->fc_b);  // BN1d as affine (s, t)
  CUDA_OK(cudaStreamSynchronize(c.stream));
  return m;
}

static Tens affine_c(Ctx& c, const Tens& x, const float* s, const float* t) {
  Tens y = c.new_tens(x.n, x.h, x.w, x.c);
  affine_c_kernel<<<grid_for(x.rows() * x.c), 256, 0, c.stream>>>(x.p, s, t, y.p, x.rows() * x.c, x.c);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  return y;
}

void arcface_embed(Ctx& c, ArcFace& m, const float* img, int B, float* out512) {
  const size_t mk = c.mark();
  Tens x = c.new_tens(B, 112, 112, 3);
  arcface_preproc_kernel<<<grid_for(x.rows() * 3), 256, 0, c.stream>>>(img, x.p, B, 224);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  {
    Epi e;
    e.bias = m.stem_bias, e.act = 4 /*PReLU*/, e.act_param = m.stem_prelu;
    x = conv3x3_t(c, x, m.stem, e);
  }
  for (auto& u : m.units) {
    Tens sc = x;
    if (u.sc_conv) {
      Epi e;
      e.bias = u.sc_bias;
      sc = conv3x3_t(c, x, u.sc, e, u.stride, 0, 0, 0, 0);
    }
    Tens r = affine_c(c, x, u.bn0_s, u.bn0_t);
    Epi e1;
    e1.act = 4, e1.act_param = u.prelu;
    r = conv3x3_t(c, r, u.c1, e1);
    Epi e2;
    e2.bias = u.c2_bias;
    r = conv3x3_t(c, r, u.c2, e2, u.stride);
    // SE: gap -> fc1 -> ReLU -> fc2 -> sigmoid -> scale; then + shortcut (MaxPool2d(1,stride) = subsample)
    float* gap = c.alloc_t<float>((size_t)B * u.depth);
    float* h1 = c.alloc_t<float>((size_t)B * (u.depth / 16));
    float* se = c.alloc_t<float>((size_t)B * u.depth);
    dim3 g((unsigned)((u.depth + 127) / 128), (unsigned)B);
    channel_mean_kernel<<<g, 128, 0, c.stream>>>(r.p, gap, r.h * r.w, u.depth);
    c.launches++;
    linear_small(c, gap, u.depth, B, u.se1, h1, u.depth / 16, 0, /*relu*/ 3);
    linear_small(c, h1, u.depth / 16, B, u.se2, se, u.depth, 0, /*sigmoid*/ 4);
    Tens y = c.new_tens(B, r.h, r.w, u.depth);
    const int sstride = u.sc_conv ? 1 : u.stride;
    se_scale_add_kernel<<<grid_for(y.rows() * u.depth), 256, 0, c.stream>>>(r.p, se, sc.p, y.p, B, r.h, r.w, u.depth, sc.h,
                                                                           sc.w, sstride);
    CUDA_OK(cudaGetLastError());
    c.launches++;
    x = y;
  }
  x = affine_c(c, x, m.out_s, m.out_t);
  float* flat = c.alloc_t<float>((size_t)B * 512 * 49);
  to_nchw_f32(c, x, flat);  // Flatten() runs over NCHW
  float* f = c.alloc_t<float>((size_t)B * 512);
  float* f2 = c.alloc_t<float>((size_t)B * 512);
  linear_small(c, flat, 512 * 49, B, m.fc, f, 512, 0, 0);
  affine_f32_kernel<<<(B * 512 + 255) / 256, 256, 0, c.stream>>>(f, m.fc_w, m.fc_b, f2, B * 512, 512);
  c.launches++;
  l2norm_kernel<<<B, 256, 0, c.stream>>>(f2, out512, 512);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  c.release(mk);
}

}  // namespace rfb
