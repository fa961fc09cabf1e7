// Face parsing: BiSeNet on a ResNet-18 context path, the 19 -> 12 label conversion and the mask /
// inpaint-image preparation -- the step BEFORE the swap path for arbitrary images and video frames.
//   BiSeNet / ContextPath / AttentionRefinementModule / FeatureFusionModule / BiSeNetOutput
//                                     pretrained/face_parsing/model.py:19-262
//   Resnet18 / BasicBlock             pretrained/face_parsing/resnet.py:19-85
//   FaceParser.preprocess_img/forward pretrained/face_parsing/face_parsing_demo.py:260-281
//   __ffhq_masks_to_faceParser_mask_detailed   face_parsing_demo.py:74-122
//   mask = 1 - isin(label, remove_tar), inpaint = image * mask   ldm/data/video_swap_dataset.py:150-222
// Every eval-mode BatchNorm FOLLOWS a bias-free conv and is folded into it (exact); all convolutions run on the
// tcgen05 GEMM / implicit-GEMM kernels of the engine, ReLU (and the BasicBlock's post-residual ReLU) in their epilogue.
#include "models.h"
#include "ptx.cuh"

namespace rfb {

// [B,3,H,W] fp32 in [0,1] -> clamp, (x - mean) / std -> NHWC fp16 [B,H,W,3]      (face_parsing_demo.py:266-268)
__global__ void parse_preproc_kernel(const float* __restrict__ img, __half* __restrict__ out, int B, int HW) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const long long total = (long long)B * HW * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 3);
    const long long p = i / 3;
    const long long b = p / HW, px = p % HW;
    const float v = fminf(fmaxf(img[(b * 3 + ch) * HW + px], 0.f), 1.f);
    out[i] = __float2half_rn((v - mean[ch]) / stdv[ch]);
  }
}
// MaxPool2d(3, stride 2, padding 1), NHWC fp16, C % 8 == 0
__global__ void maxpool3s2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int N, int H, int W, int C, int Ho,
                                  int Wo) {
  const int cv = C >> 3;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    long long p = i / cv;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + iy) * W + ix) * C + c8 * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_h2(w[t]);
          m[2 * t] = fmaxf(m[2 * t], f.x);
          m[2 * t + 1] = fmaxf(m[2 * t + 1], f.y);
        }
      }
    }
    uint4 o;
    o.x = pack_h2(m[0], m[1]), o.y = pack_h2(m[2], m[3]), o.z = pack_h2(m[4], m[5]), o.w = pack_h2(m[6], m[7]);
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}
// out = x * (att[n,c] + att_bias) + (vec ? vec[n,c] : 0) + (add ? add[i] : 0)     (ARM / FFM channel attention)
__global__ void parse_scale_add_kernel(const __half* __restrict__ x, const float* __restrict__ att, float att_bias,
                                       const float* __restrict__ vec, const __half* __restrict__ add,
                                       __half* __restrict__ out, long long total, int HW, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long n = i / ((long long)HW * C);
    float v = __half2float(x[i]) * (att[n * C + c] + att_bias);
    if (vec) v += vec[n * C + c];
    if (add) v += __half2float(add[i]);
    out[i] = __float2half_rn(v);
  }
}
// w'[o][i] = w[o][i] * s[o]  (BatchNorm folded into a 1x1 conv that runs on the fp32 GEMV path)
__global__ void scale_rows_kernel(const float* __restrict__ w, const float* __restrict__ s, float* __restrict__ out, int O,
                                  int I) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < O * I) out[i] = w[i] * s[i / I];
}
// F.interpolate(logits, (H, W), mode='bilinear', align_corners=True) + argmax over the classes + 19 -> 12 conversion.
// logits: fp32 [B, NC, h, w]; seg19 / seg12: uint8 [B, H, W] (either may be null).  Same interpolation arithmetic as
// torch's upsample_bilinear2d: src = dst * (in-1)/(out-1), lambda in fp32, w0y*(w0x*v00 + w1x*v01) + w1y*(...).
template <int NC>
__global__ void parse_upsample_argmax_kernel(const float* __restrict__ logits, uint8_t* __restrict__ seg19,
                                             uint8_t* __restrict__ seg12, int B, int h, int w, int H, int W) {
  const uint8_t map12[19] = {0, 6, 2, 2, 3, 3, 10, 7, 7, 11, 5, 9, 1, 1, 8, 0, 0, 4, 0};
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    const float fy = sy * (float)Y, fx = sx * (float)X;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const float* base = logits + b * NC * (long long)h * w;
    float best = -INFINITY;
    int arg = 0;
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
      const float* p = base + (long long)c * h * w;
      const float v = hy * (hx * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * (hx * p[y1 * w + x0] + lx * p[y1 * w + x1]);
      if (v > best) best = v, arg = c;  // first maximum wins, as torch.argmax
    }
    if (seg19) seg19[i] = (uint8_t)arg;
    if (seg12) seg12[i] = NC == 19 ? map12[arg] : (uint8_t)arg;
  }
}
// mask = 1 - isin(seg12, remove) ; inpaint = img * mask    (video_swap_dataset.py:150-222); remove_bits: bit k = label k
__global__ void parse_inpaint_kernel(const float* __restrict__ img, const uint8_t* __restrict__ seg12, unsigned remove_bits,
                                     float* __restrict__ mask, float* __restrict__ inpaint, int B, int HW) {
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, px = i % HW;
    const float m = ((remove_bits >> seg12[i]) & 1u) ? 0.f : 1.f;
    if (mask) mask[i] = m;
    if (inpaint)
      for (int ch = 0; ch < 3; ++ch) inpaint[(b * 3 + ch) * HW + px] = img[(b * 3 + ch) * HW + px] * m;
  }
}

// ---------------------------------------------------------------------------------------------- build
static ParseCBR build_cbr(Ctx& c, const std::string& conv, const std::string& bn) {
  ParseCBR r;
  float *s, *t;
  const int C = (int)c.param(conv + ".weight").shape[0];
  bn_affine(c, bn, C, &s, &t);
  r.w = pack_conv(c, conv + ".weight", "", s);
  r.bias = t;
  return r;
}
// 1x1 conv (+ optional folded BN) on the fp32 GEMV path: inputs are per-sample channel vectors
static Lin32 build_vec_conv(Ctx& c, const std::string& conv, const std::string& bn) {
  Lin32 l = lin32(c, conv + ".weight", "");
  if (!bn.empty()) {
    float *s, *t;
    bn_affine(c, bn, l.out, &s, &t);
    float* w2 = (float*)c.dmalloc((size_t)l.out * l.in * sizeof(float));
    scale_rows_kernel<<<(l.out * l.in + 255) / 256, 256, 0, c.stream>>>(l.w, s, w2, l.out, l.in);
    CUDA_OK(cudaGetLastError());
    c.launches++;
    l.w = w2, l.b = t;
  }
  return l;
}

FaceParser* build_face_parser(Ctx& c, const std::string& pfx) {
  FaceParser* m = new FaceParser();
  m->pfx = pfx;
  const std::string rn = pfx + "cp.resnet.";
  m->stem = build_cbr(c, rn + "conv1", rn + "bn1");
  static const int plan[4][2] = {{64, 1}, {128, 2}, {256, 2}, {512, 2}};  // resnet.py:65-68
  int cin = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < 2; ++bi) {
      ParseBlock b;
      const std::string p = rn + "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
      const int cout = plan[li][0];
      b.stride = bi == 0 ? plan[li][1] : 1;
      b.c1 = build_cbr(c, p + "conv1", p + "bn1");
      b.c2 = build_cbr(c, p + "conv2", p + "bn2");
      b.down = (cin != cout) || b.stride != 1;
      if (b.down) b.ds = build_cbr(c, p + "downsample.0", p + "downsample.1");
      m->blocks.push_back(b);
      cin = cout;
    }
  }
  const std::string cp = pfx + "cp.";
  m->conv_avg = build_vec_conv(c, cp + "conv_avg.conv", cp + "conv_avg.bn");
  m->arm32 = build_cbr(c, cp + "arm32.conv.conv", cp + "arm32.conv.bn");
  m->arm32_att = build_vec_conv(c, cp + "arm32.conv_atten", cp + "arm32.bn_atten");
  m->arm16 = build_cbr(c, cp + "arm16.conv.conv", cp + "arm16.conv.bn");
  m->arm16_att = build_vec_conv(c, cp + "arm16.conv_atten", cp + "arm16.bn_atten");
  m->head32 = build_cbr(c, cp + "conv_head32.conv", cp + "conv_head32.bn");
  m->head16 = build_cbr(c, cp + "conv_head16.conv", cp + "conv_head16.bn");
  m->ffm_blk = build_cbr(c, pfx + "ffm.convblk.conv", pfx + "ffm.convblk.bn");
  m->ffm1 = build_vec_conv(c, pfx + "ffm.conv1", "");
  m->ffm2 = build_vec_conv(c, pfx + "ffm.conv2", "");
  m->out_cbr = build_cbr(c, pfx + "conv_out.conv.conv", pfx + "conv_out.conv.bn");
  m->out_conv = pack_conv(c, pfx + "conv_out.conv_out.weight", "");
  m->n_classes = m->out_conv.cout;
  RFB_CHECK(m->n_classes == 19, "face parser: the label conversion expects the 19-class face-parsing.PyTorch head");
  CUDA_OK(cudaStreamSynchronize(c.stream));
  return m;
}

// ---------------------------------------------------------------------------------------------- forward
static Tens cbr(Ctx& c, const Tens& x, const ParseCBR& w, int stride = 1, int pad = -1) {
  Epi e;
  e.bias = w.bias, e.act = 5 /*ReLU*/;
  const int p = pad >= 0 ? pad : w.w.ksz / 2;
  return conv3x3_t(c, x, w.w, e, stride, p, p, p, p);
}
static float* channel_att(Ctx& c, const Tens& feat, const Lin32& l, int act_out) {
  float* gap = c.alloc_t<float>((size_t)feat.n * feat.c);
  channel_mean(c, feat, gap);
  float* att = c.alloc_t<float>((size_t)feat.n * l.out);
  linear_small(c, gap, feat.c, feat.n, l, att, l.out, 0, act_out);
  return att;
}
static Tens scale_add(Ctx& c, const Tens& x, const float* att, float att_bias, const float* vec, const Tens* add) {
  Tens y = c.new_tens(x.n, x.h, x.w, x.c);
  const long long total = x.rows() * x.c;
  parse_scale_add_kernel<<<grid_for(total), 256, 0, c.stream>>>(x.p, att, att_bias, vec, add ? add->p : nullptr, y.p, total,
                                                               x.h * x.w, x.c);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  return y;
}

// img01 [B,3,H,W] fp32 in [0,1] (H, W multiples of 32) -> logits8 fp32 [B,19,H/8,W/8] (optional), seg19 / seg12 uint8
void face_parse(Ctx& c, FaceParser& m, const float* img01, int B, int H, int W, float* logits8, uint8_t* seg19,
                uint8_t* seg12) {
  RFB_CHECK(H % 32 == 0 && W % 32 == 0, "face parser: image sides must be multiples of 32");
  const size_t mk = c.mark();
  Tens x = c.new_tens(B, H, W, 3);
  parse_preproc_kernel<<<grid_for(x.rows() * 3), 256, 0, c.stream>>>(img01, x.p, B, H * W);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  x = cbr(c, x, m.stem, 2, 3);  // 7x7 stride 2 pad 3
  {
    const int Ho = (x.h + 2 - 3) / 2 + 1, Wo = (x.w + 2 - 3) / 2 + 1;
    Tens y = c.new_tens(B, Ho, Wo, x.c);
    maxpool3s2_kernel<<<grid_for(y.rows() * (x.c / 8)), 256, 0, c.stream>>>(x.p, y.p, B, x.h, x.w, x.c, Ho, Wo);
    CUDA_OK(cudaGetLastError());
    c.launches++;
    x = y;
  }
  Tens feats[4];
  for (size_t i = 0; i < m.blocks.size(); ++i) {
    const ParseBlock& b = m.blocks[i];
    Tens sc = x;
    if (b.down) {
      Epi e;
      e.bias = b.ds.bias;
      sc = conv3x3_t(c, x, b.ds.w, e, b.stride, 0, 0, 0, 0);
    }
    Tens r = cbr(c, x, b.c1, b.stride);
    Epi e2;
    e2.bias = b.c2.bias, e2.res = sc.p, e2.ldr = sc.c, e2.relu_after_res = 1;  // relu(shortcut + bn2(conv2(r)))
    x = conv3x3_t(c, r, b.c2.w, e2);
    if (i & 1) feats[i >> 1] = x;
  }
  const Tens &f8 = feats[1], &f16 = feats[2], &f32 = feats[3];
  // ContextPath.forward, model.py:106-131
  float* avg = channel_att(c, f32, m.conv_avg, /*relu*/ 3);  // conv_avg(avg_pool(feat32)): [B,128]
  Tens a32 = cbr(c, f32, m.arm32);
  Tens f32s = scale_add(c, a32, channel_att(c, a32, m.arm32_att, /*sigmoid*/ 4), 0.f, avg, nullptr);
  Tens f32u = cbr(c, upsample2x(c, f32s), m.head32);
  Tens a16 = cbr(c, f16, m.arm16);
  Tens f16s = scale_add(c, a16, channel_att(c, a16, m.arm16_att, 4), 0.f, nullptr, &f32u);
  Tens f16u = cbr(c, upsample2x(c, f16s), m.head16);
  // FeatureFusionModule, model.py:200-212: feat * atten + feat = feat * (atten + 1)
  Tens feat = cbr(c, concat_c(c, f8, f16u), m.ffm_blk);
  float* gap = c.alloc_t<float>((size_t)B * feat.c);
  channel_mean(c, feat, gap);
  float* h1 = c.alloc_t<float>((size_t)B * m.ffm1.out);
  float* att = c.alloc_t<float>((size_t)B * m.ffm2.out);
  linear_small(c, gap, feat.c, B, m.ffm1, h1, m.ffm1.out, 0, /*relu*/ 3);
  linear_small(c, h1, m.ffm1.out, B, m.ffm2, att, m.ffm2.out, 0, /*sigmoid*/ 4);
  Tens fuse = scale_add(c, feat, att, 1.0f, nullptr, nullptr);
  // BiSeNetOutput, model.py:49-52: fp32 NCHW logits at 1/8 resolution
  Tens o = cbr(c, fuse, m.out_cbr);
  const int h = o.h, w = o.w;
  float* lg = logits8 ? logits8 : c.alloc_t<float>((size_t)B * m.n_classes * h * w);
  {
    Epi e;
    const long long hw = (long long)h * w;
    e.out32 = lg, e.o32_sn = m.n_classes * hw, e.o32_sp = 1, e.o32_sc = hw, e.o32_rpn = (int)hw;
    conv3x3_t(c, o, m.out_conv, e, 1, 0, 0, 0, 0);
  }
  if (seg19 || seg12) {
    parse_upsample_argmax_kernel<19><<<grid_for((long long)B * H * W), 256, 0, c.stream>>>(lg, seg19, seg12, B, h, w, H, W);
    CUDA_OK(cudaGetLastError());
    c.launches++;
  }
  c.release(mk);
}

void inpaint_from_parsing(Ctx& c, const float* img, const uint8_t* seg12, unsigned remove_bits, int B, int H, int W,
                          float* mask, float* inpaint) {
  parse_inpaint_kernel<<<grid_for((long long)B * H * W), 256, 0, c.stream>>>(img, seg12, remove_bits, mask, inpaint, B, H * W);
  CUDA_OK(cudaGetLastError());
  c.launches++;
}

}  // namespace rfb
