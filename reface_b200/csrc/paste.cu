// Paste-back: scripts/inference_swap_video.py:702-721 without the PIL round trip.
//   x_sample = 255. * rearrange(x, 'c h w -> h w c'); Image.fromarray(x_sample.astype(np.uint8))        :702-706
//   .resize((1024, 1024), Image.BILINEAR)                                                              :706
//   .convert('RGBA'), putalpha(255), .transform(orig.size, PERSPECTIVE, inv_coeffs, BILINEAR)           :716-720
//   pasted_image.alpha_composite(projected)                                                            :721
// The arithmetic is Pillow's (src/libImaging/Resample.c, Geometry.c, AlphaComposite.c): fixed-point two-pass triangle
// resampling with PRECISION_BITS = 22 and a uint8 intermediate; perspective source coordinates and the bilinear filter
// in double precision with truncation to uint8; alpha is exactly 0 or 255, so the composite is a select.  Everything
// is integer / IEEE-double work evaluated in Pillow's operation order (no FMA contraction): results are bit exact.
#include <algorithm>
#include <cmath>

#include "models.h"

namespace rfb {

// [B,3,h,w] fp32 -> [B,h,w,3] uint8: (uint8)(255.f * x), float32 multiply, truncation (numpy astype)
__global__ void f32chw_to_u8hwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int HW) {
  const long long total = (long long)B * HW * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 3);
    const long long p = i / 3;
    const long long b = p / HW, px = p % HW;
    out[i] = (uint8_t)(int)__fmul_rn(255.0f, x[(b * 3 + ch) * HW + px]);
  }
}
// one pass of ImagingResample{Horizontal,Vertical}_8bpc: dst[.., o, ..] = clip8((2^21 + sum_k src[.., lo+k, ..] * kk[o][k]) >> 22)
// src viewed as [outer, n_in, inner] bytes, dst as [outer, n_out, inner]
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, int ksize, long long outer, int n_in, int n_out, int inner) {
  const long long total = outer * n_out * inner;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int in_i = (int)(i % inner);
    const int o = (int)((i / inner) % n_out);
    const long long ou = i / ((long long)inner * n_out);
    const int lo = bounds[2 * o], n = bounds[2 * o + 1];
    int ss = 1 << 21;
    const uint8_t* p = src + (ou * n_in + lo) * inner + in_i;
    for (int k = 0; k < n; ++k) ss += (int)p[(long long)k * inner] * kk[o * ksize + k];
    ss >>= 22;
    dst[i] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}
// ImagingGenericTransform(perspective_transform, bilinear_filter32RGB) + alpha composite over the original frame
__global__ void perspective_paste_kernel(const uint8_t* __restrict__ big, const uint8_t* __restrict__ orig,
                                         const double* __restrict__ coef, uint8_t* __restrict__ out, int B, int h, int w,
                                         int H, int W) {
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    const double* a = coef + b * 8;
    const double xin = (double)X + 0.5, yin = (double)Y + 0.5;
    const double den = __dadd_rn(__dadd_rn(__dmul_rn(a[6], xin), __dmul_rn(a[7], yin)), 1.0);
    double sx = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[0], xin), __dmul_rn(a[1], yin)), a[2]), den);
    double sy = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[3], xin), __dmul_rn(a[4], yin)), a[5]), den);
    uint8_t* o = out + i * 3;
    const uint8_t* og = orig + i * 3;
    if (sx < 0.0 || sx >= (double)w || sy < 0.0 || sy >= (double)h || !(sx == sx) || !(sy == sy)) {
      o[0] = og[0], o[1] = og[1], o[2] = og[2];  // alpha 0: the frame shows through
      continue;
    }
    sx = __dsub_rn(sx, 0.5), sy = __dsub_rn(sy, 0.5);
    const int x = sx < 0.0 ? (int)floor(sx) : (int)sx;
    const int y = sy < 0.0 ? (int)floor(sy) : (int)sy;
    const double dx = __dsub_rn(sx, (double)x), dy = __dsub_rn(sy, (double)y);
    const int x0 = min(max(x, 0), w - 1), x1 = min(max(x + 1, 0), w - 1);
    const int y0 = min(max(y, 0), h - 1);
    const bool has2 = (y + 1 >= 0) && (y + 1 < h);
    const uint8_t* r0 = big + (b * h + y0) * (long long)w * 3;
    const uint8_t* r1 = big + (b * h + (has2 ? y + 1 : y0)) * (long long)w * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const double p00 = (double)r0[x0 * 3 + ch], p01 = (double)r0[x1 * 3 + ch];
      const double v1 = __dadd_rn(p00, __dmul_rn(__dsub_rn(p01, p00), dx));
      double v2 = v1;
      if (has2) {
        const double p10 = (double)r1[x0 * 3 + ch], p11 = (double)r1[x1 * 3 + ch];
        v2 = __dadd_rn(p10, __dmul_rn(__dsub_rn(p11, p10), dx));
      }
      o[ch] = (uint8_t)(int)__dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
    }
  }
}

// precompute_coeffs (triangle filter, support 1) + normalize_coeffs_8bpc of Pillow's Resample.c
static void pil_bilinear_coeffs(int insize, int outsize, std::vector<int>& bounds, std::vector<int>& kk, int& ksize) {
  const double scale = (double)insize / (double)outsize;
  const double fs = std::max(scale, 1.0);
  const double support = 1.0 * fs;
  ksize = (int)std::ceil(support) * 2 + 1;
  bounds.assign((size_t)outsize * 2, 0);
  kk.assign((size_t)outsize * ksize, 0);
  const double ss = 1.0 / fs;
  std::vector<double> w((size_t)ksize);
  for (int xx = 0; xx < outsize; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > insize) xmax = insize;
    xmax -= xmin;
    double ww = 0.0;
    std::fill(w.begin(), w.end(), 0.0);
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (int x = 0; x < ksize; ++x)
      kk[(size_t)xx * ksize + x] = w[x] < 0 ? (int)(-0.5 + w[x] * (double)(1 << 22)) : (int)(0.5 + w[x] * (double)(1 << 22));
    bounds[2 * xx] = xmin, bounds[2 * xx + 1] = xmax;
  }
}

static void resample_pass(Ctx& c, const uint8_t* src, uint8_t* dst, long long outer, int n_in, int n_out, int inner) {
  std::vector<int> bounds, kk;
  int ksize = 0;
  pil_bilinear_coeffs(n_in, n_out, bounds, kk, ksize);
  int* db = c.alloc_t<int>(bounds.size());
  int* dk = c.alloc_t<int>(kk.size());
  CUDA_OK(cudaMemcpyAsync(db, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  CUDA_OK(cudaMemcpyAsync(dk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  CUDA_OK(cudaStreamSynchronize(c.stream));  // the host tables go out of scope
  resample_u8_kernel<<<grid_for(outer * n_out * inner), 256, 0, c.stream>>>(src, dst, db, dk, ksize, outer, n_in, n_out, inner);
  CUDA_OK(cudaGetLastError());
  c.launches++;
}

// x01 [B,3,h,w] fp32 (device), orig [B,H,W,3] uint8 (device), coeffs [B,8] double (host) -> out [B,H,W,3] uint8 (device)
void paste_back(Ctx& c, const float* x01, const uint8_t* orig, const double* coeffs, int B, int h, int w, int up, int H, int W,
                uint8_t* out) {
  const size_t mk = c.mark();
  uint8_t* small = c.alloc_t<uint8_t>((size_t)B * h * w * 3);
  f32chw_to_u8hwc_kernel<<<grid_for((long long)B * h * w * 3), 256, 0, c.stream>>>(x01, small, B, h * w);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  const uint8_t* big = small;
  int bh = h, bw = w;
  if (up > 0 && (up != h || up != w)) {  // Image.resize((up, up), BILINEAR): horizontal pass, then vertical pass
    uint8_t* tmp = c.alloc_t<uint8_t>((size_t)B * h * up * 3);
    uint8_t* bigbuf = c.alloc_t<uint8_t>((size_t)B * up * up * 3);
    resample_pass(c, small, tmp, (long long)B * h, w, up, 3);
    resample_pass(c, tmp, bigbuf, B, h, up, up * 3);
    big = bigbuf, bh = up, bw = up;
  }
  double* dc = c.alloc_t<double>((size_t)B * 8);
  CUDA_OK(cudaMemcpyAsync(dc, coeffs, (size_t)B * 8 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  perspective_paste_kernel<<<grid_for((long long)B * H * W), 256, 0, c.stream>>>(big, orig, dc, out, B, bh, bw, H, W);
  CUDA_OK(cudaGetLastError());
  c.launches++;
  CUDA_OK(cudaStreamSynchronize(c.stream));  // `coeffs` is a host array owned by the caller
  c.release(mk);
}

}  // namespace rfb
