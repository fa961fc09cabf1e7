// HBM-bound helper kernels (layout changes, normalisation, softmax, small GEMV, DDIM update).
// All activations are NHWC fp16 ([rows = n*h*w, C]); statistics, schedules and latents stay fp32.
#pragma once
#include "ptx.cuh"

namespace rfb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// x * sigmoid(x) in five instructions: FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL.  __expf / __fdividef expand to ~12 (range
// fix-ups around the non-ftz ex2 and the division); the ftz forms need none: ex2 -> +inf gives rcp -> 0 and x * 0 = -0
// for very negative x, ex2 -> 0 gives exactly x for very positive x.  Relative error ~3 ulp, far below the fp16 output.
__device__ __forceinline__ float fast_silu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// SiLU of two values with ONE reciprocal: 1/(1+e0) = (1+e1) * r, 1/(1+e1) = (1+e0) * r with r = 1/((1+e0)(1+e1)):
// 1.5 MUFU operations per element instead of 2 (the streaming GroupNorm apply kernel is MUFU-heavy: XU pipe 62 %,
// profiles/r02_s2_gn_apply3_v2_ncu_full.txt).  Inputs are clamped at -40 so that the product stays finite
// ((1+2^58)^2 < 2^127); silu(-40) = -1.7e-16 either way.
__device__ __forceinline__ void fast_silu2(float& x0, float& x1) {
  float e0, e1, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaxf(x0, -40.0f) * -1.4426950408889634f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaxf(x1, -40.0f) * -1.4426950408889634f));
  const float d0 = 1.0f + e0, d1 = 1.0f + e1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  x0 = x0 * (d1 * r);
  x1 = x1 * (d0 * r);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ layout
// fp32 NCHW -> fp16 NHWC (channels padded with zeros up to Cp)
__global__ void nchw_f32_to_nhwc_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int N, int C,
                                            int HW, int Cp) {
  const long long total = (long long)N * HW * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long p = i / Cp;
    const int n = (int)(p / HW);
    const int px = (int)(p % HW);
    dst[i] = (c < C) ? __float2half_rn(src[((long long)n * C + c) * HW + px]) : __float2half_rn(0.f);
  }
}
// fp16 NHWC -> fp32 NCHW
__global__ void nhwc_f16_to_nchw_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, int N, int C,
                                            int HW, int ld) {
  const long long total = (long long)N * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    dst[i] = __half2float(src[((long long)n * HW + px) * ld + c]);
  }
}

// Generic im2col for the convolutions that do not fit the TMA implicit-GEMM tile (stride 2, asymmetric
// padding, tiny Cin, non power-of-two maps): dst[m, (ky*KW+kx)*C + c], row stride Kp (zero padded).
__global__ void im2col_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int N, int H, int W, int C,
                              int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo, int Kp) {
  const int taps = KH * KW;
  const int K = taps * C;
  const long long M = (long long)N * Ho * Wo;
  if ((C & 7) == 0) {
    const int cv = C >> 3;
    const int kv = Kp >> 3;
    const long long total = M * kv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int kc = (int)(i % kv);
      const long long m = i / kv;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (kc * 8 < K) {
        const int tap = kc / cv, c8 = kc % cv;
        const int ky = tap / KW, kx = tap % KW;
        const int ox = (int)(m % Wo);
        const long long t = m / Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const int iy = oy * stride - pad_t + ky, ix = ox * stride - pad_l + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          val = *reinterpret_cast<const uint4*>(src + (((long long)n * H + iy) * W + ix) * C + c8 * 8);
      }
      *reinterpret_cast<uint4*>(dst + m * Kp + kc * 8) = val;
    }
  } else {
    const long long total = M * Kp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int k = (int)(i % Kp);
      const long long m = i / Kp;
      __half val = __float2half_rn(0.f);
      if (k < K) {
        const int tap = k / C, c = k % C;
        const int ky = tap / KW, kx = tap % KW;
        const int ox = (int)(m % Wo);
        const long long t = m / Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const int iy = oy * stride - pad_t + ky, ix = ox * stride - pad_l + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = src[(((long long)n * H + iy) * W + ix) * C + c];
      }
      dst[i] = val;
    }
  }
}

// nearest 2x upsample, NHWC, C % 8 == 0
__global__ void upsample2x_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int N, int H, int W, int C) {
  const int cv = C >> 3;
  const long long total = (long long)N * (2 * H) * (2 * W) * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    long long p = i / cv;
    const int ox = (int)(p % (2 * W));
    p /= (2 * W);
    const int oy = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    reinterpret_cast<uint4*>(dst)[i] =
        *reinterpret_cast<const uint4*>(src + (((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * C + c8 * 8);
  }
}

// channel concat of two NHWC tensors ([rows,Ca] ++ [rows,Cb]); Ca, Cb % 8 == 0
__global__ void concat_c_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ dst,
                                long long rows, int Ca, int Cb) {
  const int cv = (Ca + Cb) >> 3, av = Ca >> 3;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const long long r = i / cv;
    reinterpret_cast<uint4*>(dst)[i] = (c8 < av) ? reinterpret_cast<const uint4*>(a)[r * av + c8]
                                                 : reinterpret_cast<const uint4*>(b)[r * (cv - av) + (c8 - av)];
  }
}

// ------------------------------------------------------------------ GroupNorm (fp32 statistics)
// All GroupNorm kernels read their input through GnSrc: one NHWC tensor [N, HW, C], or the channel concatenation of
// two ([.., C1] ++ [.., C - C1], torch.cat of the UNet skip connections) without materialising it; the second source
// may hold only n2mod samples (sample n reads n mod n2mod: a tensor shared by the two CFG halves).
struct GnSrc {
  const __half* x1;
  const __half* x2;  // nullptr: single source (C1 == C)
  int C1, n2mod;
};
// base pointer / row stride (in elements) of the 8-channel vector `cq` of sample n
__device__ __forceinline__ const __half* gn_src_ptr(const GnSrc& s, int n, int HW, int C, int cq, int& stride) {
  const int cv1 = s.C1 >> 3;
  if (s.x2 == nullptr || cq < cv1) {
    stride = s.C1;
    return s.x1 + (long long)n * HW * s.C1 + cq * 8;
  }
  stride = C - s.C1;
  const int n2 = s.n2mod > 0 ? n % s.n2mod : n;
  return s.x2 + (long long)n2 * HW * stride + (cq - cv1) * 8;
}

// Two-launch GroupNorm (whole-grid): gn_stats2 = per-slab statistics with 8 batched 16-byte loads in flight per thread
// and bank-conflict-free partial sums; gn_apply2 = normalise (+SiLU) with the finalize step (fixed-order fold of the
// per-slab partials) done by every CTA for its own sample instead of a separate launch.
__global__ void __launch_bounds__(512)
gn_stats2_kernel(const GnSrc src, float* __restrict__ stats, int HW, int C, int slab, int G) {
  extern __shared__ float sh[];  // [R][16][cv] + [2*C]
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int cq = threadIdx.x % cv, pr = threadIdx.x / cv;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * slab;
  const int p1 = min(HW, p0 + slab);
  int xs;
  const __half* xn = gn_src_ptr(src, n, HW, C, cq, xs);
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  for (int pb = p0 + pr; pb < p1; pb += 8 * R) {
    uint4 u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int p = pb + k * R;
      u[k] = make_uint4(0u, 0u, 0u, 0u);
      if (p < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(xn + (long long)p * xs));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = unpack_h2(w[t]);
        s[2 * t] += f.x;
        ss[2 * t] = fmaf(f.x, f.x, ss[2 * t]);
        s[2 * t + 1] += f.y;
        ss[2 * t + 1] = fmaf(f.y, f.y, ss[2 * t + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[((size_t)pr * 16 + 2 * j) * cv + cq] = s[j];
    sh[((size_t)pr * 16 + 2 * j + 1) * cv + cq] = ss[j];
  }
  __syncthreads();
  float* chs = sh + (size_t)R * 2 * C;
  for (int i = threadIdx.x; i < 16 * cv; i += blockDim.x) {
    const int k = i / cv, c8 = i - k * cv;
    float a = 0.f;
    for (int rr = 0; rr < R; ++rr) a += sh[((size_t)rr * 16 + k) * cv + c8];
    chs[2 * (c8 * 8 + (k >> 1)) + (k & 1)] = a;
  }
  __syncthreads();
  const int cpg = C / G;
  if (threadIdx.x < 2 * G) {
    const int gi = threadIdx.x >> 1, which = threadIdx.x & 1;
    float a = 0.f;
    for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) a += chs[2 * c + which];
    stats[((long long)n * gridDim.x + blockIdx.x) * 2 * G + threadIdx.x] = a;
  }
}
__global__ void __launch_bounds__(512)
gn_apply2_kernel(const GnSrc src, const float* __restrict__ partial, int nslab, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __half* __restrict__ y, int HW, int C, int G, float eps, int silu, int slab) {
  __shared__ float st[64];  // mean / rstd per group (G <= 32)
  const int n = blockIdx.y;
  const int cpg = C / G;
  {  // finalize (same order as gn_finalize_kernel): 8 threads per group fold slabs t, t+8, ..., then a fixed shuffle tree
    const int gi = threadIdx.x >> 3, t = threadIdx.x & 7;
    float s = 0.f, ss = 0.f;
    if (gi < G) {
      for (int sl = t; sl < nslab; sl += 8) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(partial + (((long long)n * nslab + sl) * G + gi) * 2));
        s += v.x;
        ss += v.y;
      }
    }
    if (threadIdx.x < 256) {  // whole warps: 8 * G <= 256
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
      }
    }
    if (gi < G && t == 0) {
      const float cnt = (float)cpg * (float)HW;
      const float mean = s / cnt;
      const float var = fmaxf(ss / cnt - mean * mean, 0.f);
      st[2 * gi] = mean;
      st[2 * gi + 1] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int cq = threadIdx.x % cv, pr = threadIdx.x / cv;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cq * 8 + j;
    const int gi = c / cpg;
    a[j] = st[2 * gi + 1] * __ldg(gamma + c);
    b[j] = __ldg(beta + c) - st[2 * gi] * a[j];
  }
  const int p0 = blockIdx.x * slab;
  const int p1 = min(HW, p0 + slab);
  int xs;
  const __half* xn = gn_src_ptr(src, n, HW, C, cq, xs);
  __half* yn = y + (long long)n * HW * C + cq * 8;
  for (int pb = p0 + pr; pb < p1; pb += 4 * R) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = pb + k * R;
      if (p < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(xn + (long long)p * xs));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = pb + k * R;
      if (p < p1) {
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_h2(w[t]);
          float v0 = fmaf(f.x, a[2 * t], b[2 * t]), v1 = fmaf(f.y, a[2 * t + 1], b[2 * t + 1]);
          if (silu) v0 = fast_silu(v0), v1 = fast_silu(v1);
          o[t] = pack_h2(v0, v1);
        }
        *reinterpret_cast<uint4*>(yn + (long long)p * C) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// GroupNorm from PRODUCER statistics: the GEMM / conv epilogue that wrote the tensor also left per-(32 rows, channel)
// partial sums (gemm_epilogue.cuh, GemmArgs::stats: [rows / 32][C][2]).  gn_finalize3 folds them per (sample, group) in a
// fixed order -> mean / rstd [N][G][2]; gn_apply3 is then a pure streaming pass (one read, one write).
struct GnStatSrc {
  const float* s1;
  const float* s2;  // partials of the second concatenated tensor (nullptr: single source)
  int C1, n2mod;
};
__global__ void __launch_bounds__(256)
gn_finalize3_kernel(const GnStatSrc st, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float2* __restrict__ ab, int P, int HW, int C, int G, float eps) {
  pdl_wait();
  // grid (G, N): one block folds ONE group (cpg <= 256 contiguous channels) of one sample.  Thread t owns channel
  // (t mod cpg) and partial rows t / cpg, t / cpg + R, ... (four independent loads in flight): consecutive threads read
  // consecutive float2, every thread's own sum runs in row order, the R row-phases and then the cpg channels are folded
  // in a fixed order -> bitwise reproducible, independent of the batch size.
  // Output: per (sample, channel) the affine map of the normalisation, ab = (rstd * gamma, beta - mean * rstd * gamma).
  __shared__ float2 red[256];
  __shared__ float2 mr;
  const int n = blockIdx.y, gi = blockIdx.x;
  const int cpg = C / G;
  const int R = 256 / cpg;
  const int ci = threadIdx.x % cpg, rp = threadIdx.x / cpg;
  float s = 0.f, ss = 0.f;
  if (rp < R) {
    const int ch = gi * cpg + ci;
    const bool first = st.s2 == nullptr || ch < st.C1;
    const int Cs = first ? st.C1 : C - st.C1;
    const float2* base = first ? reinterpret_cast<const float2*>(st.s1) + (long long)n * P * Cs + ch
                               : reinterpret_cast<const float2*>(st.s2) +
                                     (long long)(st.n2mod > 0 ? n % st.n2mod : n) * P * Cs + (ch - st.C1);
    int pr = rp;
    for (; pr + 3 * R < P; pr += 4 * R) {
      const float2 v0 = __ldg(base + (long long)pr * Cs), v1 = __ldg(base + (long long)(pr + R) * Cs);
      const float2 v2 = __ldg(base + (long long)(pr + 2 * R) * Cs), v3 = __ldg(base + (long long)(pr + 3 * R) * Cs);
      s += v0.x, ss += v0.y;
      s += v1.x, ss += v1.y;
      s += v2.x, ss += v2.y;
      s += v3.x, ss += v3.y;
    }
    for (; pr < P; pr += R) {
      const float2 v = __ldg(base + (long long)pr * Cs);
      s += v.x, ss += v.y;
    }
  }
  red[threadIdx.x] = make_float2(s, ss);
  __syncthreads();
  if (threadIdx.x < cpg) {  // row phases of one channel, in phase order
    float a = 0.f, b = 0.f;
    for (int r = 0; r < R; ++r) {
      const float2 v = red[r * cpg + threadIdx.x];
      a += v.x, b += v.y;
    }
    red[threadIdx.x] = make_float2(a, b);  // phase 0's slot of this channel: read only by this thread above
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // channels of the group, in channel order
    float a = 0.f, b = 0.f;
    for (int c = 0; c < cpg; ++c) a += red[c].x, b += red[c].y;
    const float cnt = (float)cpg * (float)HW;
    const float mean = a / cnt;
    const float var = fmaxf(b / cnt - mean * mean, 0.f);
    mr = make_float2(mean, rsqrtf(var + eps));
  }
  __syncthreads();
  if (threadIdx.x < cpg) {
    const int ch = gi * cpg + threadIdx.x;
    const float a = mr.y * __ldg(gamma + ch);
    ab[(long long)n * C + ch] = make_float2(a, __ldg(beta + ch) - mr.x * a);
  }
}
// y = x * a[n, c] + b[n, c] (* sigmoid): one read, one write.  A block streams `slab` pixels of one sample; a thread keeps
// the affine map of its 8 channels in registers and has UB 16-byte loads in flight.
__global__ void __launch_bounds__(512, 2)  // <= 64 registers: the 90-register first version ran one block per SM and was
                                           // latency-bound at 21 % warp occupancy (profiles/r02_gn3_ncu_full.txt)
gn_apply3_kernel(const GnSrc src, const float2* __restrict__ ab, __half* __restrict__ y, int HW, int C, int silu, int slab) {
  pdl_wait();
  const int n = blockIdx.y;
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int cq = threadIdx.x % cv, pr = threadIdx.x / cv;
  float a[8], b[8];
  {
    const float4* p = reinterpret_cast<const float4*>(ab + (long long)n * C + cq * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = __ldg(p + j);
      a[2 * j] = v.x, b[2 * j] = v.y, a[2 * j + 1] = v.z, b[2 * j + 1] = v.w;
    }
  }
  const int p0 = blockIdx.x * slab;
  const int p1 = min(HW, p0 + slab);
  int xs;
  const __half* xn = gn_src_ptr(src, n, HW, C, cq, xs);
  __half* yn = y + (long long)n * HW * C + cq * 8;
  constexpr int UB = 4;
  for (int pb = p0 + pr; pb < p1; pb += UB * R) {
    uint4 u[UB];
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const int p = pb + k * R;
      if (p < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(xn + (long long)p * xs));
    }
#pragma unroll
    for (int k = 0; k < UB; ++k) {
      const int p = pb + k * R;
      if (p < p1) {
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_h2(w[t]);
          float v0 = fmaf(f.x, a[2 * t], b[2 * t]), v1 = fmaf(f.y, a[2 * t + 1], b[2 * t + 1]);
          if (silu) fast_silu2(v0, v1);
          o[t] = pack_h2(v0, v1);
        }
        *reinterpret_cast<uint4*>(yn + (long long)p * C) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// An activation-free GroupNorm in front of a 1x1 convolution / linear layer (SpatialTransformer.norm -> proj_in,
// attention.py:256-262,281-283) is a per-(sample, channel) affine map, so it folds into the layer:
//   W (x * a_n + b_n) + bias = (W diag(a_n)) x + (W b_n + bias)
// -> per-sample fp16 weights Wn[n] = W * a_n (columns scaled) and a per-sample bias row; the GEMM then reads the RAW
// tensor and the normalisation pass over it disappears.  One warp per (sample, output row); fp32 originals of W.
__global__ void gn_fold_weights_kernel(const float* __restrict__ w32, const float* __restrict__ bias,
                                       const float2* __restrict__ ab, __half* __restrict__ Wn, float* __restrict__ biasn,
                                       int Cin, int Cout, int kp, int cout_p) {
  pdl_wait();
  const int n = blockIdx.y;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= cout_p) return;
  __half* dst = Wn + ((long long)n * cout_p + o) * kp;
  float acc = 0.f;
  for (int k = lane; k < kp; k += 32) {
    float v = 0.f;
    if (o < Cout && k < Cin) {
      const float w = __ldg(w32 + (long long)o * Cin + k);
      const float2 t = __ldg(ab + (long long)n * Cin + k);
      v = w * t.x;
      acc = fmaf(w, t.y, acc);
    }
    dst[k] = __float2half_rn(v);
  }
  acc = warp_sum(acc);
  if (lane == 0 && o < Cout) biasn[(long long)n * Cout + o] = acc + (bias ? bias[o] : 0.f);
}

// Single-launch GroupNorm: the CTAs of one sample (gridDim.x = cluster size, 8 or 16) form a thread-block cluster.  Each CTA reduces its pixel
// slab (fp32 sums, fixed order), the per-group partials are exchanged through distributed shared memory and folded in
// rank order (bitwise reproducible and independent of the batch size), then the CTA normalises the slab it has just
// read (second read served by L2).  HBM traffic: x once + y once; one launch instead of stats / finalize / apply.
// blockDim = (C/8) * R as above; dynamic smem = ((R + 1) * 2C + 4G) floats.
__device__ __forceinline__ float dsmem_ld_f32(uint32_t saddr, uint32_t cta) {
  float v;
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %1, %2;\n"
      "ld.shared::cluster.f32 %0, [ra];\n"
      "}\n"
      : "=f"(v)
      : "r"(saddr), "r"(cta)
      : "memory");
  return v;
}
__global__ void __launch_bounds__(512, 2)
gn_fused_kernel(const GnSrc src, const float* __restrict__ gamma, const float* __restrict__ beta,
                __half* __restrict__ y, int HW, int C, int G, float eps, int silu) {
  extern __shared__ float sh[];
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int cq = threadIdx.x % cv, pr = threadIdx.x / cv;
  const int n = blockIdx.y;
  const int S = gridDim.x;  // == cluster size
  const int rank = blockIdx.x;
  const int per = (HW + S - 1) / S;
  const int p0 = rank * per;
  const int p1 = min(HW, p0 + per);
  float* chs = sh + max(R * 2 * C, 2 * G * S);  // [2C], behind the reduction scratch / the gathered cluster partials
  float* part = chs + 2 * C;            // [2G] this CTA's per-group sums (read by the whole cluster)
  float* stat = part + 2 * G;           // [2G] mean / rstd
  int xs;
  const __half* xn = gn_src_ptr(src, n, HW, C, cq, xs);
  constexpr int UB = 4;  // 16-byte loads in flight per thread (two CTAs per SM: 64 registers)
  {
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    for (int pb = p0 + pr; pb < p1; pb += UB * R) {
      uint4 u[UB];
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        const int p = pb + k * R;
        u[k] = make_uint4(0u, 0u, 0u, 0u);  // zeros add nothing to either sum
        if (p < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(xn + (long long)p * xs));
      }
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_h2(w[t]);
          s[2 * t] += f.x;
          ss[2 * t] = fmaf(f.x, f.x, ss[2 * t]);
          s[2 * t + 1] += f.y;
          ss[2 * t + 1] = fmaf(f.y, f.y, ss[2 * t + 1]);
        }
      }
    }
    // sh[pr][k][cq], k = 2 * j + (0: sum, 1: sum of squares): consecutive lanes -> consecutive banks
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sh[((size_t)pr * 16 + 2 * j) * cv + cq] = s[j];
      sh[((size_t)pr * 16 + 2 * j + 1) * cv + cq] = ss[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * cv; i += blockDim.x) {
    const int k = i / cv, c8 = i - k * cv;
    float a = 0.f;
    for (int rr = 0; rr < R; ++rr) a += sh[((size_t)rr * 16 + k) * cv + c8];
    chs[2 * (c8 * 8 + (k >> 1)) + (k & 1)] = a;
  }
  __syncthreads();
  const int cpg = C / G;
  if (threadIdx.x < 2 * G) {
    const int gi = threadIdx.x >> 1, which = threadIdx.x & 1;
    float a = 0.f;
    for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) a += chs[2 * c + which];
    part[threadIdx.x] = a;
  }
  cluster_sync_all();  // every CTA's partials are visible cluster-wide
  // gather all S x 2G partials with independent remote loads (sh is free again), then fold them in rank order
  for (int i = threadIdx.x; i < 2 * G * S; i += blockDim.x)
    sh[i] = dsmem_ld_f32(smem_u32(part + (i % (2 * G))), (uint32_t)(i / (2 * G)));
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // done reading the peers' smem
  __syncthreads();
  if (threadIdx.x < 2 * G) {
    float a = 0.f;
    for (int rk = 0; rk < S; ++rk) a += sh[rk * 2 * G + threadIdx.x];
    chs[threadIdx.x] = a;  // group totals (sum, sum of squares interleaved)
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const float cnt = (float)cpg * (float)HW;
    const float mean = chs[2 * threadIdx.x] / cnt;
    const float var = fmaxf(chs[2 * threadIdx.x + 1] / cnt - mean * mean, 0.f);
    stat[2 * threadIdx.x] = mean;
    stat[2 * threadIdx.x + 1] = rsqrtf(var + eps);
  }
  __syncthreads();
  {
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cq * 8 + j;
      const int gi = c / cpg;
      a[j] = stat[2 * gi + 1] * __ldg(gamma + c);
      b[j] = __ldg(beta + c) - stat[2 * gi] * a[j];
    }
    __half* yn = y + (long long)n * HW * C + cq * 8;
    for (int pb = p0 + pr; pb < p1; pb += UB * R) {
      uint4 u[UB];
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        const int p = pb + k * R;
        if (p < p1) u[k] = __ldg(reinterpret_cast<const uint4*>(xn + (long long)p * xs));
      }
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        const int p = pb + k * R;
        if (p < p1) {
          const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
          uint32_t o[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_h2(w[t]);
            float v0 = fmaf(f.x, a[2 * t], b[2 * t]), v1 = fmaf(f.y, a[2 * t + 1], b[2 * t + 1]);
            if (silu) v0 = fast_silu(v0), v1 = fast_silu(v1);
            o[t] = pack_h2(v0, v1);
          }
          *reinterpret_cast<uint4*>(yn + (long long)p * C) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // peers may still be reading `part`
}

// ------------------------------------------------------------------ LayerNorm: one warp per row, C % 64 == 0, C <= 2048
__global__ void layernorm_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __half* __restrict__ y, long long rows, int C,
                                 long long ldx, long long ldy, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int npair = C >> 6;  // half2 per lane
  float2 v[32];
  const __half2* xp = reinterpret_cast<const __half2*>(x + row * ldx);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < npair) {
      v[k] = __half22float2(xp[lane + 32 * k]);
      s += v[k].x + v[k].y;
    }
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < npair) {
      const float a = v[k].x - mean, b = v[k].y - mean;
      ss += a * a + b * b;
    }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
  __half2* yp = reinterpret_cast<__half2*>(y + row * ldy);
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < npair) {
      const int c = 2 * (lane + 32 * k);
      const float a = (v[k].x - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      const float b = (v[k].y - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
      yp[lane + 32 * k] = __floats2half2_rn(a, b);
    }
}

// Vectorised LayerNorm: LPR lanes share a row (32/LPR rows per warp), every lane keeps VPL 16-byte vectors of the row
// in registers (C = 8 * LPR * VPL).  Global traffic is 16 B per lane and fully coalesced (a warp instruction covers
// 32/LPR contiguous 128-byte-aligned row segments); same two-pass fp32 statistics as layernorm_kernel.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256)
layernorm_vec_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __half* __restrict__ y, long long rows, long long ldx, long long ldy, float eps) {
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  constexpr int C = 8 * LPR * VPL;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = ((long long)blockIdx.x * 8 + warp) * RPW + lane / LPR;
  const int sub = lane % LPR;
  const bool ok = row < rows;
  const uint4* xp = reinterpret_cast<const uint4*>(x + (ok ? row : 0) * ldx);
  uint4 u[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) u[k] = __ldg(xp + sub + LPR * k);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_h2(w[t]);
      s += f.x + f.y;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / (float)C);
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_h2(w[t]);
      const float a = f.x - mean, b = f.y - mean;
      ss = fmaf(a, a, ss);
      ss = fmaf(b, b, ss);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss * (1.0f / (float)C) + eps);
  uint4* yp = reinterpret_cast<uint4*>(y + (ok ? row : 0) * ldy);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c0 = 8 * (sub + LPR * k);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float2 f0 = unpack_h2(u[k].x), f1 = unpack_h2(u[k].y), f2 = unpack_h2(u[k].z), f3 = unpack_h2(u[k].w);
    uint4 o;
    o.x = pack_h2(fmaf((f0.x - mean) * rstd, g0.x, b0.x), fmaf((f0.y - mean) * rstd, g0.y, b0.y));
    o.y = pack_h2(fmaf((f1.x - mean) * rstd, g0.z, b0.z), fmaf((f1.y - mean) * rstd, g0.w, b0.w));
    o.z = pack_h2(fmaf((f2.x - mean) * rstd, g1.x, b1.x), fmaf((f2.y - mean) * rstd, g1.y, b1.y));
    o.w = pack_h2(fmaf((f3.x - mean) * rstd, g1.z, b1.z), fmaf((f3.y - mean) * rstd, g1.w, b1.w));
    if (ok) yp[sub + LPR * k] = o;
  }
}

// ------------------------------------------------------------------ row softmax (materialised attention path)
// x: [rows, ld] fp16 in place over the first L columns; columns [L, ld) are zeroed (K padding for the P.V GEMM).
__global__ void softmax_rows_kernel(__half* __restrict__ x, long long rows, int L, int ld) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  __half* p = x + row * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) mx = fmaxf(mx, __half2float(p[i]));
  const int nw = blockDim.x >> 5;
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int w = 0; w < nw; ++w) mx = fmaxf(mx, red[w]);  // every thread reduces the per-warp partials
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += __expf(__half2float(p[i]) - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];
  const float inv = 1.0f / s;
  for (int i = threadIdx.x; i < ld; i += blockDim.x)
    p[i] = (i < L) ? __float2half_rn(__expf(__half2float(p[i]) - mx) * inv) : __float2half_rn(0.f);
}

// Same contract, one WARP per row with the row held in registers (ld <= 32 * EPL): one read, one exp per element, shuffle
// reductions instead of three block-wide barriers.  Used for the short rows of the materialised path (L = 64 / 256 at the
// UNet's 8x8 / 16x16 levels, L = 257 in CLIP).
template <int EPL>
__global__ void softmax_rows_warp_kernel(__half* __restrict__ x, long long rows, int L, int ld) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  __half* p = x + row * ld;
  float v[EPL];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < L ? __half2float(p[i]) : -INFINITY;
    mx = fmaxf(mx, v[k]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < L ? __expf(v[k] - mx) : 0.f;
    s += v[k];
  }
  const float inv = 1.0f / warp_sum(s);
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const int i = lane + 32 * k;
    if (i < ld) p[i] = __float2half_rn(v[k] * inv);  // columns [L, ld) become 0
  }
}

// V section of a fused [N, L, ldq] projection -> Vt [N*heads, d, Lp] (keys contiguous), zero padded to Lp
__global__ void transpose_v_kernel(const __half* __restrict__ v, __half* __restrict__ vt, int N, int L, int heads, int d,
                                   long long ldq, int Lp, int hs) {
  __shared__ __half tile[32][33];
  const int z = blockIdx.z;  // n*heads + h
  const int n = z / heads, h = z % heads;
  const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int l = l0 + i, dd = d0 + threadIdx.x;
    tile[i][threadIdx.x] = (l < L && dd < d) ? v[((long long)n * L + l) * ldq + h * hs + dd] : __float2half_rn(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int dd = d0 + i, l = l0 + threadIdx.x;
    if (dd < d && l < Lp) vt[((long long)z * d + dd) * Lp + l] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ cross-attention with a short context (T <= 16)
// attention.py:204-221 for context length T > 1 (`stack_feat`-style conditioning): one thread per (token, head).
// q: [N*L, C] fp16 (to_q of LN2(x)); kc, vc: [N*T, C] fp32 (to_k / to_v of the context, staged in smem per sample);
// out: [N*L, C] fp16 = softmax_T(scale * q k^T) v.  The score row has only T entries: no tensor cores needed.
template <int TMAX>
__global__ void cross_attn_small_kernel(const __half* __restrict__ q, const float* __restrict__ kc,
                                        const float* __restrict__ vc, __half* __restrict__ out, int L, int T, int C,
                                        int heads, float scale) {
  pdl_wait();
  extern __shared__ float kv[];  // [2][T][C]
  const int n = blockIdx.y;
  const int d = C / heads;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    kv[i] = kc[(long long)n * T * C + i];
    kv[T * C + i] = vc[(long long)n * T * C + i];
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // token * heads + head
  if (idx >= L * heads) return;
  const int tok = idx / heads, h = idx - tok * heads;
  const __half* qp = q + ((long long)n * L + tok) * C + h * d;
  float s[TMAX];
#pragma unroll
  for (int t = 0; t < TMAX; ++t) s[t] = 0.f;
  for (int j = 0; j < d; j += 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(qp + j));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float qv[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_h2(w[e]);
      qv[2 * e] = f.x, qv[2 * e + 1] = f.y;
    }
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < T) {
        const float* kp = kv + t * C + h * d + j;
#pragma unroll
        for (int e = 0; e < 8; ++e) s[t] = fmaf(qv[e], kp[e], s[t]);
      }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < T) s[t] *= scale, mx = fmaxf(mx, s[t]);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < T) s[t] = __expf(s[t] - mx), sum += s[t];
  const float inv = 1.0f / sum;
  __half* op = out + ((long long)n * L + tok) * C + h * d;
  for (int j = 0; j < d; j += 8) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < T) {
        const float* vp = kv + (T + t) * C + h * d + j;
        const float pt = s[t] * inv;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(pt, vp[e], o[e]);
      }
    uint4 r;
    r.x = pack_h2(o[0], o[1]), r.y = pack_h2(o[2], o[3]), r.z = pack_h2(o[4], o[5]), r.w = pack_h2(o[6], o[7]);
    *reinterpret_cast<uint4*>(op + j) = r;
  }
}

// ------------------------------------------------------------------ small-M linear (fp32 in, fp32 weights)
// out[r, o] = act_out(bias[o] + sum_k act_in(x[r, k]) * W[o, k]) (+ res[r, o]) for r < R <= RMAX.
// One warp per output column; x is staged (activated once) through shared memory in 512-wide K chunks,
// weights are streamed with 16-byte loads.  K % 4 == 0.
template <int RMAX>
__global__ void linear_small_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ out, int R, int K, int O,
                                    long long ldx, long long ldo, int act_in, int act_out,
                                    const float* __restrict__ res) {
  pdl_wait();
  constexpr int KC = 512;
  __shared__ __align__(16) float xs[RMAX][KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  float acc[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) acc[r] = 0.f;
  const float* w = W + (long long)min(o, O - 1) * K;
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int kc = min(KC, K - k0);
    for (int idx = threadIdx.x; idx < R * KC; idx += blockDim.x) {
      const int r = idx / KC, kk = idx % KC;
      float v = 0.f;
      if (kk < kc) {
        v = x[r * ldx + k0 + kk];
        if (act_in == 1) v = v / (1.0f + __expf(-v));
      }
      xs[r][kk] = v;
    }
    __syncthreads();
    for (int kk = lane * 4; kk < kc; kk += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k0 + kk));
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < R) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[r][kk]);
          acc[r] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
        }
    }
    __syncthreads();
  }
  if (o >= O) return;
#pragma unroll
  for (int r = 0; r < RMAX; ++r) {
    if (r < R) {
      float v = warp_sum(acc[r]);
      if (lane == 0) {
        if (bias) v += bias[o];
        if (act_out == 1) v = v / (1.0f + __expf(-v));
        else if (act_out == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
        else if (act_out == 3) v = fmaxf(v, 0.f);
        else if (act_out == 4) v = 1.0f / (1.0f + __expf(-v));
        if (res) v += res[r * ldo + o];
        out[r * ldo + o] = v;
      }
    }
  }
}

// timestep_embedding (cos || sin), fp32  [ref: ldm/modules/diffusionmodules/util.py:151-171]
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, float* __restrict__ out, int N, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * half) return;
  const int n = i / half, j = i % half;
  const float freq = expf(-logf(10000.0f) * (float)j / (float)half);
  const float a = (float)t[n] * freq;
  out[n * dim + j] = cosf(a);
  out[n * dim + half + j] = sinf(a);
}

// ------------------------------------------------------------------ DDIM glue (fp32, bit-exact data movement)
// x9[n] = cat(x[n%B], z[n%B], mask[n%B]) for n < dup*B   [ref: ldm/models/diffusion/ddim.py:330,338]
__global__ void concat9_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ mask,
                               float* __restrict__ out, int B, int HW, int dup) {
  pdl_wait();
  const long long total = (long long)dup * B * 9 * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % 9);
    const int b = (int)((t / 9) % B);
    float v;
    if (c < 4) v = x[((long long)b * 4 + c) * HW + px];
    else if (c < 8) v = z[((long long)b * 4 + (c - 4)) * HW + px];
    else v = mask[(long long)b * HW + px];
    out[i] = v;
  }
}
// classifier-free guidance combine + DDIM x_{t-1} (same fp32 op order as ddim.py:346,363-374; no FMA contraction)
__global__ void cfg_ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eps2,
                                       const float* __restrict__ noise, float* __restrict__ x_prev,
                                       float* __restrict__ pred_x0, long long count, float scale, float a_t, float a_prev,
                                       float sigma, float sqrt_one_minus_at, int has_uncond) {
  const float sqrt_at = __fsqrt_rn(a_t);
  const float sqrt_aprev = __fsqrt_rn(a_prev);
  const float dir_coef = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, a_prev), __fmul_rn(sigma, sigma)));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float e;
    if (has_uncond) {
      const float eu = eps2[i], ec = eps2[count + i];
      e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(ec, eu)));
    } else {
      e = eps2[i];
    }
    const float p0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(sqrt_one_minus_at, e)), sqrt_at);
    const float dir = __fmul_rn(dir_coef, e);
    const float nz = noise ? __fmul_rn(sigma, noise[i]) : 0.0f;
    x_prev[i] = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_aprev, p0), dir), nz);
    if (pred_x0) pred_x0[i] = p0;
  }
}

// The UNet's output convolution (320 -> 4 channels, 3x3, openaimodel.py:829-836) runs as ONE thin GEMM over the 9 taps:
// taps[p, tap * 4 + co] = sum_c h[p, c] * W[co, c, tap] (fp32, [N*L*L, 36]); a pixel's eps is then the sum of the 9 tap
// partials of its 3x3 neighbourhood (zero outside the map = the conv's padding) plus the bias.  As a 3x3 implicit GEMM the
// layer moved 9x the activation bytes through 128x32 tiles of which 4 columns were real (55-73 us); the tap GEMM reads
// the activation once.
__device__ __forceinline__ float4 eps_gather(const float* __restrict__ taps, const float* __restrict__ bias, int n, int y,
                                             int x, int L) {
  float4 acc = make_float4(bias[0], bias[1], bias[2], bias[3]);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy >= 0 && yy < L && xx >= 0 && xx < L) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(taps + (((long long)n * L + yy) * L + xx) * 36 + tap * 4));
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  return acc;
}
// eps [N,4,L,L] fp32 from the tap partials (UNetModel.forward's return value)
__global__ void eps_from_taps_kernel(const float* __restrict__ taps, const float* __restrict__ bias, float* __restrict__ eps,
                                     int N, int L) {
  const long long total = (long long)N * L * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % L), y = (int)((i / L) % L), n = (int)(i / ((long long)L * L));
    const float4 e = eps_gather(taps, bias, n, y, x, L);
    float* o = eps + (long long)n * 4 * L * L + (long long)y * L + x;
    o[0] = e.x, o[(long long)L * L] = e.y, o[2ll * L * L] = e.z, o[3ll * L * L] = e.w;
  }
}
// p_sample_ddim's tail as the END of the output convolution: tap gather (+bias) for the uncond / cond halves ->
// e_t = e_u + s (e_c - e_u) (ddim.py:346) -> pred_x0, dir_xt, x_{t-1} (ddim.py:363-374), same fp32 op order as
// cfg_ddim_update_kernel; eps never goes to HBM.  One thread per latent pixel (4 channels).
__global__ void taps_cfg_ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ taps,
                                            const float* __restrict__ bias, const float* __restrict__ noise,
                                            float* __restrict__ x_prev, float* __restrict__ pred_x0, int B, int L, float scale,
                                            float a_t, float a_prev, float sigma, float sqrt_one_minus_at, int has_uncond) {
  pdl_wait();
  const float sqrt_at = __fsqrt_rn(a_t);
  const float sqrt_aprev = __fsqrt_rn(a_prev);
  const float dir_coef = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, a_prev), __fmul_rn(sigma, sigma)));
  const long long HW = (long long)L * L, total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % L), yy = (int)((i / L) % L), b = (int)(i / HW);
    const float4 eu4 = eps_gather(taps, bias, b, yy, xx, L);
    float4 ec4 = eu4;
    if (has_uncond) ec4 = eps_gather(taps, bias, B + b, yy, xx, L);
    const float eu[4] = {eu4.x, eu4.y, eu4.z, eu4.w}, ec[4] = {ec4.x, ec4.y, ec4.z, ec4.w};
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const long long idx = ((long long)b * 4 + ch) * HW + (long long)yy * L + xx;
      const float e = has_uncond ? __fadd_rn(eu[ch], __fmul_rn(scale, __fsub_rn(ec[ch], eu[ch]))) : eu[ch];
      const float p0 = __fdiv_rn(__fsub_rn(x[idx], __fmul_rn(sqrt_one_minus_at, e)), sqrt_at);
      const float dir = __fmul_rn(dir_coef, e);
      const float nz = noise ? __fmul_rn(sigma, noise[idx]) : 0.0f;
      x_prev[idx] = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_aprev, p0), dir), nz);
      if (pred_x0) pred_x0[idx] = p0;
    }
  }
}
// [4, C, 3, 3] fp32 -> [64 (36 used), C] fp16, row = tap * 4 + co
__global__ void pack_out_taps_kernel(const float* __restrict__ w, __half* __restrict__ dst, int C, int kp) {
  const int total = 64 * kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % kp, r = i / kp;
    const int tap = r >> 2, co = r & 3;
    dst[i] = (r < 36 && k < C) ? __float2half_rn(w[((long long)co * C + k) * 9 + tap]) : __float2half_rn(0.f);
  }
}

// classifier-free guidance combine alone (get_model_output, plms.py:184-188): e = e_u + scale * (e_c - e_u)
__global__ void cfg_combine_kernel(const float* __restrict__ eps2, float* __restrict__ out, long long count, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float eu = eps2[i], ec = eps2[count + i];
    out[i] = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(ec, eu)));
  }
}
// e_t' of p_sample_plms (plms.py:225-240), same fp32 op order, no FMA contraction.  order = len(old_eps) clipped to 3;
// o1/o2/o3 = old_eps[-1], [-2], [-3]; e_next only for order 0 (pseudo improved Euler).
__global__ void plms_combine_kernel(const float* __restrict__ e_t, const float* __restrict__ o1,
                                    const float* __restrict__ o2, const float* __restrict__ o3,
                                    const float* __restrict__ e_next, float* __restrict__ out, long long count, int order) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float e = e_t[i];
    float r;
    if (order == 0) {
      r = __fdiv_rn(__fadd_rn(e, e_next[i]), 2.0f);
    } else if (order == 1) {
      r = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, e), o1[i]), 2.0f);
    } else if (order == 2) {
      r = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, e), __fmul_rn(16.0f, o1[i])), __fmul_rn(5.0f, o2[i])), 12.0f);
    } else {
      r = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.0f, e), __fmul_rn(59.0f, o1[i])), __fmul_rn(37.0f, o2[i])),
                              __fmul_rn(9.0f, o3[i])),
                    24.0f);
    }
    out[i] = r;
  }
}
// DDPM.q_sample (ddpm.py:412-415): out = sqrt_ac[t[b]] * x0 + sqrt_1m_ac[t[b]] * noise, per sample b
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const float* __restrict__ coef, float* __restrict__ out, long long per_sample, int B) {
  const long long total = per_sample * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    out[i] = __fadd_rn(__fmul_rn(coef[2 * b], x0[i]), __fmul_rn(coef[2 * b + 1], noise[i]));
  }
}

// split-K reduction: out[m, n] = fp16(sum_z part[z][m][n] (fixed order) + bias[n] + rowvec[(m / rpv) * ldv + n] + res[m, n])
// -- the epilogue terms of the GEMM in the same order as the fused epilogue (bias, row vector, residual).  N % 4 == 0.
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int ks, long long M, int N,
                                     const float* __restrict__ bias, const float* __restrict__ rowvec, int rpv, int ldv,
                                     const __half* __restrict__ res, long long ldr, __half* __restrict__ out, long long ldo) {
  pdl_wait();
  const int nv = N >> 2;
  const long long total = M * nv;
  const long long zs = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % nv);
    const long long m = i / nv;
    const float* p = part + m * N + c4 * 4;
    float4 a = *reinterpret_cast<const float4*>(p);
    for (int z = 1; z < ks; ++z) {
      const float4 b = *reinterpret_cast<const float4*>(p + z * zs);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    const int n = c4 * 4;
    if (bias) a.x += bias[n], a.y += bias[n + 1], a.z += bias[n + 2], a.w += bias[n + 3];
    if (rowvec) {
      const float* rv = rowvec + (m / rpv) * ldv + n;
      a.x += rv[0], a.y += rv[1], a.z += rv[2], a.w += rv[3];
    }
    if (res) {
      const uint2 u = *reinterpret_cast<const uint2*>(res + m * ldr + n);
      const float2 f0 = unpack_h2(u.x), f1 = unpack_h2(u.y);
      a.x += f0.x, a.y += f0.y, a.z += f1.x, a.w += f1.y;
    }
    uint2 o;
    o.x = pack_h2(a.x, a.y), o.y = pack_h2(a.z, a.w);
    *reinterpret_cast<uint2*>(out + m * ldo + n) = o;
  }
}

// The same reduction, organised so that it can also leave the GroupNorm partial statistics the GEMM epilogues write
// (GemmArgs::stats: per 32 rows and channel the sum / sum of squares of the fp16 results): a block of 128 threads owns 32
// rows x 128 columns; thread (cg, rq) reduces columns 4 cg .. 4 cg + 3 of rows 8 rq .. 8 rq + 7, the four row quarters
// are folded through shared memory in quarter order.  M % 32 == 0, N % 128 == 0.
__global__ void __launch_bounds__(128)
splitk_reduce_stats_kernel(const float* __restrict__ part, int ks, long long M, int N, const float* __restrict__ bias,
                           const float* __restrict__ rowvec, int rpv, int ldv, const __half* __restrict__ res, long long ldr,
                           __half* __restrict__ out, long long ldo, float* __restrict__ stats) {
  pdl_wait();
  __shared__ float sh[4][32][8];
  const int cg = threadIdx.x & 31, rq = threadIdx.x >> 5;
  const int n = blockIdx.y * 128 + cg * 4;
  const long long m0 = (long long)blockIdx.x * 32 + rq * 8;
  const long long zs = M * N;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = *reinterpret_cast<const float4*>(bias + n);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int r = 0; r < 8; ++r) {
    const long long m = m0 + r;
    const float* p = part + m * N + n;
    float4 a = *reinterpret_cast<const float4*>(p);
    for (int z = 1; z < ks; ++z) {
      const float4 b = *reinterpret_cast<const float4*>(p + z * zs);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    if (bias) a.x += bv.x, a.y += bv.y, a.z += bv.z, a.w += bv.w;
    if (rowvec) {
      const float* rv = rowvec + (m / rpv) * ldv + n;
      a.x += rv[0], a.y += rv[1], a.z += rv[2], a.w += rv[3];
    }
    if (res) {
      const uint2 u = *reinterpret_cast<const uint2*>(res + m * ldr + n);
      const float2 f0 = unpack_h2(u.x), f1 = unpack_h2(u.y);
      a.x += f0.x, a.y += f0.y, a.z += f1.x, a.w += f1.y;
    }
    uint2 o;
    o.x = pack_h2(a.x, a.y), o.y = pack_h2(a.z, a.w);
    *reinterpret_cast<uint2*>(out + m * ldo + n) = o;
    if (stats) {  // statistics of the ROUNDED values, like the GEMM epilogue's
      const float2 g0 = unpack_h2(o.x), g1 = unpack_h2(o.y);
      s[0] += g0.x, q[0] = fmaf(g0.x, g0.x, q[0]);
      s[1] += g0.y, q[1] = fmaf(g0.y, g0.y, q[1]);
      s[2] += g1.x, q[2] = fmaf(g1.x, g1.x, q[2]);
      s[3] += g1.y, q[3] = fmaf(g1.y, g1.y, q[3]);
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) sh[rq][cg][2 * j] = s[j], sh[rq][cg][2 * j + 1] = q[j];
    __syncthreads();
    if (rq == 0) {
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = ((sh[0][cg][j] + sh[1][cg][j]) + sh[2][cg][j]) + sh[3][cg][j];
      float4* sp = reinterpret_cast<float4*>(stats + ((long long)blockIdx.x * N + n) * 2);
      sp[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
      sp[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
    }
  }
}

// fp32 -> fp16 copy with optional row padding: dst[r, 0..Kp) = src[r, 0..K) (zeros beyond K)
__global__ void pack_rows_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long rows, int K,
                                     int Kp) {
  const long long total = rows * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const long long r = i / Kp;
    dst[i] = (k < K) ? __float2half_rn(src[r * K + k]) : __float2half_rn(0.f);
  }
}
// conv weight [O, I, KH, KW] fp32 -> [O, taps*Ip] fp16 with k = tap*Ip + i (Ip >= I zero padded)
__global__ void pack_conv_w_kernel(const float* __restrict__ src, __half* __restrict__ dst, int O, int I, int taps,
                                   int Ip, int Kp, const float* __restrict__ oscale) {
  const long long total = (long long)O * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const int o = (int)(idx / Kp);
    const int tap = k / Ip, i = k % Ip;
    const float sc = oscale ? oscale[o] : 1.0f;
    dst[idx] = (tap < taps && i < I) ? __float2half_rn(sc * src[((long long)o * I + i) * taps + tap])
                                     : __float2half_rn(0.f);
  }
}

}  // namespace rfb
