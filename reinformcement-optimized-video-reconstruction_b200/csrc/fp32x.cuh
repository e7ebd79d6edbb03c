// fp32x.cuh — support kernels for the EMULATED-FP32 policy trunks (PolicyNetwork1UNet.unet,
// PolicyNetwork2UNet.video_conv; reference rovr/policy_net_1.py:60-84, rovr/policy_net_2.py:41-60).
//
// Why: every gradient of those trunks flows through max-pools (8x8, 4x4, 2x2) whose arg-max is
// decided between values that differ by ~1e-4 relative in a few percent of the windows. With bf16
// operands (2^-9 per product) a few percent of the windows pick another element than the fp32
// reference and the gradients move by 0.2-0.6 L2-rel (scripts/exp/precision_policy.py measures it:
// fp32 storage alone does not help, the conv arithmetic itself must be fp32-class). The north_star
// tolerance is 2e-2. So these (tiny, launch-bound) trunks keep fp32 activations and run their
// contractions as SPLIT-bf16 products on the tcgen05 tensor cores:
//
//     v = hi + mid + lo     hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid)   (24 mantissa bits)
//
//     forward   x*w ~= xh*wh + xm*wh + xl*wh + xh*wm + xm*wm + xh*wl        (6 terms, error ~2^-24)
//     backward  g*w ~= gh*wh + gm*wh + gh*wm                                (3 terms, error ~2^-16)
//
// The terms are not separate launches: the pieces are STACKED along the channel (= contraction)
// dimension — x' = [xh|xm|xl|xh|xm|xh], w' = [wh|wh|wh|wm|wm|wl] — so ONE pass of the existing
// implicit-GEMM kernel over 6*C channels accumulates all of them in the fp32 TMEM accumulator
// (bf16 x bf16 products are exact in fp32). The epilogue writes fp32 (igemm direct epilogue).
// Weight gradients contract over pixels, so there the pieces are stacked along the OUTPUT dims:
// dy' = [gh|gm] (first 2 blocks of the 3-term stack), x' = [xh|xm] (first 2 blocks of the 6-term
// stack) give a [2*Cout][2*Cin] result whose four blocks are summed (blocksum kernel).
//
// This file: split + stack (with an optional fused max-pool window), weight splitting, fp32
// BatchNorm (train-mode statistics in fp64), fp32 max-pool forward / backward, fp32 flatten /
// unflatten, the wgrad block sum. All HBM-bound elementwise work, coalesced on the channel dim.
#pragma once
#include "ptx.cuh"

namespace rovr {

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// piece k of v: 0 = hi, 1 = mid, 2 = lo
__device__ __forceinline__ void split3(float v, float pc[3]) {
  pc[0] = bf16_round(v);
  const float r = v - pc[0];          // exact in fp32
  pc[1] = bf16_round(r);
  pc[2] = bf16_round(r - pc[1]);
}
// which piece goes into stacked block t. Activation side / weight side of the same product list.
//   NT = 6: act [h m l h m h]  x  wgt [h h h m m l]
//   NT = 3: act [h m h]        x  wgt [h h m]
__device__ __forceinline__ int piece_act(int nt, int t) {
  if (nt == 6) return t < 3 ? t : (t == 4 ? 1 : 0);   // h m l h m h
  if (nt == 3) return t == 1 ? 1 : 0;                 // h m h
  return t;                                           // h m (l)
}
__device__ __forceinline__ int piece_wgt(int nt, int t) {
  if (nt == 6) return t < 3 ? 0 : (t < 5 ? 1 : 2);    // h h h m m l
  if (nt == 3) return t == 2 ? 1 : 0;                 // h h m
  return t;
}

struct SplitSrc {
  const float* p;
  const float* p2;            // optional second tensor with the same strides: channels c_split .. C come from it
  int c_split;                //   (torch.cat([image, context], 1) of rovr/policy_net_1.py:88 without the copy)
  const float* relu_mask;     // optional tensor with the same strides: the value is taken as 0 where mask <= 0 (the
                              //   ReLU backward of the producing layer folded into the gradient split; no pooling then)
  long long sb, sy, sx, sc;   // element strides of (batch, row, column, channel): NHWC views and NCHW tensors alike
};

// dst[b][oy][ox][t * cb + c_off + c] = piece_act(t) of max_{window} src[b][oy*sh+dy][ox*sw+dx][c]   for c < C,
// zero for C <= c < cw. One thread per (output pixel, channel pair). kh = kw = 1: plain split.
template <int NT>
__global__ void split_stack_kernel(SplitSrc s, int B, int H, int W, int C, int kh, int kw, int sh, int sw_, int Ho,
                                   int Wo, __nv_bfloat16* __restrict__ dst, int dst_ld, int cb, int c_off, int cw) {
  const int cp2 = cw >> 1;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * Ho * Wo * cp2;
  if (i >= n) return;
  const int cp = static_cast<int>(i % cp2);
  long long r = i / cp2;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  float v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int c = 2 * cp + e;
    float m = 0.f;
    if (c < C) {
      const float* base = (c < s.c_split ? s.p + c * s.sc : s.p2 + (c - s.c_split) * s.sc) + b * s.sb +
                          static_cast<long long>(oy) * sh * s.sy + static_cast<long long>(ox) * sw_ * s.sx;
      m = __ldg(base);
      if (s.relu_mask != nullptr && !(__ldg(s.relu_mask + (base - s.p)) > 0.f)) m = 0.f;
      for (int dy = 0; dy < kh; ++dy)
        for (int dx = 0; dx < kw; ++dx) {
          if (dy == 0 && dx == 0) continue;
          m = fmaxf(m, __ldg(base + dy * s.sy + dx * s.sx));
        }
    }
    v[e] = m;
  }
  float p0[3], p1[3];
  split3(v[0], p0);
  split3(v[1], p1);
  __nv_bfloat16* o = dst + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * dst_ld + c_off + 2 * cp;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int k = piece_act(NT, t);
    *reinterpret_cast<uint32_t*>(o + t * cb) = pack_bf16x2(p0[k], p1[k]);
  }
}

// Weight side: w [d0][d1][inner] fp32 -> out with dim `stack_dim` (0 or 1) replaced by nt blocks of cb
// entries, block t holding piece_wgt(t) of the weights (exactly bf16-representable fp32 values, zero for
// the padding entries d <= j < cb); the existing repack kernels then make the bf16 GEMM operands.
__global__ void split_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int d0, int d1, int inner,
                                     int stack_dim, int nt, int cb) {
  const int o0 = stack_dim == 0 ? nt * cb : d0, o1 = stack_dim == 1 ? nt * cb : d1;
  const long long n = static_cast<long long>(o0) * o1 * inner;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = static_cast<int>(i % inner);
  long long r = i / inner;
  int i1 = static_cast<int>(r % o1);
  int i0 = static_cast<int>(r / o1);
  int blk;
  if (stack_dim == 0) { blk = i0 / cb; i0 -= blk * cb; } else { blk = i1 / cb; i1 -= blk * cb; }
  float v = 0.f;
  if (i0 < d0 && i1 < d1) {
    float pc[3];
    split3(w[(static_cast<long long>(i0) * d1 + i1) * inner + t], pc);
    v = pc[piece_wgt(nt, blk)];
  }
  out[i] = v;
}

// dw[i][j][t] = sum_{a < 2, b < 2} dwp[a * cb0 + i][b * cb1 + j][t]; dwp is [2 * cb0][2 * cb1][inner]
__global__ void blocksum4_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int d0, int d1, int inner,
                                 int cb0, int cb1) {
  const long long n = static_cast<long long>(d0) * d1 * inner;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = static_cast<int>(i % inner);
  const long long r = i / inner;
  const int j = static_cast<int>(r % d1), i0 = static_cast<int>(r / d1);
  const long long row = 2ll * cb1 * inner;
  const float* p = dwp + static_cast<long long>(i0) * row + static_cast<long long>(j) * inner + t;
  // smallest terms first: (mid, mid), then the two cross terms, then (hi, hi)
  const float s = ((p[cb0 * row + static_cast<long long>(cb1) * inner] + p[cb0 * row]) + p[static_cast<long long>(cb1) * inner]) + p[0];
  dw[i] = s;
}

// ---- BatchNorm2d on fp32 NHWC views ---------------------------------------------------------------
// statistics: [grid][2 * C] partial (sum, sum of squares); same layout as bn_stats_partial_kernel, so the
// fp64 finalize kernels of norm.cuh are shared.
__global__ void bn_stats_partial_f32_kernel(const float* __restrict__ x, int ld, long long npix, int C,
                                            float* __restrict__ partial) {
  extern __shared__ float ssum[];  // [blockDim.x * 4]
  x += static_cast<long long>(blockIdx.y) * npix * ld;        // gridDim.y > 1: one statistics group per frame
  partial += static_cast<long long>(blockIdx.y) * gridDim.x * 2 * C;
  const int c2 = C >> 1;
  const int pl = threadIdx.x / c2;
  const int cp = threadIdx.x - pl * c2;
  const int plane = blockDim.x / c2;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (pl < plane) {
    for (long long px = static_cast<long long>(blockIdx.x) * plane + pl; px < npix;
         px += static_cast<long long>(gridDim.x) * plane) {
      const float2 u = __ldg(reinterpret_cast<const float2*>(x + px * ld) + cp);
      s0 += u.x; s1 += u.y;
      q0 = fmaf(u.x, u.x, q0); q1 = fmaf(u.y, u.y, q1);
    }
  }
  float* my = ssum + 4 * threadIdx.x;
  my[0] = s0; my[1] = s1; my[2] = q0; my[3] = q1;
  __syncthreads();
  if (threadIdx.x < c2) {
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    for (int l = 0; l < plane; ++l) {
      const float* o = ssum + 4 * (l * c2 + threadIdx.x);
      a += o[0]; b += o[1]; c += o[2]; d += o[3];
    }
    float* dst = partial + static_cast<long long>(blockIdx.x) * 2 * C;
    dst[2 * threadIdx.x] = a;
    dst[2 * threadIdx.x + 1] = b;
    dst[C + 2 * threadIdx.x] = c;
    dst[C + 2 * threadIdx.x + 1] = d;
  }
}

// centred second pass: partial[block][C + c] = sum (x - mean[c])^2 (the fp32 sum-of-squares shortcut
// E[x^2] - mean^2 loses the digits the fp32-class parity needs when |mean| >> std); [0..C) = sum (x - mean).
__global__ void bn_var_partial_f32_kernel(const float* __restrict__ x, int ld, long long npix, int C,
                                          const float* __restrict__ mean, float* __restrict__ partial) {
  extern __shared__ float ssum[];
  const int c2 = C >> 1;
  const int pl = threadIdx.x / c2;
  const int cp = threadIdx.x - pl * c2;
  const int plane = blockDim.x / c2;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (pl < plane) {
    const float m0 = mean[2 * cp], m1 = mean[2 * cp + 1];
    for (long long px = static_cast<long long>(blockIdx.x) * plane + pl; px < npix;
         px += static_cast<long long>(gridDim.x) * plane) {
      const float2 u = __ldg(reinterpret_cast<const float2*>(x + px * ld) + cp);
      const float a = u.x - m0, b = u.y - m1;
      s0 += a; s1 += b;
      q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
    }
  }
  float* my = ssum + 4 * threadIdx.x;
  my[0] = s0; my[1] = s1; my[2] = q0; my[3] = q1;
  __syncthreads();
  if (threadIdx.x < c2) {
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    for (int l = 0; l < plane; ++l) {
      const float* o = ssum + 4 * (l * c2 + threadIdx.x);
      a += o[0]; b += o[1]; c += o[2]; d += o[3];
    }
    float* dst = partial + static_cast<long long>(blockIdx.x) * 2 * C;
    dst[2 * threadIdx.x] = a;
    dst[2 * threadIdx.x + 1] = b;
    dst[C + 2 * threadIdx.x] = c;
    dst[C + 2 * threadIdx.x + 1] = d;
  }
}
// stage A: mean[c] = sum / n (fp64 combine, block order)
__global__ void bn_mean_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, long long npix,
                                        float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += static_cast<double>(partial[static_cast<long long>(b) * 2 * C + c]);
  mean[c] = static_cast<float>(s / static_cast<double>(npix));
}
// stage B: from the centred sums; corrects the mean by the (tiny) residual sum, updates the running stats
__global__ void bn_var_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, long long npix, float eps,
                                       float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       long long* __restrict__ num_batches_tracked, int c_valid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    s += static_cast<double>(partial[static_cast<long long>(b) * 2 * C + c]);
    q += static_cast<double>(partial[static_cast<long long>(b) * 2 * C + C + c]);
  }
  const double n = static_cast<double>(npix);
  const double d = s / n;                       // residual of the first-pass mean
  const double m = static_cast<double>(mean[c]) + d;
  double var = q / n - d * d;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (c < c_valid && running_mean != nullptr) {
    const double unbiased = npix > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// y = [relu](gamma * (x - mean) * rstd + beta) in the operation order of ATen's batch_norm
// ((x - mean) * invstd * weight + bias); channels >= c_valid are written as 0.
__global__ void bn_apply_f32_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld,
                                    long long npix, int C, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, int c_valid, int relu) {
  const int c4 = C >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix * c4) return;
  const int cb = static_cast<int>(i % c4) * 4;
  const long long px = i / c4;
  const float4 u = __ldg(reinterpret_cast<const float4*>(x + px * x_ld + cb));
  float v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cb + e;
    float r = 0.f;
    if (c < c_valid) {
      r = (v[e] - mean[c]) * rstd[c] * gamma[c] + beta[c];
      if (relu) r = fmaxf(r, 0.f);
    }
    v[e] = r;
  }
  *reinterpret_cast<float4*>(y + px * y_ld + cb) = make_float4(v[0], v[1], v[2], v[3]);
}

// backward partials: g = dy * (y > 0) [if relu]; [0..C) = sum g, [C..2C) = sum g * xhat
__global__ void bn_bwd_partial_f32_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ y,
                                          int y_ld, const float* __restrict__ x, int x_ld, long long npix, int C,
                                          const float* __restrict__ mean, const float* __restrict__ rstd, int relu,
                                          float* __restrict__ partial) {
  extern __shared__ float ssum[];
  const int c2 = C >> 1;
  const int pl = threadIdx.x / c2;
  const int cp = threadIdx.x - pl * c2;
  const int plane = blockDim.x / c2;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (pl < plane) {
    const float m0 = mean[2 * cp], m1 = mean[2 * cp + 1], r0 = rstd[2 * cp], r1 = rstd[2 * cp + 1];
    for (long long px = static_cast<long long>(blockIdx.x) * plane + pl; px < npix;
         px += static_cast<long long>(gridDim.x) * plane) {
      float2 g = __ldg(reinterpret_cast<const float2*>(dy + px * dy_ld) + cp);
      const float2 xv = __ldg(reinterpret_cast<const float2*>(x + px * x_ld) + cp);
      if (relu) {
        const float2 yv = __ldg(reinterpret_cast<const float2*>(y + px * y_ld) + cp);
        if (!(yv.x > 0.f)) g.x = 0.f;
        if (!(yv.y > 0.f)) g.y = 0.f;
      }
      s0 += g.x; s1 += g.y;
      q0 = fmaf(g.x, (xv.x - m0) * r0, q0);
      q1 = fmaf(g.y, (xv.y - m1) * r1, q1);
    }
  }
  float* my = ssum + 4 * threadIdx.x;
  my[0] = s0; my[1] = s1; my[2] = q0; my[3] = q1;
  __syncthreads();
  if (threadIdx.x < c2) {
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    for (int l = 0; l < plane; ++l) {
      const float* o = ssum + 4 * (l * c2 + threadIdx.x);
      a += o[0]; b += o[1]; c += o[2]; d += o[3];
    }
    float* dst = partial + static_cast<long long>(blockIdx.x) * 2 * C;
    dst[2 * threadIdx.x] = a;
    dst[2 * threadIdx.x + 1] = b;
    dst[C + 2 * threadIdx.x] = c;
    dst[C + 2 * threadIdx.x + 1] = d;
  }
}

// dx = gamma * rstd * (g - sum_g / N - xhat * sum_gxhat / N)   (eval mode: gamma * rstd * g)
__global__ void bn_bwd_apply_f32_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ y, int y_ld,
                                        const float* __restrict__ x, int x_ld, float* __restrict__ dx, int dx_ld,
                                        long long npix, int C, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                        const float* __restrict__ sums, int c_valid, int relu, int eval_mode) {
  const int c4 = C >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix * c4) return;
  const int cb = static_cast<int>(i % c4) * 4;
  const long long px = i / c4;
  const float inv_n = 1.f / static_cast<float>(npix);
  const float4 ug = __ldg(reinterpret_cast<const float4*>(dy + px * dy_ld + cb));
  const float4 ux = __ldg(reinterpret_cast<const float4*>(x + px * x_ld + cb));
  float4 uy = make_float4(1.f, 1.f, 1.f, 1.f);
  if (relu) uy = __ldg(reinterpret_cast<const float4*>(y + px * y_ld + cb));
  const float g[4] = {ug.x, ug.y, ug.z, ug.w}, xv[4] = {ux.x, ux.y, ux.z, ux.w}, yv[4] = {uy.x, uy.y, uy.z, uy.w};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cb + e;
    float r = 0.f;
    if (c < c_valid) {
      const float gg = (yv[e] > 0.f) ? g[e] : 0.f;
      const float xh = (xv[e] - mean[c]) * rstd[c];
      r = eval_mode ? gamma[c] * rstd[c] * gg : gamma[c] * rstd[c] * (gg - sums[c] * inv_n - xh * sums[C + c] * inv_n);
    }
    o[e] = r;
  }
  *reinterpret_cast<float4*>(dx + px * dx_ld + cb) = make_float4(o[0], o[1], o[2], o[3]);
}

// ---- max-pool on fp32 NHWC views ---------------------------------------------------------------------
__global__ void maxpool_f32_fwd_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld, int B,
                                       int H, int W, int C, int kh, int kw, int sh, int sw_, int Ho, int Wo) {
  const int c4 = C >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * Ho * Wo * c4;
  if (i >= n) return;
  const int cv = static_cast<int>(i % c4);
  long long r = i / c4;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const float* xb = x + ((static_cast<long long>(b) * H + oy * sh) * W + ox * sw_) * x_ld + cv * 4;
  float4 m = __ldg(reinterpret_cast<const float4*>(xb));
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      if (dy == 0 && dx == 0) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (static_cast<long long>(dy) * W + dx) * x_ld));
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  *reinterpret_cast<float4*>(y + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * y_ld + cv * 4) = m;
}

// non-overlapping windows (kernel == stride, H % kh == W % kw == 0): one thread per (window, 4 channels)
// scans the window once (first maximum in row-major order wins, like ATen) and writes all kh*kw outputs:
// gx = gskip + (position == arg-max ? gp : 0).
__global__ void maxpool_f32_bwd_tiled_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ gp,
                                             int gp_ld, const float* __restrict__ gskip, int gs_ld,
                                             float* __restrict__ gx, int gx_ld, int B, int H, int W, int C, int kh,
                                             int kw) {
  const int c4 = C >> 2;
  const int Ho = H / kh, Wo = W / kw;
  const long long n = static_cast<long long>(B) * Ho * Wo * c4;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cv = static_cast<int>(i % c4);
  long long r = i / c4;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const long long pix0 = (static_cast<long long>(b) * H + oy * kh) * W + ox * kw;
  const float4 gv = __ldg(reinterpret_cast<const float4*>(gp + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * gp_ld + cv * 4));
  const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (pix0 + static_cast<long long>(dy) * W + dx) * x_ld + cv * 4));
      const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (f[j] > best[j]) { best[j] = f[j]; arg[j] = dy * kw + dx; }
    }
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      const long long pix = pix0 + static_cast<long long>(dy) * W + dx;
      float4 sk = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gskip) sk = __ldg(reinterpret_cast<const float4*>(gskip + pix * gs_ld + cv * 4));
      float o[4] = {sk.x, sk.y, sk.z, sk.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (arg[j] == dy * kw + dx) o[j] += g4[j];
      *reinterpret_cast<float4*>(gx + pix * gx_ld + cv * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// any window geometry (overlapping windows of PolicyNetwork2UNet's MaxPool2d(2, stride=(2, 1))): one thread
// per input element gathers from the windows that contain it.
__global__ void maxpool_f32_bwd_generic_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ gp,
                                               int gp_ld, float* __restrict__ gx, int gx_ld, int B, int H, int W,
                                               int C, int kh, int kw, int sh, int sw_, int Ho, int Wo) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * H * W * C;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  long long r = i / C;
  const int xx = static_cast<int>(r % W);
  r /= W;
  const int yy = static_cast<int>(r % H);
  const int b = static_cast<int>(r / H);
  // windows oy with oy*sh <= yy < oy*sh + kh
  const int oy_lo = yy - kh + 1 <= 0 ? 0 : (yy - kh + sh) / sh, oy_hi = min(Ho - 1, yy / sh);
  const int ox_lo = xx - kw + 1 <= 0 ? 0 : (xx - kw + sw_) / sw_, ox_hi = min(Wo - 1, xx / sw_);
  float g = 0.f;
  for (int oy = oy_lo; oy <= oy_hi; ++oy)
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      float best = -INFINITY;
      int ay = 0, ax = 0;
      for (int dy = 0; dy < kh; ++dy)
        for (int dx = 0; dx < kw; ++dx) {
          const float f = x[((static_cast<long long>(b) * H + oy * sh + dy) * W + ox * sw_ + dx) * x_ld + c];
          if (f > best) { best = f; ay = oy * sh + dy; ax = ox * sw_ + dx; }
        }
      if (ay == yy && ax == xx) g += gp[((static_cast<long long>(b) * Ho + oy) * Wo + ox) * gp_ld + c];
    }
  gx[((static_cast<long long>(b) * H + yy) * W + xx) * gx_ld + c] = g;
}

// ---- NCHW-flatten of an fp32 NHWC view (nn.Flatten on NCHW) and its inverse --------------------------
// rows[b][c * HW + p] = x[b][p][c], c < C
__global__ void flatten_f32_kernel(const float* __restrict__ x, int ld, float* __restrict__ rows, long long rows_ld,
                                   int B, int HW, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * HW * C;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  long long r = i / C;
  const int p = static_cast<int>(r % HW);
  const int b = static_cast<int>(r / HW);
  rows[b * rows_ld + static_cast<long long>(c) * HW + p] = x[(static_cast<long long>(b) * HW + p) * ld + c];
}
// x[b][p][c] = rows[b][c * HW + p] for c < C, 0 for C <= c < cpad
__global__ void unflatten_f32_kernel(const float* __restrict__ rows, long long rows_ld, float* __restrict__ x, int ld,
                                     int B, int HW, int C, int cpad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * HW * cpad;
  if (i >= n) return;
  const int c = static_cast<int>(i % cpad);
  long long r = i / cpad;
  const int p = static_cast<int>(r % HW);
  const int b = static_cast<int>(r / HW);
  x[(static_cast<long long>(b) * HW + p) * ld + c] = c < C ? rows[b * rows_ld + static_cast<long long>(c) * HW + p] : 0.f;
}

}  // namespace rovr
