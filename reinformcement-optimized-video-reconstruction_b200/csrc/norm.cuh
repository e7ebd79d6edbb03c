// norm.cuh — train-mode BatchNorm2d (+ReLU) for the policy networks and LayerNorm for the
// attention blocks. All kernels are HBM-bound: NHWC bf16 activations are read with 16-byte
// vector loads (8 channels per thread), statistics are fp32 partials combined in a fixed order
// in fp64, so results are bitwise reproducible run to run.
//
// Reference semantics:
//   nn.BatchNorm2d in training mode — rovr/policy_net_1.py:20-49,61-81, rovr/policy_net_2.py:43-55:
//     y = gamma * (x - mean_batch) / sqrt(var_biased + eps) + beta, running stats updated with
//     momentum 0.1 and the UNBIASED variance, num_batches_tracked += 1;
//   nn.LayerNorm — rovr/common_layers.py:59,71-72,86.
#pragma once
#include "ptx.cuh"

namespace rovr {

// ---- BatchNorm statistics: stage 1, [grid][2*C] partial (sum, sum of squares) ------------------
// blockDim = 256 (or C/2 if larger): thread t owns channel pair (t % (C/2)) of pixel lane t / (C/2).
// gridDim.y > 1: one statistics group per FRAME of npix pixels (partial is [frame][gridDim.x][2 * C]).
__global__ void bn_stats_partial_kernel(const __nv_bfloat16* __restrict__ x, int ld, long long npix,
                                        int C, float* __restrict__ partial) {
  extern __shared__ float ssum[];  // [blockDim.x * 4]
  x += static_cast<long long>(blockIdx.y) * npix * ld;
  partial += static_cast<long long>(blockIdx.y) * gridDim.x * 2 * C;
  const int c2 = C >> 1;
  const int pl = threadIdx.x / c2;
  const int cp = threadIdx.x - pl * c2;
  const int plane = blockDim.x / c2;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (pl < plane) {
    for (long long px = static_cast<long long>(blockIdx.x) * plane + pl; px < npix;
         px += static_cast<long long>(gridDim.x) * plane) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(x + px * ld) + cp);
      const float a = bf16_lo(u), b = bf16_hi(u);
      s0 += a; s1 += b;
      q0 += a * a; q1 += b * b;
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

// stage 2: one thread per channel; fp64 combine in block order.
__global__ void bn_stats_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                         long long npix, float eps, float momentum,
                                         float* __restrict__ mean, float* __restrict__ rstd,
                                         float* __restrict__ running_mean,
                                         float* __restrict__ running_var,
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
  const double m = s / n;
  double var = q / n - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (c < c_valid && running_mean != nullptr) {
    const double unbiased = npix > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// Per-frame statistics: the reference encodes every frame on its own (rovr/resnet_extractor.py:42-47,
// `.unsqueeze(0)` -> a batch of ONE), so a train-mode trunk (the default `pretrained=False` constructor,
// rovr/resnet_extractor.py:6-8) normalises each frame with its own mean / variance and updates the running
// buffers once per frame, in frame order. One thread per channel walks the frames in that order.
// mean / rstd: [frames][C].
__global__ void bn_frames_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, long long pix,
                                          int frames, float eps, float momentum, float* __restrict__ mean,
                                          float* __restrict__ rstd, float* __restrict__ running_mean,
                                          float* __restrict__ running_var,
                                          long long* __restrict__ num_batches_tracked, int c_valid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += frames;
  if (c >= C) return;
  const bool track = c < c_valid && running_mean != nullptr;
  float rm = track ? running_mean[c] : 0.f, rv = track ? running_var[c] : 0.f;
  const double n = static_cast<double>(pix);
  for (int f = 0; f < frames; ++f) {
    const float* pf = partial + static_cast<long long>(f) * nblocks * 2 * C;
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblocks; ++b) {
      s += static_cast<double>(pf[static_cast<long long>(b) * 2 * C + c]);
      q += static_cast<double>(pf[static_cast<long long>(b) * 2 * C + C + c]);
    }
    const double m = s / n;
    double var = q / n - m * m;
    if (var < 0.0) var = 0.0;
    mean[static_cast<long long>(f) * C + c] = static_cast<float>(m);
    rstd[static_cast<long long>(f) * C + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const double unbiased = pix > 1 ? var * n / (n - 1.0) : var;
    rm = (1.f - momentum) * rm + momentum * static_cast<float>(m);
    rv = (1.f - momentum) * rv + momentum * static_cast<float>(unbiased);
  }
  if (track) {
    running_mean[c] = rm;
    running_var[c] = rv;
  }
}

// eval mode: rstd[c] = 1 / sqrt(running_var[c] + eps) (mean is running_mean itself)
__global__ void bn_eval_rstd_kernel(const float* __restrict__ running_var, float eps, float* __restrict__ rstd,
                                    int c_valid, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) rstd[c] = c < c_valid ? rsqrtf(running_var[c] + eps) : 0.f;
}

// y = [relu](gamma * (x - mean) * rstd + beta); channels >= c_valid (zero padding) are written as 0.
// frame_pix > 0: mean / rstd are [frames][C], pixel px belongs to frame px / frame_pix.
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ x, int x_ld,
                                __nv_bfloat16* __restrict__ y, int y_ld, long long npix, int C,
                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                int c_valid, int relu, long long frame_pix) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix * c8) return;
  const int cb = static_cast<int>(i % c8) * 8;
  const long long px = i / c8;
  if (frame_pix > 0) {
    mean += (px / frame_pix) * C;
    rstd += (px / frame_pix) * C;
  }
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + px * x_ld + cb));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[2] = {bf16_lo(w[j]), bf16_hi(w[j])};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = cb + 2 * j + e;
      float r = 0.f;
      if (c < c_valid) {
        r = (v[e] - mean[c]) * rstd[c] * gamma[c] + beta[c];
        if (relu) r = fmaxf(r, 0.f);
      }
      v[e] = r;
    }
    o[j] = pack_bf16x2(v[0], v[1]);
  }
  *reinterpret_cast<uint4*>(y + px * y_ld + cb) = make_uint4(o[0], o[1], o[2], o[3]);
}

// the same from an fp32 convolution output (x_ld in floats, a multiple of 4) to the bf16 activation
__global__ void bn_apply_f32in_kernel(const float* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y, int y_ld,
                                      long long npix, int C, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int c_valid, int relu, long long frame_pix) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix * c8) return;
  const int cb = static_cast<int>(i % c8) * 8;
  const long long px = i / c8;
  if (frame_pix > 0) {
    mean += (px / frame_pix) * C;
    rstd += (px / frame_pix) * C;
  }
  const float4 a = __ldg(reinterpret_cast<const float4*>(x + px * x_ld + cb));
  const float4 b = __ldg(reinterpret_cast<const float4*>(x + px * x_ld + cb + 4));
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cb + e;
    float r = 0.f;
    if (c < c_valid) {
      r = (v[e] - mean[c]) * rstd[c] * gamma[c] + beta[c];
      if (relu) r = fmaxf(r, 0.f);
    }
    v[e] = r;
  }
  *reinterpret_cast<uint4*>(y + px * y_ld + cb) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ---- BatchNorm backward ---------------------------------------------------------------------
// g = dy * (y > 0) [if relu]; partial[block][0..C) = sum g, [C..2C) = sum g * xhat.
__global__ void bn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld,
                                      const __nv_bfloat16* __restrict__ y, int y_ld,
                                      const __nv_bfloat16* __restrict__ x, int x_ld, long long npix,
                                      int C, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, int relu,
                                      float* __restrict__ partial) {
  extern __shared__ float ssum[];  // [blockDim.x * 4]
  const int c2 = C >> 1;
  const int pl = threadIdx.x / c2;
  const int cp = threadIdx.x - pl * c2;
  const int plane = blockDim.x / c2;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (pl < plane) {
    const float m0 = mean[2 * cp], m1 = mean[2 * cp + 1], r0 = rstd[2 * cp], r1 = rstd[2 * cp + 1];
    for (long long px = static_cast<long long>(blockIdx.x) * plane + pl; px < npix;
         px += static_cast<long long>(gridDim.x) * plane) {
      const uint32_t ug = __ldg(reinterpret_cast<const uint32_t*>(dy + px * dy_ld) + cp);
      const uint32_t ux = __ldg(reinterpret_cast<const uint32_t*>(x + px * x_ld) + cp);
      float g0 = bf16_lo(ug), g1 = bf16_hi(ug);
      if (relu) {
        const uint32_t uy = __ldg(reinterpret_cast<const uint32_t*>(y + px * y_ld) + cp);
        if (!(bf16_lo(uy) > 0.f)) g0 = 0.f;
        if (!(bf16_hi(uy) > 0.f)) g1 = 0.f;
      }
      s0 += g0; s1 += g1;
      q0 += g0 * (bf16_lo(ux) - m0) * r0;
      q1 += g1 * (bf16_hi(ux) - m1) * r1;
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

// sums[c] = sum g, sums[C + c] = sum g*xhat (fp32, stays in the workspace for the apply kernel);
// dgamma / dbeta (length c_valid) are the parameter gradients.
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                       float* __restrict__ sums, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int c_valid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    s += static_cast<double>(partial[static_cast<long long>(b) * 2 * C + c]);
    q += static_cast<double>(partial[static_cast<long long>(b) * 2 * C + C + c]);
  }
  sums[c] = static_cast<float>(s);
  sums[C + c] = static_cast<float>(q);
  if (c < c_valid) {
    if (dbeta) dbeta[c] = static_cast<float>(s);
    if (dgamma) dgamma[c] = static_cast<float>(q);
  }
}

// dx = gamma * rstd * (g - sum_g / N - xhat * sum_gxhat / N)
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld,
                                    const __nv_bfloat16* __restrict__ y, int y_ld,
                                    const __nv_bfloat16* __restrict__ x, int x_ld,
                                    __nv_bfloat16* __restrict__ dx, int dx_ld, long long npix, int C,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sums,
                                    int c_valid, int relu, int eval_mode) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix * c8) return;
  const int cb = static_cast<int>(i % c8) * 8;
  const long long px = i / c8;
  const float inv_n = 1.f / static_cast<float>(npix);
  const uint4 ug = __ldg(reinterpret_cast<const uint4*>(dy + px * dy_ld + cb));
  const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + px * x_ld + cb));
  uint4 uy = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  if (relu) uy = __ldg(reinterpret_cast<const uint4*>(y + px * y_ld + cb));
  const uint32_t wg[4] = {ug.x, ug.y, ug.z, ug.w}, wx[4] = {ux.x, ux.y, ux.z, ux.w},
                 wy[4] = {uy.x, uy.y, uy.z, uy.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float g[2] = {bf16_lo(wg[j]), bf16_hi(wg[j])};
    const float xv[2] = {bf16_lo(wx[j]), bf16_hi(wx[j])};
    const float yv[2] = {bf16_lo(wy[j]), bf16_hi(wy[j])};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = cb + 2 * j + e;
      float r = 0.f;
      if (c < c_valid) {
        const float gg = (yv[e] > 0.f) ? g[e] : 0.f;
        const float xh = (xv[e] - mean[c]) * rstd[c];
        // eval mode: mean / rstd are constants (running statistics), so only the direct term remains
        r = eval_mode ? gamma[c] * rstd[c] * gg
                      : gamma[c] * rstd[c] * (gg - sums[c] * inv_n - xh * sums[C + c] * inv_n);
      }
      g[e] = r;
    }
    o[j] = pack_bf16x2(g[0], g[1]);
  }
  *reinterpret_cast<uint4*>(dx + px * dx_ld + cb) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- LayerNorm over the last dim (E elements, contiguous rows) -----------------------------------
// One warp per row. x fp32 in; y bf16 (GEMM operand) and/or fp32 out; saves mean / rstd.
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long rows, int E, float eps,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ y_f32, __nv_bfloat16* __restrict__ y_bf16,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * E;
  float s = 0.f;
  for (int i = lane; i < E; i += 32) s += xr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float m = s / E;
  float q = 0.f;
  for (int i = lane; i < E; i += 32) {
    const float d = xr[i] - m;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float r = rsqrtf(q / E + eps);
  if (lane == 0) {
    mean[row] = m;
    rstd[row] = r;
  }
  for (int i = lane; i < E; i += 32) {
    const float v = (xr[i] - m) * r * gamma[i] + beta[i];
    if (y_f32) y_f32[row * E + i] = v;
    if (y_bf16) y_bf16[row * E + i] = __float2bfloat16_rn(v);
  }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)); per-row partials of dgamma /
// dbeta are accumulated by a second kernel (column sums over rows).
__global__ void layernorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                     long long rows, int E, const float* __restrict__ gamma,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ dx, int accumulate) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float m = mean[row], r = rstd[row];
  const float* xr = x + row * E;
  const float* gr = g + row * E;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < E; i += 32) {
    const float gg = gr[i] * gamma[i];
    a += gg;
    b += gg * (xr[i] - m) * r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  a /= E;
  b /= E;
  for (int i = lane; i < E; i += 32) {
    const float xh = (xr[i] - m) * r;
    const float v = r * (gr[i] * gamma[i] - a - xh * b);
    float* d = dx + row * E + i;
    *d = accumulate ? *d + v : v;
  }
}

// dgamma[i] = sum_rows g*xhat, dbeta[i] = sum_rows g: block = 32 columns x 8 row groups,
// partial over gridDim.y row chunks -> [chunks][2][E], reduced by reduce_rows_kernel.
__global__ void layernorm_param_grad_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                            long long rows, int E, const float* __restrict__ mean,
                                            const float* __restrict__ rstd, int rows_per_chunk,
                                            float* __restrict__ partial) {
  __shared__ float sa[8][33], sb[8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cl;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float a = 0.f, b = 0.f;
  if (j < E) {
    for (long long r = r0 + rg; r < r1; r += 8) {
      const float gg = g[r * E + j];
      a += gg * (x[r * E + j] - mean[r]) * rstd[r];
      b += gg;
    }
  }
  sa[rg][cl] = a;
  sb[rg][cl] = b;
  __syncthreads();
  if (rg == 0 && j < E) {
    float s = 0.f, t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s += sa[k][cl];
      t += sb[k][cl];
    }
    partial[(static_cast<long long>(blockIdx.y) * 2) * E + j] = s;
    partial[(static_cast<long long>(blockIdx.y) * 2 + 1) * E + j] = t;
  }
}

}  // namespace rovr
