// policy.cuh — fp32 kernels of the policy heads, the frame-feature projection and ActionLSTM.
//
// These layers have a tiny row count (M = batch <= a few dozen), so they are weight-read-bound
// GEMVs, not tensor-core work: one warp streams one weight row with 16-byte loads and keeps the
// activations of up to 8 batch rows in registers. Everything is fp32 so that the frame indices the
// policies select are a deterministic function of the logits (SURVEY.md §8c "RNG for bit-exact
// indices").
//
// Reference semantics:
//   nn.Linear chains            rovr/policy_net_1.py:54-57,94; rovr/policy_net_2.py:63-69,79;
//                               rovr/resnet_extractor.py:9,46; rovr/action_lstm.py:14,35
//   standardisation quirks      rovr/policy_net_1.py:91-93,100; rovr/policy_net_2.py:104-106,121-122
//   gumbel-softmax / selection  rovr/policy_net_1.py:101-103,113-114; rovr/policy_net_2.py:98-102,138-141
//   LSTMCell pointwise          rovr/action_lstm.py:33
#pragma once
#include "ptx.cuh"

namespace rovr {

constexpr int LIN_MT = 8;  // batch rows per pass

// y[m][n] = bias[n] + sum_k x[m][k] * w[n][k]   (+ optional second operand: + sum_k x2[m][k] * w2[n][k])
// one warp per output column n; grid.y walks the batch in chunks of LIN_MT rows.
__global__ void __launch_bounds__(256)
linear_f32_fwd_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ y, int y_ld, int M, int N,
                      int K, int accumulate) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int m0 = blockIdx.y * LIN_MT;
  const int mc = min(LIN_MT, M - m0);
  float acc[LIN_MT];
#pragma unroll
  for (int i = 0; i < LIN_MT; ++i) acc[i] = 0.f;
  const float* wr = w + static_cast<long long>(n) * K;
  if ((K & 3) == 0 && (x_ld & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int i = 0; i < LIN_MT; ++i) {
        if (i < mc) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(m0 + i) * x_ld + k));
          acc[i] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int i = 0; i < LIN_MT; ++i)
        if (i < mc) acc[i] += wv * __ldg(x + static_cast<long long>(m0 + i) * x_ld + k);
    }
  }
#pragma unroll
  for (int i = 0; i < LIN_MT; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (lane == 0) {
    const float b = bias ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < LIN_MT; ++i)
      if (i < mc) {
        float* d = y + static_cast<long long>(m0 + i) * y_ld + n;
        *d = accumulate ? (*d + acc[i] + b) : (acc[i] + b);
      }
  }
}

// dx[m][k] = sum_n dy[m][n] * w[n][k]; thread per k (coalesced weight reads along k), n split over
// gridDim.z slices whose partial sums land in ws[slice][M][K] (combined by reduce_rows_kernel).
__global__ void __launch_bounds__(128)
linear_f32_dgrad_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ w,
                        float* __restrict__ part, int M, int N, int K, int n_per_slice) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m0 = blockIdx.y * LIN_MT;
  const int mc = min(LIN_MT, M - m0);
  const int n0 = blockIdx.z * n_per_slice, n1 = min(N, n0 + n_per_slice);
  if (k >= K) return;
  float acc[LIN_MT];
#pragma unroll
  for (int i = 0; i < LIN_MT; ++i) acc[i] = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float wv = __ldg(w + static_cast<long long>(n) * K + k);
#pragma unroll
    for (int i = 0; i < LIN_MT; ++i)
      if (i < mc) acc[i] += wv * __ldg(dy + static_cast<long long>(m0 + i) * dy_ld + n);
  }
  float* dst = part + (static_cast<long long>(blockIdx.z) * M + m0) * K + k;
#pragma unroll
  for (int i = 0; i < LIN_MT; ++i)
    if (i < mc) dst[static_cast<long long>(i) * K] = acc[i];
}

// dw[n][k] = sum_m dy[m][n] * x[m][k]; db[n] = sum_m dy[m][n]. Thread per (n, k) element.
__global__ void __launch_bounds__(256)
linear_f32_wgrad_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ x, int x_ld,
                        float* __restrict__ dw, float* __restrict__ db, int M, int N, int K) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<long long>(n) * K);
  float s = 0.f, b = 0.f;
  for (int m = 0; m < M; ++m) {
    const float g = __ldg(dy + static_cast<long long>(m) * dy_ld + n);
    s += g * __ldg(x + static_cast<long long>(m) * x_ld + k);
    b += g;
  }
  dw[i] = s;
  if (k == 0 && db) db[n] = b;
}

// ---- standardisation along one dimension ---------------------------------------------------------
// y[o][i] = (x[o][i] - mean_o) / (std_o + eps_add), std unbiased (n-1). Element (o, i) lives at
// o*so + i*si. One warp per outer index. sig[o] = std_o is saved for the backward pass.
__global__ void standardize_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                       float* __restrict__ sig, int outer, int len, long long so,
                                       long long si, float eps_add) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= outer) return;
  const float* xo = x + o * so;
  float s = 0.f;
  for (int i = lane; i < len; i += 32) s += xo[i * si];
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  const float m = s / len;
  float q = 0.f;
  for (int i = lane; i < len; i += 32) {
    const float d = xo[i * si] - m;
    q += d * d;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) q += __shfl_xor_sync(0xffffffffu, q, k);
  const float sd = sqrtf(q / (len - 1));
  if (lane == 0) sig[o] = sd;
  const float inv = 1.f / (sd + eps_add);
  for (int i = lane; i < len; i += 32) y[o * so + i * si] = (xo[i * si] - m) * inv;
}
// dx_i = (g_i - mean(g)) / s - (sum_j g_j y_j) * y_i / ((n-1) * sigma),   s = sigma + eps_add
__global__ void standardize_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                       const float* __restrict__ sig, float* __restrict__ dx,
                                       int outer, int len, long long so, long long si, float eps_add) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= outer) return;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < len; i += 32) {
    const float gg = g[o * so + i * si];
    a += gg;
    b += gg * y[o * so + i * si];
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, k);
    b += __shfl_xor_sync(0xffffffffu, b, k);
  }
  const float sd = sig[o];
  const float inv = 1.f / (sd + eps_add);
  const float c = b / ((len - 1) * sd);
  a /= len;
  for (int i = lane; i < len; i += 32)
    dx[o * so + i * si] = (g[o * so + i * si] - a) * inv - c * y[o * so + i * si];
}

// ---- policy heads: logits are [b][n] with n <= 32 and b <= 1024, one block, thread per row -------
constexpr int HEAD_MAX_N = 32;

// (a) scatter 0 at the target columns (in place: the reference uses scatter_ on the Linear output,
//     rovr/policy_net_2.py:121,138);
// (b) if standardize: out[i][j] = (l[i][j] - mean_b(i, j)) / (std_i + 0.1) where, reproducing
//     `logits - logits.mean(dim=1)` WITHOUT keepdim (rovr/policy_net_1.py:100, policy_net_2.py:122),
//     the mean vector [b] broadcasts along the LAST dim: mean_b(i, j) = mean of row j when b == n,
//     the single row mean when b == 1 (any other shape is a broadcast error in the reference too).
__global__ void head_mask_std_fwd_kernel(float* __restrict__ logits, const long long* __restrict__ target,
                                         int tk, float* __restrict__ out, float* __restrict__ sig,
                                         int b, int n, int standardize) {
  __shared__ float smean[1024];
  const int i = threadIdx.x;
  float l[HEAD_MAX_N];
  if (i < b) {
    for (int t = 0; t < tk; ++t) {
      const long long c = target[static_cast<long long>(i) * tk + t];
      if (c >= 0 && c < n) logits[static_cast<long long>(i) * n + c] = 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j) {
      l[j] = j < n ? logits[static_cast<long long>(i) * n + j] : 0.f;
      s += l[j];
    }
    smean[i] = s / n;
  }
  __syncthreads();
  if (i >= b || !standardize) return;
  const float mu = smean[i];
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j)
    if (j < n) q += (l[j] - mu) * (l[j] - mu);
  const float sd = sqrtf(q / (n - 1));
  sig[i] = sd;
  const float inv = 1.f / (sd + 0.1f);
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j)
    if (j < n) out[static_cast<long long>(i) * n + j] = (l[j] - (b == 1 ? mu : smean[j])) * inv;
}

// backward of (b) followed by (a): g = dL/dout -> dlogits (zero at the scattered columns).
//   out_ij = (l_ij - m_j) / s_i,  m_j = mean(row j),  s_i = sigma_i + 0.1
//   dl_ab = g_ab / s_a                                         direct
//         - (1/n) * sum_i g_ia / s_i                           through m_a (column a of g / s)   [b == n]
//         - (sum_j g_aj * out_aj) / s_a * (l_ab - mu_a) / ((n-1) sigma_a)     through sigma_a
//   (for b == 1 the second term is -(1/n) sum_j g_0j / s_0).
__global__ void head_mask_std_bwd_kernel(const float* __restrict__ g, const float* __restrict__ logits,
                                         const float* __restrict__ out, const float* __restrict__ sig,
                                         const long long* __restrict__ target, int tk,
                                         float* __restrict__ dl, int b, int n, int standardize) {
  __shared__ float scol[HEAD_MAX_N];
  const int i = threadIdx.x;
  if (threadIdx.x < HEAD_MAX_N) scol[threadIdx.x] = 0.f;
  __syncthreads();
  if (standardize && b == n) {
    // column sums of g / s: thread j < n walks the rows (b == n <= 32)
    if (i < n) {
      float s = 0.f;
      for (int r = 0; r < b; ++r) s += g[static_cast<long long>(r) * n + i] / (sig[r] + 0.1f);
      scol[i] = s;
    }
  }
  __syncthreads();
  if (i >= b) return;
  float d[HEAD_MAX_N];
  if (!standardize) {
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j) d[j] = j < n ? g[static_cast<long long>(i) * n + j] : 0.f;
  } else {
    const float sd = sig[i], s = sd + 0.1f;
    float mu = 0.f, dot = 0.f, rowsum = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j)
      if (j < n) {
        mu += logits[static_cast<long long>(i) * n + j];
        dot += g[static_cast<long long>(i) * n + j] * out[static_cast<long long>(i) * n + j];
        rowsum += g[static_cast<long long>(i) * n + j];
      }
    mu /= n;
    const float c = dot / (s * (n - 1) * sd);
    const float through_mean = (b == 1) ? rowsum / (s * n) : scol[i] / n;
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j)
      if (j < n)
        d[j] = g[static_cast<long long>(i) * n + j] / s - through_mean -
               c * (logits[static_cast<long long>(i) * n + j] - mu);
  }
  for (int t = 0; t < tk; ++t) {
    const long long c = target[static_cast<long long>(i) * tk + t];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j)
      if (j == c) d[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j)
    if (j < n) dl[static_cast<long long>(i) * n + j] = d[j];
}

// F.gumbel_softmax(hard=False, dim=1) with the Exp(1) draw supplied by the caller (torch's RNG):
// probs = softmax((logits - log(expo)) / tau). Then, per mode:
//   mode 0: nothing else (probs only)
//   mode 1: idx[i] = argmax, val[i] = log p_max                       (rovr/policy_net_1.py:102-103)
//   mode 2: idx[i][0..1] = top-2 (descending), val[i] = (log p1 + log p2)/2 + 0.69314   (policy_net_2.py:100-102)
//   mode 3: val[i] = log p[action[i]]                                   (policy_net_1.py:114)
//   mode 4: val[i] = log(p[a0] * p[a1]) / 2 + 0.69314                   (policy_net_2.py:139-141)
__global__ void head_gumbel_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ expo,
                                       float tau, float* __restrict__ probs, int b, int n, int mode,
                                       const long long* __restrict__ action, long long* __restrict__ idx,
                                       float* __restrict__ val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float z[HEAD_MAX_N];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j) {
    z[j] = -INFINITY;
    if (j < n) {
      z[j] = (logits[static_cast<long long>(i) * n + j] - logf(expo[static_cast<long long>(i) * n + j])) / tau;
      mx = fmaxf(mx, z[j]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j)
    if (j < n) {
      z[j] = expf(z[j] - mx);
      s += z[j];
    }
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j)
    if (j < n) {
      z[j] = z[j] / s;
      probs[static_cast<long long>(i) * n + j] = z[j];
    }
  if (mode == 1 || mode == 2) {
    int i1 = 0;
    float p1 = -1.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j)
      if (j < n && z[j] > p1) { p1 = z[j]; i1 = j; }
    if (mode == 1) {
      idx[i] = i1;
      val[i] = logf(p1);
    } else {
      int i2 = 0;
      float p2 = -1.f;
#pragma unroll
      for (int j = 0; j < HEAD_MAX_N; ++j)
        if (j < n && j != i1 && z[j] > p2) { p2 = z[j]; i2 = j; }
      idx[2 * i] = i1;
      idx[2 * i + 1] = i2;
      val[i] = (logf(p1) + logf(p2)) / 2.f + 0.69314f;
    }
  } else if (mode == 3) {
    const long long a = action[i];
    float p = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j)
      if (j == a) p = z[j];
    val[i] = logf(p);
  } else if (mode == 4) {
    const long long a0 = action[2 * i], a1 = action[2 * i + 1];
    float p0 = 0.f, p1 = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j) {
      if (j == a0) p0 = z[j];
      if (j == a1) p1 = z[j];
    }
    val[i] = logf(p0 * p1) / 2.f + 0.69314f;
  }
}

// backward of modes 3 / 4 through the softmax: dlogits = p * (gp - sum_j gp_j p_j) / tau, where
// gp = dL/dprobs: mode 3: gval / p[a] at a; mode 4: gval / (2 p[a]) at a0 and at a1 (summed if equal).
__global__ void head_gumbel_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ gval,
                                       float tau, int b, int n, int mode,
                                       const long long* __restrict__ action, float* __restrict__ dl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float p[HEAD_MAX_N], gp[HEAD_MAX_N];
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j) {
    p[j] = j < n ? probs[static_cast<long long>(i) * n + j] : 0.f;
    gp[j] = 0.f;
  }
  const float gv = gval[i];
  if (mode == 3) {
    const long long a = action[i];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j)
      if (j == a) gp[j] = gv / p[j];
  } else {
    const long long a0 = action[2 * i], a1 = action[2 * i + 1];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_N; ++j) {
      if (j == a0) gp[j] += gv / (2.f * p[j]);
      if (j == a1) gp[j] += gv / (2.f * p[j]);
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j) dot += gp[j] * p[j];
#pragma unroll
  for (int j = 0; j < HEAD_MAX_N; ++j)
    if (j < n) dl[static_cast<long long>(i) * n + j] = p[j] * (gp[j] - dot) / tau;
}

// ---- LSTMCell pointwise (gates = x W_ih^T + b_ih + h W_hh^T + b_hh already summed) ---------------
// gate order i, f, g, o (torch.nn.LSTMCell); act saves the activated gates for the backward pass.
__global__ void lstm_pointwise_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                          float* __restrict__ h, float* __restrict__ c,
                                          float* __restrict__ act, int B, int Hd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * Hd) return;
  const int b = t / Hd, j = t - b * Hd;
  const float* gr = gates + static_cast<long long>(b) * 4 * Hd;
  const float ig = 1.f / (1.f + expf(-gr[j]));
  const float fg = 1.f / (1.f + expf(-gr[Hd + j]));
  const float gg = tanhf(gr[2 * Hd + j]);
  const float og = 1.f / (1.f + expf(-gr[3 * Hd + j]));
  const float cn = fg * c_prev[t] + ig * gg;
  c[t] = cn;
  h[t] = og * tanhf(cn);
  if (act) {
    float* ar = act + static_cast<long long>(b) * 4 * Hd;
    ar[j] = ig; ar[Hd + j] = fg; ar[2 * Hd + j] = gg; ar[3 * Hd + j] = og;
  }
}
// dgates (pre-activation) and dc_prev from dh, dc.
__global__ void lstm_pointwise_bwd_kernel(const float* __restrict__ act, const float* __restrict__ c_prev,
                                          const float* __restrict__ c, const float* __restrict__ dh,
                                          const float* __restrict__ dc_in, float* __restrict__ dgates,
                                          float* __restrict__ dc_prev, int B, int Hd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * Hd) return;
  const int b = t / Hd, j = t - b * Hd;
  const float* ar = act + static_cast<long long>(b) * 4 * Hd;
  const float ig = ar[j], fg = ar[Hd + j], gg = ar[2 * Hd + j], og = ar[3 * Hd + j];
  const float tc = tanhf(c[t]);
  const float dhv = dh ? dh[t] : 0.f;
  const float dcv = (dc_in ? dc_in[t] : 0.f) + dhv * og * (1.f - tc * tc);
  float* dg = dgates + static_cast<long long>(b) * 4 * Hd;
  dg[j] = dcv * gg * ig * (1.f - ig);
  dg[Hd + j] = dcv * c_prev[t] * fg * (1.f - fg);
  dg[2 * Hd + j] = dcv * ig * (1.f - gg * gg);
  dg[3 * Hd + j] = dhv * tc * og * (1.f - og);
  if (dc_prev) dc_prev[t] = dcv * fg;
}

// ---- small data movers -----------------------------------------------------------------------
// NHWC bf16 [B][H][W][ld] (first C channels) -> NCHW-flattened fp32 rows [B][C*H*W] at column offset
// (the Flatten of rovr/policy_net_2.py:59 and `rearrange b c h w -> b (c h w)` of policy_net_1.py:90)
__global__ void flatten_nhwc_to_rows_kernel(const __nv_bfloat16* __restrict__ src, int ld, float* __restrict__ dst,
                                            int dst_ld, int B, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW * C) return;
  const int c = i % C, p = (i / C) % HW, b = i / (C * HW);
  dst[static_cast<long long>(b) * dst_ld + c * HW + p] = __bfloat162float(src[(static_cast<long long>(b) * HW + p) * ld + c]);
}
__global__ void unflatten_rows_to_nhwc_kernel(const float* __restrict__ src, int src_ld, __nv_bfloat16* __restrict__ dst,
                                              int ld, int B, int HW, int C, int cpad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW * cpad) return;
  const int c = i % cpad, p = (i / cpad) % HW, b = i / (cpad * HW);
  const float v = c < C ? src[static_cast<long long>(b) * src_ld + c * HW + p] : 0.f;
  dst[(static_cast<long long>(b) * HW + p) * ld + c] = __float2bfloat16_rn(v);
}
// fp32 strided 2-D copy: dst[r][c] = src[r][c] (+ optional scale), used for concat / slice of small
// feature rows (torch.cat([vector_out, image_out], 1), rovr/policy_net_2.py:92).
__global__ void copy2d_f32_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int dst_ld,
                                  int rows, int cols, float scale, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  const float v = src[static_cast<long long>(r) * src_ld + c] * scale;
  float* d = dst + static_cast<long long>(r) * dst_ld + c;
  *d = accumulate ? *d + v : v;
}

}  // namespace rovr
