// attention.cuh — the memory-bound pieces around the tensor-core GEMMs of the attention blocks
// (rovr/common_layers.py:54-118): per-head operand transposes, the row softmax over key tokens and
// its gradient, exact GELU and its gradient, fp32 -> bf16 casts. QK^T, PV, their gradients and all
// projections run on igemm_kernel (rovr_gemm_bf16 / rovr_gemm_batched_bf16 / rovr_gemm_wgrad).
#pragma once
#include "ptx.cuh"

namespace rovr {

// out[i2][i1][c][r] = in[i2][i1][r][c] for r < R, c < C; out columns r in [R, r_pad) are zero.
// 64x64 tiles through shared memory, bf16 PAIRS on both sides (a warp reads / writes 128 contiguous
// bytes per row); grid = (ceil(r_pad/64), ceil(C/64), n1*n2), block = (32, 8). C and the strides are
// even (all callers: multiples of 8).
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long in_ld, long long in_s1,
                                      long long in_s2, __nv_bfloat16* __restrict__ out, long long out_ld,
                                      long long out_s1, long long out_s2, int R, int C, int r_pad, int n1) {
  __shared__ uint32_t tile[64][33];   // [row r][pair of columns]
  const int i1 = blockIdx.z % n1, i2 = blockIdx.z / n1;
  const __nv_bfloat16* src = in + i2 * in_s2 + i1 * in_s1;
  __nv_bfloat16* dst = out + i2 * out_s2 + i1 * out_s1;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  for (int j = threadIdx.y; j < 64; j += 8) {
    const int r = r0 + j, c = c0 + 2 * threadIdx.x;
    uint32_t v = 0u;
    if (r < R && c < C) v = __ldg(reinterpret_cast<const uint32_t*>(src + r * in_ld + c));
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  // output row = input column c0 + j; this thread writes the pair of input rows (r0 + 2 tx, r0 + 2 tx + 1)
  for (int j = threadIdx.y; j < 64; j += 8) {
    const int c = c0 + j, r = r0 + 2 * threadIdx.x;
    if (c < C && r < r_pad) {
      const uint32_t a = tile[2 * threadIdx.x][j >> 1], bb = tile[2 * threadIdx.x + 1][j >> 1];
      const uint32_t lo = (j & 1) ? (a >> 16) : (a & 0xFFFFu);
      const uint32_t hi = (j & 1) ? (bb >> 16) : (bb & 0xFFFFu);
      *reinterpret_cast<uint32_t*>(dst + c * out_ld + r) = lo | (hi << 16);
    }
  }
}

// P[row][t] = softmax_t(scale * s[row][t]) for t < T, 0 for T <= t < t_pad. One warp per row.
__global__ void softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, long long rows,
                                   int T, int t_pad, float scale) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + row * t_pad;
  float mx = -INFINITY;
  for (int t = lane; t < T; t += 32) mx = fmaxf(mx, sr[t] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int t = lane; t < T; t += 32) sum += expf(sr[t] * scale - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int t = lane; t < t_pad; t += 32)
    p[row * t_pad + t] = __float2bfloat16_rn(t < T ? expf(sr[t] * scale - mx) * inv : 0.f);
}
// dS = scale * P * (dP - sum_t dP * P)
__global__ void softmax_bwd_kernel(const float* __restrict__ dp, const __nv_bfloat16* __restrict__ p,
                                   __nv_bfloat16* __restrict__ ds, long long rows, int T, int t_pad, float scale) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int t = lane; t < T; t += 32) dot += dp[row * t_pad + t] * __bfloat162float(p[row * t_pad + t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int t = lane; t < t_pad; t += 32) {
    const float v = t < T ? scale * __bfloat162float(p[row * t_pad + t]) * (dp[row * t_pad + t] - dot) : 0.f;
    ds[row * t_pad + t] = __float2bfloat16_rn(v);
  }
}

// exact GELU (F.gelu default, rovr/common_layers.py:91): a = h * Phi(h)
__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ a, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __bfloat162float(h[i]);
  a[i] = __float2bfloat16_rn(0.5f * x * (1.f + erff(x * 0.70710678118654752f)));
}
// dh = da * (Phi(h) + h * phi(h))
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ h,
                                __nv_bfloat16* __restrict__ dh, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __bfloat162float(h[i]);
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  dh[i] = __float2bfloat16_rn(__bfloat162float(da[i]) * (cdf + x * pdf));
}

// out = a + b on fp32 vectors (residual joins of the encoder / decoder blocks), 16 bytes per thread
__global__ void add_f32_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                               long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = __ldg(a + i), y = __ldg(b + i);
  out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// out[p][j] = x[b][p][j] + (w1[j] * p1 + b1[j]) + (w2 ? w2[j] * p2 + b2[j] : 0), p1 = p % n1, p2 = p / n1:
// the learned positional tables of rovr/common_layers.py:7-52 are Linear(1, D) applied to arange(n),
// i.e. rank-1 in the position index, so they are evaluated on the fly instead of materialised.
__global__ void posenc_add_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ w1,
                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                  const float* __restrict__ b2, long long total, int P, int D, int n1) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = static_cast<int>(i % D);
  const int p = static_cast<int>((i / D) % P);
  float v = x[i] + w1[j] * static_cast<float>(p % n1) + b1[j];
  if (w2) v += w2[j] * static_cast<float>(p / n1) + b2[j];
  out[i] = v;
}
// gradients of the two Linear(1, D): gw[j] = sum_{b,p} g * pos(p), gb[j] = sum_{b,p} g. One block per
// 32 columns, 8 row groups, fixed-order combine.
__global__ void posenc_grad_kernel(const float* __restrict__ g, long long rows, int P, int D, int n1, int which,
                                   float* __restrict__ gw, float* __restrict__ gb) {
  __shared__ float sa[8][33], sb[8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cl;
  float a = 0.f, b = 0.f;
  if (j < D) {
    for (long long r = rg; r < rows; r += 8) {
      const int p = static_cast<int>(r % P);
      const float pos = static_cast<float>(which == 0 ? p % n1 : p / n1);
      const float gg = g[r * D + j];
      a += gg * pos;
      b += gg;
    }
  }
  sa[rg][cl] = a;
  sb[rg][cl] = b;
  __syncthreads();
  if (rg == 0 && j < D) {
    float s = 0.f, t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += sa[k][cl]; t += sb[k][cl]; }
    gw[j] = s;
    gb[j] = t;
  }
}

// ---- dropout (nn.Dropout / the attention-probability dropout of nn.MultiheadAttention) -------------------
// rovr/common_layers.py:58,70 (MultiheadAttention(dropout=p)) and :87,91 (nn.Dropout after GELU).
// The keep mask is a pure function of (seed, step counter, call site, element index): nothing is stored,
// the backward pass recomputes it. `state` = device int64 [2] = {seed, counter} (a snapshot taken by the
// forward; the live counter is advanced by dropout_advance_kernel, which a CUDA graph captures too, so
// every replay draws a new mask). out = x * keep / (1 - p).
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long ctr, int site, unsigned long long i,
                                             uint32_t threshold) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr * 64ull + static_cast<unsigned long long>(site) + 1ull);
  z ^= i * 0xD1342543DE82EF95ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;     // splitmix64 finaliser
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z >> 32) >= threshold;
}
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ out, long long n, uint32_t threshold, float inv_keep,
                               const long long* __restrict__ state, int site) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long seed = static_cast<unsigned long long>(state[0]), ctr = static_cast<unsigned long long>(state[1]);
  const float v = dropout_keep(seed, ctr, site, static_cast<unsigned long long>(i), threshold) ? static_cast<float>(x[i]) * inv_keep : 0.f;
  out[i] = static_cast<T>(v);
}
__global__ void dropout_advance_kernel(long long* state) { state[1] += 1; }

}  // namespace rovr
