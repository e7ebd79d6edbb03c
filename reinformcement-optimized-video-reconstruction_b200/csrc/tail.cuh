// tail.cuh — LocalNet tail: conv8 (1x1, 64 -> 3) + sigmoid (+ fused L2 loss) and its backward.
// Reference: rovr/local_net.py:39,71 and nn.MSELoss of rovr/train_local_net_unet.py:90,107.
//
// Both kernels are HBM-bound streams over y7 [pixels][64] bf16. Eight lanes share one pixel
// (16 bytes = 8 channels each), so a warp instruction moves 4 consecutive pixels = 512 contiguous
// bytes, and each lane keeps eight such loads in flight.
#pragma once
#include "ptx.cuh"

namespace rovr {

constexpr int TAIL_C = 64;
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_CHUNKS = 8;                                      // 32-pixel chunks per warp
constexpr int TAIL_PIX_PER_BLOCK = (TAIL_THREADS / 32) * 32 * TAIL_CHUNKS;  // 2048
constexpr int TAILB_COLS = 3 * TAIL_C + 3 + TAIL_C;                 // dW8 | db8 | db7

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}

// out[b][k][px] = sigmoid(b8[k] + sum_c w8[k][c] y7[px][c]); optional per-block sum of squared
// error against `target` (fixed-order reduction).
__global__ void __launch_bounds__(TAIL_THREADS)
tail_fwd_kernel(const __nv_bfloat16* __restrict__ y7, const float* __restrict__ w8,
                const float* __restrict__ b8, float* __restrict__ out,
                const float* __restrict__ target, float* __restrict__ loss_partial, int B, int HW) {
  __shared__ float sred[TAIL_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, cg = lane & 7;  // pixel within a group of 4, channel group of 8
  const long long npix = static_cast<long long>(B) * HW;
  float w[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 8; ++c) w[k][c] = __ldg(w8 + k * TAIL_C + cg * 8 + c);
  const float bias = cg < 3 ? __ldg(b8 + cg) : 0.f;
  float se = 0.f;
  const long long wbase = (static_cast<long long>(blockIdx.x) * (TAIL_THREADS / 32) + warp) * 32ll * TAIL_CHUNKS;
  for (int ch = 0; ch < TAIL_CHUNKS; ++ch) {
    const long long cbase = wbase + ch * 32;
    if (cbase >= npix) break;
    uint4 q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long px = cbase + 4 * j + sub;
      q[j] = px < npix ? __ldg(reinterpret_cast<const uint4*>(y7 + px * TAIL_C) + cg) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[8];
      unpack8(q[j], f);
      float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        a[0] = fmaf(f[c], w[0][c], a[0]);
        a[1] = fmaf(f[c], w[1][c], a[1]);
        a[2] = fmaf(f[c], w[2][c], a[2]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 1);
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 2);
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 4);
      }
      const long long px = cbase + 4 * j + sub;
      if (cg < 3 && px < npix) {  // lane cg of the group finishes output channel cg
        const float z = (cg == 0 ? a[0] : (cg == 1 ? a[1] : a[2])) + bias;
        const float y = 1.f / (1.f + expf(-z));
        const int b = static_cast<int>(px / HW);
        const long long o = (static_cast<long long>(b) * 3 + cg) * HW + (px - static_cast<long long>(b) * HW);
        out[o] = y;
        if (target) {
          const float d = y - __ldg(target + o);
          se = fmaf(d, d, se);
        }
      }
    }
  }
  if (loss_partial) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) sred[warp] = se;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < TAIL_THREADS / 32; ++i) v += sred[i];
      loss_partial[blockIdx.x] = v;
    }
  }
}

// Tail backward. gz_k = g_k * y_k (1 - y_k) with g_k = gout_k (if given) + mse_scale * gloss *
// (y_k - t_k) (if a fused L2 target is given; mse_scale = 2 / numel, gloss = dL/dloss on device);
//   g7[c]     = (sum_k w8[k][c] gz_k) * (y7[c] > 0)            (bf16, NHWC)
//   dW8[k][c] = sum_p gz_k y7[c];  db8[k] = sum_p gz_k;  db7[c] = sum_p g7[c]   (conv7's bias grad)
// written as per-block partial rows [grid][TAILB_COLS] (combined later in a fixed order).
__global__ void __launch_bounds__(TAIL_THREADS)
tail_bwd_kernel(const __nv_bfloat16* __restrict__ y7, const float* __restrict__ w8,
                const float* __restrict__ yout, const float* __restrict__ gout,
                const float* __restrict__ target, float mse_scale,
                const float* __restrict__ gloss, __nv_bfloat16* __restrict__ g7,
                float* __restrict__ partial, int B, int HW, int chunks) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sacc[TAILB_COLS];
  for (int i = threadIdx.x; i < TAILB_COLS; i += blockDim.x) sacc[i] = 0.f;
  const float lscale = mse_scale * (gloss ? __ldg(gloss) : 1.f);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, cg = lane & 7;
  const long long npix = static_cast<long long>(B) * HW;
  float w[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 8; ++c) w[k][c] = __ldg(w8 + k * TAIL_C + cg * 8 + c);
  float dw[3][8], dbc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { dw[0][c] = dw[1][c] = dw[2][c] = 0.f; dbc[c] = 0.f; }
  float db[3] = {0.f, 0.f, 0.f};
  __syncthreads();
  const long long wbase = (static_cast<long long>(blockIdx.x) * (TAIL_THREADS / 32) + warp) * 32ll * chunks;
  for (int ch = 0; ch < chunks; ++ch) {
    const long long cbase = wbase + ch * 32;
    if (cbase >= npix) break;
    uint4 q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long px = cbase + 4 * j + sub;
      q[j] = px < npix ? __ldg(reinterpret_cast<const uint4*>(y7 + px * TAIL_C) + cg) : make_uint4(0, 0, 0, 0);
    }
    // lane l owns pixel cbase + l for the 3-channel part (coalesced plane reads)
    float gz[3] = {0.f, 0.f, 0.f};
    {
      const long long px = cbase + lane;
      if (px < npix) {
        const int b = static_cast<int>(px / HW);
        const long long o = static_cast<long long>(b) * 3 * HW + (px - static_cast<long long>(b) * HW);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float y = __ldg(yout + o + static_cast<long long>(k) * HW);
          float g = 0.f;
          if (gout) g = __ldg(gout + o + static_cast<long long>(k) * HW);
          if (target) g += lscale * (y - __ldg(target + o + static_cast<long long>(k) * HW));
          gz[k] = g * y * (1.f - y);
          db[k] += gz[k];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int src = 4 * j + sub;  // lane that owns this pixel's gz
      const float z0 = __shfl_sync(0xffffffffu, gz[0], src);
      const float z1 = __shfl_sync(0xffffffffu, gz[1], src);
      const float z2 = __shfl_sync(0xffffffffu, gz[2], src);
      const long long px = cbase + src;
      float f[8], g[8];
      unpack8(q[j], f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        dw[0][c] = fmaf(z0, f[c], dw[0][c]);
        dw[1][c] = fmaf(z1, f[c], dw[1][c]);
        dw[2][c] = fmaf(z2, f[c], dw[2][c]);
        float v = z0 * w[0][c] + z1 * w[1][c] + z2 * w[2][c];
        if (!(f[c] > 0.f)) v = 0.f;
        g[c] = v;
        dbc[c] += v;
      }
      if (px < npix)
        reinterpret_cast<uint4*>(g7 + px * TAIL_C)[cg] =
            make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]),
                       pack_bf16x2(g[6], g[7]));
    }
  }
  // combine the 4 pixel sub-groups of the warp (lanes with equal cg), then the warps in order
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dw[k][c] += __shfl_xor_sync(0xffffffffu, dw[k][c], 8);
      dw[k][c] += __shfl_xor_sync(0xffffffffu, dw[k][c], 16);
    }
    dbc[c] += __shfl_xor_sync(0xffffffffu, dbc[c], 8);
    dbc[c] += __shfl_xor_sync(0xffffffffu, dbc[c], 16);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) db[k] += __shfl_xor_sync(0xffffffffu, db[k], o);
  for (int wv = 0; wv < TAIL_THREADS / 32; ++wv) {
    if (warp == wv && sub == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < 3; ++k) sacc[k * TAIL_C + cg * 8 + c] += dw[k][c];
        sacc[3 * TAIL_C + 3 + cg * 8 + c] += dbc[c];
      }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) sacc[3 * TAIL_C + k] += db[k];
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < TAILB_COLS; i += blockDim.x)
    partial[static_cast<long long>(blockIdx.x) * TAILB_COLS + i] = sacc[i];
}

// One launch that reduces the [grid][TAILB_COLS] partial rows of tail_bwd_kernel and routes the
// three column ranges to their gradients: dW8 (3*64) | db8 (3) | db7 (64, may be NULL).
__global__ void __launch_bounds__(1024)
tail_reduce_kernel(const float* __restrict__ partial, int nrows, float* __restrict__ dw8,
                   float* __restrict__ db8, float* __restrict__ db7) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sred[32][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cl;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (j < TAILB_COLS) {
    const float* p = partial + j;
    int r = rg;
    for (; r + 96 < nrows; r += 128) {
      a0 += __ldg(p + static_cast<long long>(r) * TAILB_COLS);
      a1 += __ldg(p + static_cast<long long>(r + 32) * TAILB_COLS);
      a2 += __ldg(p + static_cast<long long>(r + 64) * TAILB_COLS);
      a3 += __ldg(p + static_cast<long long>(r + 96) * TAILB_COLS);
    }
    for (; r < nrows; r += 32) a0 += __ldg(p + static_cast<long long>(r) * TAILB_COLS);
  }
  sred[rg][cl] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (rg == 0 && j < TAILB_COLS) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) s += sred[g][cl];
    if (j < 3 * TAIL_C) dw8[j] = s;
    else if (j < 3 * TAIL_C + 3) db8[j - 3 * TAIL_C] = s;
    else if (db7) db7[j - 3 * TAIL_C - 3] = s;
  }
}

}  // namespace rovr
