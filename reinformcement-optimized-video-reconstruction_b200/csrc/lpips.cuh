// lpips.cuh — the HBM-bound parts of the LPIPS(net='vgg') perceptual loss (call sites
// rovr/train_local_net_unet.py:91,109-113 and rovr/rovr.py:54,255). The VGG16 convolutions themselves
// run on the igemm engine (bf16 NHWC, ReLU and 2x2 max-pool fused into the epilogue); this file holds
//   * the ScalingLayer + NCHW fp32 -> NHWC bf16 pack of BOTH images into one 2N batch,
//   * the distance head of one tap: unit-normalise both feature vectors of a pixel, weighted squared
//     difference (the 1x1 `lin` conv), spatial mean — and, in the same pass over the features, the
//     gradient of that value w.r.t. the first image's features (the target image carries none),
//   * the conversion of the gradient w.r.t. the packed input back to NCHW fp32 d/d(in0).
// Published algorithm: Zhang et al. 2018, lpips v0.1 (restated in oracle/rovr_oracle.py::lpips_vgg).
#pragma once
#include "ptx.cuh"

namespace rovr {

struct LpipsScale {
  float shift[3], inv_scale[3];
  float pre_mul, pre_add;     // normalize=True: x -> 2x - 1 first
};

// dst [2N][H][W][16] bf16: image n < N from in0, image N + n from in1; channel c < 3 =
// ((pre_mul * v + pre_add) - shift[c]) * inv_scale[c], channels 3..15 zero.
__global__ void lpips_pack_kernel(const float* __restrict__ in0, const float* __restrict__ in1, int N, long long hw,
                                  LpipsScale sc, __nv_bfloat16* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2ll * N * hw) return;
  const long long n = i / hw, p = i - n * hw;
  const float* src = (n < N ? in0 + n * 3 * hw : in1 + (n - N) * 3 * hw) + p;
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = ((sc.pre_mul * __ldg(src + c * hw) + sc.pre_add) - sc.shift[c]) * sc.inv_scale[c];
  uint4* o = reinterpret_cast<uint4*>(dst + i * 16);
  o[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], 0.f), 0u, 0u);
  o[1] = make_uint4(0u, 0u, 0u, 0u);
}

// gout [N][3][H][W] fp32 = gx16[n][p][c] * pre_mul * inv_scale[c] * gval[n]
__global__ void lpips_unpack_grad_kernel(const __nv_bfloat16* __restrict__ gx16, int N, long long hw, LpipsScale sc,
                                         const float* __restrict__ gval, float* __restrict__ gout) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * hw) return;
  const long long n = i / hw, p = i - n * hw;
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(gx16 + i * 16));
  const float g = gval[n] * sc.pre_mul;
  float* o = gout + n * 3 * hw + p;
  o[0] = bf16_lo(u.x) * sc.inv_scale[0] * g;
  o[hw] = bf16_hi(u.x) * sc.inv_scale[1] * g;
  o[2 * hw] = bf16_lo(u.y) * sc.inv_scale[2] * g;
}

// fp32 variants for the emulated-fp32 mode (LPIPS(precision="fp32x")): dst [2N][H][W][4] fp32 (3 valid channels)
__global__ void lpips_pack_f32_kernel(const float* __restrict__ in0, const float* __restrict__ in1, int N, long long hw,
                                      LpipsScale sc, float* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2ll * N * hw) return;
  const long long n = i / hw, p = i - n * hw;
  const float* src = (n < N ? in0 + n * 3 * hw : in1 + (n - N) * 3 * hw) + p;
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = ((sc.pre_mul * __ldg(src + c * hw) + sc.pre_add) - sc.shift[c]) * sc.inv_scale[c];
  reinterpret_cast<float4*>(dst)[i] = make_float4(v[0], v[1], v[2], 0.f);
}
// gx [N][H][W][ld] fp32 (gradient w.r.t. the packed input, first 3 channels) -> gout NCHW
__global__ void lpips_unpack_grad_f32_kernel(const float* __restrict__ gx, int ld, int N, long long hw, LpipsScale sc,
                                             const float* __restrict__ gval, float* __restrict__ gout) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * hw) return;
  const long long n = i / hw, p = i - n * hw;
  const float g = gval[n] * sc.pre_mul;
  float* o = gout + n * 3 * hw + p;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c * hw] = gx[i * ld + c] * sc.inv_scale[c] * g;
}

constexpr int LP_WARPS = 8;

// One tap. F: [2N][hw][C] bf16 (post-ReLU features; images n and N + n are compared), w: [C] lin weights.
// partial[n * gridDim.x + blockIdx.x] = sum over this block's pixels of sum_c w_c (a_c - b_c)^2 with
// a = f0 / (|f0| + 1e-10), b = f1 / (|f1| + 1e-10). G (optional, [N][hw][C] bf16) = inv_hw * d(that)/d f0,
// masked by f0 > 0 (the ReLU the features came through). One warp per pixel, CPL channels per lane.
// fp32 features (emulated-fp32 mode): same arithmetic, scalar loads / stores (lane l owns channels l, l + 32, ...)
template <int CPL>
__global__ void __launch_bounds__(LP_WARPS * 32)
lpips_head_f32_kernel(const float* __restrict__ F, int N, long long hw, const float* __restrict__ w, float inv_hw,
                      float* __restrict__ G, float* __restrict__ partial) {
  constexpr int C = CPL * 32;
  __shared__ float sacc[LP_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y;
  const float* f0p = F + static_cast<long long>(n) * hw * C + lane;
  const float* f1p = F + static_cast<long long>(N + n) * hw * C + lane;
  float wv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) wv[j] = __ldg(w + lane + 32 * j);
  float acc = 0.f;
  for (long long p = static_cast<long long>(blockIdx.x) * LP_WARPS + warp; p < hw; p += static_cast<long long>(gridDim.x) * LP_WARPS) {
    float a[CPL], b[CPL];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      a[j] = __ldg(f0p + p * C + 32 * j);
      b[j] = __ldg(f1p + p * C + 32 * j);
      s0 = fmaf(a[j], a[j], s0);
      s1 = fmaf(b[j], b[j], s1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float n0 = sqrtf(s0), n1 = sqrtf(s1);
    const float i0 = 1.f / (n0 + 1e-10f), i1 = 1.f / (n1 + 1e-10f);
    float v = 0.f, dot = 0.f, u[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float d = a[j] * i0 - b[j] * i1;
      v = fmaf(wv[j] * d, d, v);
      u[j] = 2.f * wv[j] * d;
      dot = fmaf(u[j], a[j], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(0xffffffffu, v, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    acc += v;
    if (G != nullptr) {
      const float coef = n0 > 0.f ? dot * i0 * i0 / n0 : 0.f;
      float* gp = G + (static_cast<long long>(n) * hw + p) * C + lane;
#pragma unroll
      for (int j = 0; j < CPL; ++j) gp[32 * j] = a[j] > 0.f ? inv_hw * (u[j] * i0 - coef * a[j]) : 0.f;
    }
  }
  if (lane == 0) sacc[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LP_WARPS; ++k) s += sacc[k];
    partial[static_cast<long long>(n) * gridDim.x + blockIdx.x] = s;
  }
}

template <int CPL>
__global__ void __launch_bounds__(LP_WARPS * 32)
lpips_head_kernel(const __nv_bfloat16* __restrict__ F, int N, long long hw, const float* __restrict__ w, float inv_hw,
                  __nv_bfloat16* __restrict__ G, float* __restrict__ partial) {
  constexpr int C = CPL * 32;
  __shared__ float sacc[LP_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y;
  const __nv_bfloat16* f0p = F + static_cast<long long>(n) * hw * C + lane * CPL;
  const __nv_bfloat16* f1p = F + static_cast<long long>(N + n) * hw * C + lane * CPL;
  float wv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) wv[j] = __ldg(w + lane * CPL + j);
  float acc = 0.f;
  for (long long p = static_cast<long long>(blockIdx.x) * LP_WARPS + warp; p < hw; p += static_cast<long long>(gridDim.x) * LP_WARPS) {
    float a[CPL], b[CPL];
    {
      uint32_t ua[CPL / 2], ub[CPL / 2];
      if constexpr (CPL == 2) {
        ua[0] = __ldg(reinterpret_cast<const uint32_t*>(f0p + p * C));
        ub[0] = __ldg(reinterpret_cast<const uint32_t*>(f1p + p * C));
      } else if constexpr (CPL == 4) {
        const uint2 x = __ldg(reinterpret_cast<const uint2*>(f0p + p * C)), y = __ldg(reinterpret_cast<const uint2*>(f1p + p * C));
        ua[0] = x.x; ua[1] = x.y; ub[0] = y.x; ub[1] = y.y;
      } else {
#pragma unroll
        for (int q = 0; q < CPL / 8; ++q) {
          const uint4 x = __ldg(reinterpret_cast<const uint4*>(f0p + p * C) + q), y = __ldg(reinterpret_cast<const uint4*>(f1p + p * C) + q);
          ua[4 * q] = x.x; ua[4 * q + 1] = x.y; ua[4 * q + 2] = x.z; ua[4 * q + 3] = x.w;
          ub[4 * q] = y.x; ub[4 * q + 1] = y.y; ub[4 * q + 2] = y.z; ub[4 * q + 3] = y.w;
        }
      }
#pragma unroll
      for (int j = 0; j < CPL / 2; ++j) {
        a[2 * j] = bf16_lo(ua[j]); a[2 * j + 1] = bf16_hi(ua[j]);
        b[2 * j] = bf16_lo(ub[j]); b[2 * j + 1] = bf16_hi(ub[j]);
      }
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) { s0 = fmaf(a[j], a[j], s0); s1 = fmaf(b[j], b[j], s1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float n0 = sqrtf(s0), n1 = sqrtf(s1);
    const float i0 = 1.f / (n0 + 1e-10f), i1 = 1.f / (n1 + 1e-10f);
    float v = 0.f, dot = 0.f, u[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float d = a[j] * i0 - b[j] * i1;
      v = fmaf(wv[j] * d, d, v);
      u[j] = 2.f * wv[j] * d;
      dot = fmaf(u[j], a[j], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(0xffffffffu, v, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    acc += v;
    if (G != nullptr) {
      // d/d f0_c [ sum_k w_k (f0_k i0 - b_k)^2 ] = u_c i0 - (sum_k u_k f0_k) i0^2 / n0 * f0_c     (d n0 / d f0_c = f0_c / n0)
      const float coef = n0 > 0.f ? dot * i0 * i0 / n0 : 0.f;
      uint32_t og[CPL / 2];
#pragma unroll
      for (int j = 0; j < CPL / 2; ++j) {
        const float g0 = a[2 * j] > 0.f ? inv_hw * (u[2 * j] * i0 - coef * a[2 * j]) : 0.f;
        const float g1 = a[2 * j + 1] > 0.f ? inv_hw * (u[2 * j + 1] * i0 - coef * a[2 * j + 1]) : 0.f;
        og[j] = pack_bf16x2(g0, g1);
      }
      __nv_bfloat16* gp = G + (static_cast<long long>(n) * hw + p) * C + lane * CPL;
      if constexpr (CPL == 2) {
        *reinterpret_cast<uint32_t*>(gp) = og[0];
      } else if constexpr (CPL == 4) {
        *reinterpret_cast<uint2*>(gp) = make_uint2(og[0], og[1]);
      } else {
#pragma unroll
        for (int q = 0; q < CPL / 8; ++q)
          reinterpret_cast<uint4*>(gp)[q] = make_uint4(og[4 * q], og[4 * q + 1], og[4 * q + 2], og[4 * q + 3]);
      }
    }
  }
  if (lane == 0) sacc[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LP_WARPS; ++k) s += sacc[k];
    partial[static_cast<long long>(n) * gridDim.x + blockIdx.x] = s;
  }
}

struct LpipsTaps {
  const float* partial[5];
  int nblocks[5];
  float inv_hw[5];
  int ntaps;
};
// val[n] = sum_taps inv_hw * sum_blocks partial   (one thread per image, fixed order)
__global__ void lpips_finalize_kernel(LpipsTaps t, int N, float* __restrict__ val) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float tot = 0.f;
  for (int k = 0; k < t.ntaps; ++k) {
    float s = 0.f;
    for (int b = 0; b < t.nblocks[k]; ++b) s += t.partial[k][static_cast<long long>(n) * t.nblocks[k] + b];
    tot += s * t.inv_hw[k];
  }
  val[n] = tot;
}

}  // namespace rovr
