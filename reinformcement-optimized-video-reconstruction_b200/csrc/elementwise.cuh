// elementwise.cuh — the HBM-bound kernels around the tensor-core engine: layout packing, max
// pooling (forward / backward fused with the skip-gradient add and ReLU mask), the 1x1-conv +
// sigmoid + L2-loss tail of LocalNet and its backward, per-channel column sums (bias gradients)
// and weight repacking. All are coalesced 16-byte-vector kernels; reductions are two-stage and
// deterministic (no float atomics).
#pragma once
#include "ptx.cuh"

namespace rovr {

// ---------------------------------------------------------------------------------------------
// NCHW fp32 (up to 3 source tensors, channel-concatenated) -> NHWC bf16 padded to cpad channels.
// Reference: the input pack of LocalNet, rovr/local_net.py:48-49 (cat + 'b n c h w -> b (n c) h w'),
// and torch.cat([x, context], 1) of rovr/policy_net_1.py:88.
// ---------------------------------------------------------------------------------------------
struct PackSrc {
  const float* ptr[3];
  int ch[3];
  int nsrc;
};
__global__ void pack_nchw_to_nhwc_kernel(PackSrc src, __nv_bfloat16* __restrict__ dst, int B,
                                         int HW, int cpad) {
  pdl_trigger();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * HW) return;
  const int b = static_cast<int>(i / HW);
  const int px = static_cast<int>(i - static_cast<long long>(b) * HW);
  // gather the pixel's channels (plane reads are coalesced across the warp), then write the
  // NHWC row with 16-byte stores
  uint4* o = reinterpret_cast<uint4*>(dst + i * cpad);
  int s = 0, k = 0;
  for (int c0 = 0; c0 < cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      while (s < src.nsrc && k >= src.ch[s]) { ++s; k = 0; }
      if (s < src.nsrc) {
        v[j] = __ldg(src.ptr[s] + (static_cast<long long>(b) * src.ch[s] + k) * HW + px);
        ++k;
      } else {
        v[j] = 0.f;
      }
    }
    o[c0 >> 3] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                            pack_bf16x2(v[6], v[7]));
  }
}

// torchvision ToTensor on the device: uint8 -> fp32 / denom (255; a true IEEE division, so that the result is
// bit-identical to the host's `.div(255)`), 16 values per thread. Video frames are
// decoded to uint8 (rovr/video_ds.py:107-114: cv2.imread / resize, then the ToTensor transform on the host);
// shipping the uint8 frames and converting here cuts the host -> device bytes of a step 4x.
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n, float denom) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + i));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float4* o = reinterpret_cast<float4*>(dst + i);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_float4(__fdiv_rn(static_cast<float>(w[j] & 0xffu), denom), __fdiv_rn(static_cast<float>((w[j] >> 8) & 0xffu), denom),
                         __fdiv_rn(static_cast<float>((w[j] >> 16) & 0xffu), denom), __fdiv_rn(static_cast<float>(w[j] >> 24), denom));
  } else {
    for (long long k = i; k < n; ++k) dst[k] = __fdiv_rn(static_cast<float>(src[k]), denom);
  }
}

// The dataset's mask corruption on the device (rovr/video_ds.py:62-87, difficulty < 2 branch): frame n of a clip
// gets a zeroed box_w x box_h box at x0 = (n % 8) * W / 8, y0 = (n / 8) * H / 3, clipped to the frame;
// out = clean * mask, mask (optional) = 1 outside the box. frame_index[i] is the clip-frame index n of image i.
__global__ void corrupt_frames_kernel(const float* __restrict__ clean, const long long* __restrict__ frame_index,
                                      float* __restrict__ out, float* __restrict__ mask, int N, int C, int H, int W,
                                      int box_w, int box_h) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(N) * C * H * W;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  const int y = static_cast<int>((i / W) % H);
  const int n = static_cast<int>(i / (static_cast<long long>(C) * H * W));
  const long long f = frame_index[n];
  const int x0 = static_cast<int>(f % 8) * W / 8, y0 = static_cast<int>(f / 8) * H / 3;
  const bool inside = x >= x0 && x < min(W, x0 + box_w) && y >= y0 && y < min(H, y0 + box_h);
  out[i] = inside ? 0.f : clean[i];
  if (mask != nullptr) mask[i] = inside ? 0.f : 1.f;
}

// NHWC bf16 (ld) -> NCHW fp32, used to hand results back at the module boundary.
__global__ void unpack_nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int ld,
                                           float* __restrict__ dst, int B, int HW, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * HW) return;
  const int b = static_cast<int>(i / HW);
  const int px = static_cast<int>(i - static_cast<long long>(b) * HW);
  for (int c = 0; c < C; ++c)
    dst[(static_cast<long long>(b) * C + c) * HW + px] = __bfloat162float(src[i * ld + c]);
}

// ---------------------------------------------------------------------------------------------
// Max pooling, NHWC bf16, 8 channels (16 B) per thread. Reference: nn.MaxPool2d,
// rovr/local_net.py:21,53-55; rovr/policy_net_1.py:29; rovr/policy_net_2.py:45-58.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void max8(uint4& a, const uint4& b) {
  __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int j = 0; j < 4; ++j) pa[j] = __hmax2(pa[j], pb[j]);
}
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld,
                                   __nv_bfloat16* __restrict__ y, int y_ld, int B, int H, int W,
                                   int C, int kh, int kw, int sh, int sw_, int Ho, int Wo) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * Ho * Wo * c8;
  if (i >= n) return;
  const int cv = static_cast<int>(i % c8);
  long long r = i / c8;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const __nv_bfloat16* xb = x + ((static_cast<long long>(b) * H + oy * sh) * W + ox * sw_) * x_ld + cv * 8;
  uint4 m = __ldg(reinterpret_cast<const uint4*>(xb));
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      if (dy == 0 && dx == 0) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(xb + (static_cast<long long>(dy) * W + dx) * x_ld));
      max8(m, v);
    }
  *reinterpret_cast<uint4*>(y + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * y_ld + cv * 8) = m;
}

// 2x2 / stride 2 specialisation: the four 16-byte loads of a window are issued together and two
// windows are processed per thread, so eight requests are in flight per thread.
__global__ void __launch_bounds__(256)
maxpool_fwd_2x2_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y, int y_ld,
                       int B, int H, int W, int C) {
  const int c8 = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long n = static_cast<long long>(B) * Ho * Wo * c8;
  const long long i0 = (static_cast<long long>(blockIdx.x) * blockDim.x * 2) + threadIdx.x;
  uint4 v[2][4];
  long long oaddr[2];
  bool ok[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long i = i0 + static_cast<long long>(u) * blockDim.x;
    ok[u] = i < n;
    const long long ii = ok[u] ? i : 0;
    const int cv = static_cast<int>(ii % c8);
    long long r = ii / c8;
    const int ox = static_cast<int>(r % Wo);
    r /= Wo;
    const int oy = static_cast<int>(r % Ho);
    const int b = static_cast<int>(r / Ho);
    const __nv_bfloat16* xb = x + ((static_cast<long long>(b) * H + oy * 2) * W + ox * 2) * x_ld + cv * 8;
    v[u][0] = __ldg(reinterpret_cast<const uint4*>(xb));
    v[u][1] = __ldg(reinterpret_cast<const uint4*>(xb + x_ld));
    v[u][2] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(W) * x_ld));
    v[u][3] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(W + 1) * x_ld));
    oaddr[u] = ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * y_ld + cv * 8;
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    max8(v[u][0], v[u][1]);
    max8(v[u][2], v[u][3]);
    max8(v[u][0], v[u][2]);
    if (ok[u]) *reinterpret_cast<uint4*>(y + oaddr[u]) = v[u][0];
  }
}

// Backward for non-overlapping windows (stride == kernel, H % kh == 0, W % kw == 0): one thread
// owns one window x 8 channels, routes the pooled gradient to the first maximum in row-major scan
// order (ATen's rule), adds the skip-connection gradient arriving at the same tensor, and applies
// the ReLU mask of the pooled tensor (x > 0). This is the fused form of
//   g_x = relu'(x) * (g_skip + maxpool_backward(g_pool))
// that autograd evaluates for x1/x2/x3 of rovr/local_net.py:52-68.
__global__ void maxpool_bwd_tiled_kernel(const __nv_bfloat16* __restrict__ x, int x_ld,
                                         const __nv_bfloat16* __restrict__ gp, int gp_ld,
                                         const __nv_bfloat16* __restrict__ gskip, int gs_ld,
                                         __nv_bfloat16* __restrict__ gx, int gx_ld, int B, int H,
                                         int W, int C, int kh, int kw, int relu_mask,
                                         float* __restrict__ colsum_partial) {
  // colsum_partial (optional, [gridDim.x][C]): per-block column sums of gx — the bias gradient of
  // the convolution that produced x — so that no separate pass over gx is needed. Requires
  // blockDim.x % (C/8) == 0 (every thread of a block keeps the same channel group).
  __shared__ float scs[8][32][9];
  const int c8 = C >> 3;
  const int Ho = H / kh, Wo = W / kw;
  const long long n = static_cast<long long>(B) * Ho * Wo * c8;
  float csum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) csum[j] = 0.f;
  // grid-stride: the stride is a multiple of blockDim.x, hence of C/8, so a thread keeps its channel group
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
  const int cv = static_cast<int>(i % c8);
  long long r = i / c8;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const long long pix0 = (static_cast<long long>(b) * H + oy * kh) * W + ox * kw;
  const uint4 gv = __ldg(reinterpret_cast<const uint4*>(
      gp + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * gp_ld + cv * 8));
  const __nv_bfloat16* g8 = reinterpret_cast<const __nv_bfloat16*>(&gv);
  float best[8];
  int arg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (pix0 + static_cast<long long>(dy) * W + dx) * x_ld + cv * 8));
      const __nv_bfloat16* v8 = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __bfloat162float(v8[j]);
        if (f > best[j]) { best[j] = f; arg[j] = dy * kw + dx; }
      }
    }
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      const long long pix = pix0 + static_cast<long long>(dy) * W + dx;
      uint4 sk = make_uint4(0, 0, 0, 0);
      if (gskip) sk = __ldg(reinterpret_cast<const uint4*>(gskip + pix * gs_ld + cv * 8));
      const __nv_bfloat16* s8 = reinterpret_cast<const __nv_bfloat16*>(&sk);
      // re-read this element's own value (L1 hit) for the ReLU mask
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + pix * x_ld + cv * 8));
      const __nv_bfloat16* x8 = reinterpret_cast<const __nv_bfloat16*>(&xv);
      uint4 o;
      __nv_bfloat16* o8 = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float g = __bfloat162float(s8[j]);
        if (arg[j] == dy * kw + dx) g += __bfloat162float(g8[j]);
        if (relu_mask && !(__bfloat162float(x8[j]) > 0.f)) g = 0.f;
        o8[j] = __float2bfloat16_rn(g);
        csum[j] += g;
      }
      *reinterpret_cast<uint4*>(gx + pix * gx_ld + cv * 8) = o;
    }
  }
  if (colsum_partial == nullptr) return;
  // lanes l, l + c8, l + 2*c8, ... of a warp hold the same channel group: fold them with shuffles,
  // then the eight warps through shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o >= c8; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], o);
  }
  if (lane < c8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) scs[warp][lane][j] = csum[j];
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < C) {
    const int g8 = threadIdx.x >> 3, e = threadIdx.x & 7;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += scs[w][g8][e];
    colsum_partial[static_cast<long long>(blockIdx.x) * C + threadIdx.x] = s;
  }
}

// 2x2 / stride 2 specialisation of the kernel above (every LocalNet / PolicyNetwork1 pool): all nine
// 16-byte loads of a window (gp, 4 x, 4 gskip) are issued before anything is consumed, so each
// thread keeps nine requests in flight instead of one or two.
__global__ void __launch_bounds__(256)
maxpool_bwd_2x2_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, const __nv_bfloat16* __restrict__ gp,
                       int gp_ld, const __nv_bfloat16* __restrict__ gskip, int gs_ld,
                       __nv_bfloat16* __restrict__ gx, int gx_ld, int B, int H, int W, int C, int relu_mask,
                       float* __restrict__ colsum_partial) {
  pdl_trigger();
  pdl_wait();
  __shared__ float scs[8][32][9];
  const int c8 = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long n = static_cast<long long>(B) * Ho * Wo * c8;
  float csum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) csum[j] = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % c8);
    long long r = i / c8;
    const int ox = static_cast<int>(r % Wo);
    r /= Wo;
    const int oy = static_cast<int>(r % Ho);
    const int b = static_cast<int>(r / Ho);
    const long long pix0 = (static_cast<long long>(b) * H + oy * 2) * W + ox * 2;
    const long long pix[4] = {pix0, pix0 + 1, pix0 + W, pix0 + W + 1};
    uint4 xv[4], sk[4];
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gp + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * gp_ld + cv * 8));
#pragma unroll
    for (int q = 0; q < 4; ++q) xv[q] = __ldg(reinterpret_cast<const uint4*>(x + pix[q] * x_ld + cv * 8));
#pragma unroll
    for (int q = 0; q < 4; ++q)
      sk[q] = gskip ? __ldg(reinterpret_cast<const uint4*>(gskip + pix[q] * gs_ld + cv * 8)) : make_uint4(0, 0, 0, 0);
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t o[4][4];
#pragma unroll
    for (int wq = 0; wq < 4; ++wq) {            // one bf16 pair at a time
      float xs[4][2], ss[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t xw = (&xv[q].x)[wq], sw_ = (&sk[q].x)[wq];
        xs[q][0] = bf16_lo(xw); xs[q][1] = bf16_hi(xw);
        ss[q][0] = bf16_lo(sw_); ss[q][1] = bf16_hi(sw_);
      }
      const float gg[2] = {bf16_lo(gw[wq]), bf16_hi(gw[wq])};
      float res[4][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int arg = 0;                              // first maximum in window scan order (ATen's rule)
        float best = xs[0][e];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (xs[q][e] > best) { best = xs[q][e]; arg = q; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float g = ss[q][e] + (arg == q ? gg[e] : 0.f);
          if (relu_mask && !(xs[q][e] > 0.f)) g = 0.f;
          res[q][e] = g;
          csum[2 * wq + e] += g;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][wq] = pack_bf16x2(res[q][0], res[q][1]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(gx + pix[q] * gx_ld + cv * 8) = make_uint4(o[q][0], o[q][1], o[q][2], o[q][3]);
  }
  if (colsum_partial == nullptr) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o2 = 16; o2 >= c8; o2 >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], o2);
  }
  if (lane < c8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) scs[warp][lane][j] = csum[j];
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < C) {
    const int g8 = threadIdx.x >> 3, e = threadIdx.x & 7;
    float s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s2 += scs[w][g8][e];
    colsum_partial[static_cast<long long>(blockIdx.x) * C + threadIdx.x] = s2;
  }
}

// Generic (possibly overlapping / ragged) backward: one thread per input element vector, gathers
// from every window that contains it. Only used on the tiny 5x5 maps of PolicyNetwork2UNet.
__global__ void maxpool_bwd_generic_kernel(const __nv_bfloat16* __restrict__ x, int x_ld,
                                           const __nv_bfloat16* __restrict__ gp, int gp_ld,
                                           __nv_bfloat16* __restrict__ gx, int gx_ld, int B, int H,
                                           int W, int C, int kh, int kw, int sh, int sw_, int Ho,
                                           int Wo, int relu_mask) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * H * W * C;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  long long r = i / C;
  const int xx = static_cast<int>(r % W);
  r /= W;
  const int yy = static_cast<int>(r % H);
  const int b = static_cast<int>(r / H);
  float g = 0.f;
  for (int oy = 0; oy < Ho; ++oy) {
    if (yy < oy * sh || yy >= oy * sh + kh) continue;
    for (int ox = 0; ox < Wo; ++ox) {
      if (xx < ox * sw_ || xx >= ox * sw_ + kw) continue;
      float best = -INFINITY;
      int ay = 0, ax = 0;
      for (int dy = 0; dy < kh; ++dy)
        for (int dx = 0; dx < kw; ++dx) {
          const float f = __bfloat162float(
              x[((static_cast<long long>(b) * H + oy * sh + dy) * W + ox * sw_ + dx) * x_ld + c]);
          if (f > best) { best = f; ay = oy * sh + dy; ax = ox * sw_ + dx; }
        }
      if (ay == yy && ax == xx)
        g += __bfloat162float(gp[((static_cast<long long>(b) * Ho + oy) * Wo + ox) * gp_ld + c]);
    }
  }
  const long long pix = (static_cast<long long>(b) * H + yy) * W + xx;
  if (relu_mask && !(__bfloat162float(x[pix * x_ld + c]) > 0.f)) g = 0.f;
  gx[pix * gx_ld + c] = __float2bfloat16_rn(g);
}

// sum a vector of per-block partials in a fixed order -> out[0] = scale * sum
__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float scale,
                                    float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sred[32];
  // eight independent accumulators: the loads of a thread are in flight together (one dependent chain
  // of 48 L2 round trips took 24 us for the 49152 partials of a B=24 step); fixed order throughout
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int bd = blockDim.x;
  int i = threadIdx.x;
  for (; i + 7 * bd < n; i += 8 * bd) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += __ldg(partial + i + k * bd);
  }
  for (; i < n; i += bd) a[0] += __ldg(partial + i);
  float s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[0] = v * scale;
  }
}

// out[j] = sum_{r < nrows} partial[r][j]   (fixed order), optional accumulate into out
// Block = 32 columns x 32 row groups (1024 threads); each thread sums its rows (r = group,
// group + 32, ...) with 4 independent accumulators, then the 32 groups are combined in a fixed
// order. These reductions are pure latency (a few hundred KB): the wide block keeps 4 x 1024 loads
// in flight per 32 columns instead of walking the rows with a handful of threads.
constexpr int RR_GROUPS = 32;
constexpr int RR_THREADS = 32 * RR_GROUPS;
__global__ void __launch_bounds__(RR_THREADS)
reduce_rows_kernel(const float* __restrict__ partial, int nrows, int row_stride, int ncols,
                   float* __restrict__ out, int accumulate, int rows_per_chunk = 0,
                   int out_stride = 0) {
  pdl_trigger();
  pdl_wait();
  // blockIdx.y selects a chunk of rows (two-stage reduction of very tall partial buffers);
  // with gridDim.y == 1 the whole buffer is reduced straight into `out`.
  __shared__ float sred[RR_GROUPS][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cl;
  const int r_begin = rows_per_chunk > 0 ? blockIdx.y * rows_per_chunk : 0;
  const int r_end = rows_per_chunk > 0 ? min(nrows, r_begin + rows_per_chunk) : nrows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (j < ncols) {
    const float* p = partial + j;
    int r = r_begin + rg;
    for (; r + 3 * RR_GROUPS < r_end; r += 4 * RR_GROUPS) {
      a0 += __ldg(p + static_cast<long long>(r) * row_stride);
      a1 += __ldg(p + static_cast<long long>(r + RR_GROUPS) * row_stride);
      a2 += __ldg(p + static_cast<long long>(r + 2 * RR_GROUPS) * row_stride);
      a3 += __ldg(p + static_cast<long long>(r + 3 * RR_GROUPS) * row_stride);
    }
    for (; r < r_end; r += RR_GROUPS) a0 += __ldg(p + static_cast<long long>(r) * row_stride);
  }
  sred[rg][cl] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (rg == 0 && j < ncols) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < RR_GROUPS; ++g) s += sred[g][cl];
    float* o = out + static_cast<long long>(blockIdx.y) * out_stride + j;
    *o = accumulate ? *o + s : s;
  }
}

// [nrows][row_stride] partials -> out[ncols], fixed order. Up to 2048 rows one block per 32 columns walks
// them all (64 loads per thread, four in flight); taller buffers go through `tmp` ([<=128][row_stride]) in
// chunks of at least 256 rows, so that no block is launched for a handful of rows.
constexpr int RR_MAX_CHUNKS = 128;
inline void launch_reduce_rows(const float* partial, int nrows, int row_stride, int ncols, float* out, float* tmp,
                               cudaStream_t st) {
  if (nrows <= 2048 || tmp == nullptr) {
    launch_chain(reduce_rows_kernel, dim3((ncols + 31) / 32), dim3(RR_THREADS), 0, st, 1, partial, nrows, row_stride, ncols,
                 out, 0, 0, 0);
    return;
  }
  int rpc = (nrows + RR_MAX_CHUNKS - 1) / RR_MAX_CHUNKS;
  if (rpc < 256) rpc = 256;
  const int chunks = (nrows + rpc - 1) / rpc;
  launch_chain(reduce_rows_kernel, dim3((ncols + 31) / 32, chunks), dim3(RR_THREADS), 0, st, 1, partial, nrows, row_stride,
               ncols, tmp, 0, rpc, row_stride);
  launch_chain(reduce_rows_kernel, dim3((ncols + 31) / 32), dim3(RR_THREADS), 0, st, 1, tmp, chunks, row_stride, ncols, out,
               0, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// Column sums of an NHWC bf16 tensor (bias gradients): stage 1 writes [grid][C] partials.
// blockDim = 256: thread t owns channel pair (t % (C/2)) of pixel lane (t / (C/2)).
// ---------------------------------------------------------------------------------------------
__global__ void colsum_partial_kernel(const __nv_bfloat16* __restrict__ g, int ld, long long npix,
                                      int C, float* __restrict__ partial) {
  extern __shared__ float ssum[];  // [blockDim.x * 2]
  const int c2 = C >> 1;
  const int pl = threadIdx.x / c2;          // pixel lane
  const int cp = threadIdx.x - pl * c2;     // channel pair
  const int plane = blockDim.x / c2;        // pixel lanes per block
  float s0 = 0.f, s1 = 0.f;
  if (pl < plane) {
    for (long long px = static_cast<long long>(blockIdx.x) * plane + pl; px < npix;
         px += static_cast<long long>(gridDim.x) * plane) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(g + px * ld) + cp);
      s0 += bf16_lo(u);
      s1 += bf16_hi(u);
    }
  }
  ssum[2 * threadIdx.x] = s0;
  ssum[2 * threadIdx.x + 1] = s1;
  __syncthreads();
  if (threadIdx.x < c2) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < plane; ++l) {
      a += ssum[2 * (l * c2 + threadIdx.x)];
      b += ssum[2 * (l * c2 + threadIdx.x) + 1];
    }
    partial[static_cast<long long>(blockIdx.x) * C + 2 * threadIdx.x] = a;
    partial[static_cast<long long>(blockIdx.x) * C + 2 * threadIdx.x + 1] = b;
  }
}

// Column sums of a wide bf16 row-major matrix [M][C] (bias gradients of the attention / feed-forward
// projections, C up to 9216): block = 64 columns (32 lanes x one bf16 pair) x 8 row groups over a
// chunk of rows; partial [chunks][C] is combined by reduce_rows_kernel.
__global__ void __launch_bounds__(256)
colsum_wide_kernel(const __nv_bfloat16* __restrict__ g, long long ld, long long M, int C, int rows_per_chunk,
                   float* __restrict__ partial) {
  __shared__ float sred[8][64];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * lane;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_chunk;
  const long long r1 = min(M, r0 + rows_per_chunk);
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    for (long long r = r0 + rg; r < r1; r += 8) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(g + r * ld + c));
      s0 += bf16_lo(u);
      s1 += bf16_hi(u);
    }
  }
  sred[rg][2 * lane] = s0;
  sred[rg][2 * lane + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sred[k][threadIdx.x];
    partial[static_cast<long long>(blockIdx.y) * C + blockIdx.x * 64 + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Weight repack: dst[i0][i1][i2] (bf16, dense) = src[i0*s0 + i1*s1 + i2*s2] (fp32) for i2 < v2,
// i0 < v0, else 0.
// ---------------------------------------------------------------------------------------------
__global__ void repack_weights_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      int d0, int d1, int d2, long long s0, long long s1,
                                      long long s2, int v0, int v2) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(d0) * d1 * d2;
  if (i >= n) return;
  const int i2 = static_cast<int>(i % d2);
  const int i1 = static_cast<int>((i / d2) % d1);
  const int i0 = static_cast<int>(i / (static_cast<long long>(d1) * d2));
  float v = 0.f;
  if (i2 < v2 && i0 < v0) v = src[i0 * s0 + i1 * s1 + i2 * s2];
  dst[i] = __float2bfloat16_rn(v);
}

// The same for up to REPACK_MAX weight tensors in ONE launch (all layers of a network after an optimizer
// step): entry e covers flat output elements [start_e, start_{e+1}).
constexpr int REPACK_MAX = 40;
struct RepackEntry {
  const float* src;
  __nv_bfloat16* dst;
  long long s0, s1, s2, start;
  int d0, d1, d2, v0, v2, pad;
};
struct RepackTable {
  int n;
  int pad;
  long long total;
  RepackEntry e[REPACK_MAX];
};
__global__ void repack_batch_kernel(const __grid_constant__ RepackTable tab) {
  pdl_trigger();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= tab.total) return;
  int e = 0;
  while (e + 1 < tab.n && i >= tab.e[e + 1].start) ++e;
  const RepackEntry& t = tab.e[e];
  const long long k = i - t.start;
  const int i2 = static_cast<int>(k % t.d2);
  const int i1 = static_cast<int>((k / t.d2) % t.d1);
  const int i0 = static_cast<int>(k / (static_cast<long long>(t.d1) * t.d2));
  float v = 0.f;
  if (i2 < t.v2 && i0 < t.v0) v = t.src[i0 * t.s0 + i1 * t.s1 + i2 * t.s2];
  t.dst[k] = __float2bfloat16_rn(v);
}

}  // namespace rovr
