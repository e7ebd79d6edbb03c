// resnet.cuh — the memory-bound pieces of the ResNet-50 frame-feature extractor
// (rovr/resnet_extractor.py:5-67 around torchvision.models.resnet50). The convolutions themselves
// run on igemm_kernel (3x3) and as plain GEMMs (1x1, and the 7x7 stem after im2col); what is here:
//   * eval-mode BatchNorm folded into the preceding convolution (weights * gamma/sqrt(var+eps),
//     bias = beta - mean * gamma/sqrt(var+eps)) — the trunk is frozen + .eval() when pretrained
//     (rovr/resnet_extractor.py:11-14);
//   * ToPILImage -> ToTensor quantisation of rovr/resnet_extractor.py:18-23 (uint8 round trip) fused
//     with the stem's im2col (7x7, stride 2, pad 3);
//   * max-pool 3x3 s2 p1, 2x spatial subsampling (stride-2 convolutions), residual add + ReLU,
//     global average pool, and the 3x16x16 tile paste into the 5x5 mosaic (:25-55).
#pragma once
#include "ptx.cuh"

namespace rovr {

// wf[co][k] = w[co][k] * s_co,  bf[co] = beta - mean * s_co,  s_co = gamma / sqrt(var + eps)
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ wf,
                               float* __restrict__ bf, int Cout, int K) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(Cout) * K) return;
  const int co = static_cast<int>(i / K);
  const float s = gamma[co] * rsqrtf(var[co] + eps);
  wf[i] = w[i] * s;
  if (i - static_cast<long long>(co) * K == 0) bf[co] = beta[co] - mean[co] * s;
}

// im2col of the 7x7 stride-2 pad-3 stem: src NCHW fp32 [B][3][H][W] -> dst [B*Ho*Wo][kpad] bf16 with
// k = c*49 + r*7 + s (the flattening of the PyTorch weight [64][3][7][7]); columns >= 147 are 0.
// quantise != 0 applies v -> floor(clamp(v,0,1)*255)/255, the ToPILImage/ToTensor round trip.
__global__ void stem_im2col_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B,
                                   int H, int W, int Ho, int Wo, int kpad, int quantise) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * Ho * Wo * kpad;
  if (i >= n) return;
  const int k = static_cast<int>(i % kpad);
  long long r = i / kpad;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  float v = 0.f;
  if (k < 147) {
    const int c = k / 49, rr = (k % 49) / 7, ss = k % 7;
    const int iy = oy * 2 - 3 + rr, ix = ox * 2 - 3 + ss;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      v = src[((static_cast<long long>(b) * 3 + c) * H + iy) * W + ix];
      if (quantise) v = floorf(fminf(fmaxf(v, 0.f), 1.f) * 255.f) * (1.f / 255.f);
    }
  }
  dst[i] = __float2bfloat16_rn(v);
}

// max-pool with zero... no: -inf padding (nn.MaxPool2d(3, 2, 1)); NHWC bf16, 8 channels per thread
__global__ void maxpool_pad_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld,
                                       __nv_bfloat16* __restrict__ y, int y_ld, int B, int H, int W, int C,
                                       int k, int s, int pad, int Ho, int Wo) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * Ho * Wo * c8) return;
  const int cv = static_cast<int>(i % c8);
  long long r = i / c8;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  float best[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
  for (int dy = 0; dy < k; ++dy) {
    const int iy = oy * s - pad + dy;
    if (iy < 0 || iy >= H) continue;
    for (int dx = 0; dx < k; ++dx) {
      const int ix = ox * s - pad + dx;
      if (ix < 0 || ix >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + iy) * W + ix) * x_ld + cv * 8));
      const __nv_bfloat16* v8 = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], __bfloat162float(v8[j]));
    }
  }
  uint4 o;
  __nv_bfloat16* o8 = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
  for (int j = 0; j < 8; ++j) o8[j] = __float2bfloat16_rn(best[j]);
  *reinterpret_cast<uint4*>(y + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * y_ld + cv * 8) = o;
}

// y[b][oy][ox][:] = x[b][oy*s][ox*s][:]  (the sampling pattern of a stride-s convolution)
__global__ void subsample_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y,
                                 int y_ld, int B, int H, int W, int C, int s, int Ho, int Wo) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * Ho * Wo * c8) return;
  const int cv = static_cast<int>(i % c8);
  long long r = i / c8;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + oy * s) * W + ox * s) * x_ld + cv * 8));
  *reinterpret_cast<uint4*>(y + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * y_ld + cv * 8) = v;
}

// out = relu(a + b) on dense bf16 tensors (the residual join of a bottleneck block)
__global__ void add_relu_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a) + i);
  const uint4 ub = __ldg(reinterpret_cast<const uint4*>(b) + i);
  const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = pack_bf16x2(fmaxf(bf16_lo(wa[j]) + bf16_lo(wb[j]), 0.f), fmaxf(bf16_hi(wa[j]) + bf16_hi(wb[j]), 0.f));
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// global average pool: NHWC bf16 [B][HW][C] -> fp32 [B][C]; one thread per (b, channel pair)
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ x, int ld, float* __restrict__ out, int B,
                               int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = C >> 1;
  if (i >= B * c2) return;
  const int b = i / c2, cp = i - b * c2;
  float s0 = 0.f, s1 = 0.f;
  for (int p = 0; p < HW; ++p) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(x + (static_cast<long long>(b) * HW + p) * ld) + cp);
    s0 += bf16_lo(u);
    s1 += bf16_hi(u);
  }
  out[static_cast<long long>(b) * C + 2 * cp] = s0 / HW;
  out[static_cast<long long>(b) * C + 2 * cp + 1] = s1 / HW;
}

// feature rows [n][ch*t*t] (fp32, channel-major like .view(3,16,16)) pasted into mosaics
// [nb][ch][side][side] at (slot // per_row * t, slot % per_row * t): rovr/resnet_extractor.py:30-40,49-55.
// row r goes to mosaic batch[r] (or r / slots_per_mosaic when batch == NULL), slot slot[r] (or r % spm).
__global__ void mosaic_paste_kernel(const float* __restrict__ feat, float* __restrict__ mosaic,
                                    const long long* __restrict__ batch, const long long* __restrict__ slot,
                                    int n, int spm, int ch, int t, int per_row, int side) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = ch * t * t;
  if (i >= n * per) return;
  const int r = i / per, e = i - r * per;
  const int c = e / (t * t), yy = (e / t) % t, xx = e % t;
  const long long bsel = batch ? batch[r] : r / spm;
  const long long ssel = slot ? slot[r] : r % spm;
  const int y0 = static_cast<int>(ssel / per_row) * t, x0 = static_cast<int>(ssel % per_row) * t;
  mosaic[((bsel * ch + c) * side + y0 + yy) * side + x0 + xx] = feat[i];
}
// the inverse (gradient of the paste w.r.t. the feature rows)
__global__ void mosaic_gather_kernel(const float* __restrict__ mosaic, float* __restrict__ feat,
                                     const long long* __restrict__ batch, const long long* __restrict__ slot,
                                     int n, int spm, int ch, int t, int per_row, int side) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = ch * t * t;
  if (i >= n * per) return;
  const int r = i / per, e = i - r * per;
  const int c = e / (t * t), yy = (e / t) % t, xx = e % t;
  const long long bsel = batch ? batch[r] : r / spm;
  const long long ssel = slot ? slot[r] : r % spm;
  const int y0 = static_cast<int>(ssel / per_row) * t, x0 = static_cast<int>(ssel % per_row) * t;
  feat[i] = mosaic[((bsel * ch + c) * side + y0 + yy) * side + x0 + xx];
}

// transforms.Resize on a PIL image = PIL's two-pass 8-bit resampler (bilinear filter whose support
// grows with the down-sampling factor, i.e. antialiased): horizontal pass into a uint8 image, then
// vertical pass, each with 22-bit fixed-point coefficients and round-half-up. One pass here resamples
// dimension `len_in -> len_out` of a tensor viewed as [outer][len][inner]; values are uint8 levels
// held in fp32. first != 0 quantises the input like ToPILImage (floor(clamp(v,0,1)*255)); last != 0
// divides by 255 like ToTensor.
__global__ void resize_pass_kernel(const float* __restrict__ src, float* __restrict__ dst, long long outer,
                                   int len_in, int len_out, int inner, int first, int last) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= outer * len_out * inner) return;
  const int in = static_cast<int>(i % inner);
  const int o = static_cast<int>((i / inner) % len_out);
  const long long ou = i / (static_cast<long long>(inner) * len_out);
  const double scale = static_cast<double>(len_in) / len_out;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * fscale;  // bilinear filter support = 1
  const double center = (o + 0.5) * scale;
  int xmin = static_cast<int>(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(center + support + 0.5);
  if (xmax > len_in) xmax = len_in;
  const int cnt = xmax - xmin;
  double wsum = 0.0;
  for (int x = 0; x < cnt; ++x) {
    double t = (x + xmin - center + 0.5) / fscale;
    if (t < 0) t = -t;
    wsum += t < 1.0 ? 1.0 - t : 0.0;
  }
  long long acc = 1ll << 21;  // 0.5 in 22-bit fixed point
  for (int x = 0; x < cnt; ++x) {
    double t = (x + xmin - center + 0.5) / fscale;
    if (t < 0) t = -t;
    const double w = (t < 1.0 ? 1.0 - t : 0.0) / wsum;
    const long long k = static_cast<long long>(w < 0 ? -0.5 + w * 4194304.0 : 0.5 + w * 4194304.0);
    float v = src[(ou * len_in + xmin + x) * inner + in];
    if (first) v = floorf(fminf(fmaxf(v, 0.f), 1.f) * 255.f);
    acc += k * static_cast<long long>(v);
  }
  long long q = acc >> 22;
  q = q < 0 ? 0 : (q > 255 ? 255 : q);
  dst[i] = last ? static_cast<float>(q) * (1.f / 255.f) : static_cast<float>(q);
}

}  // namespace rovr
