// rovr_b200.cu — C-ABI entry points (include/rovr_b200.h): tensor-map construction, tiling
// decisions and kernel launches. Single translation unit; compiled for sm_100a only.
#include "../../include/rovr_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "elementwise.cuh"
#include "fp32x.cuh"
#include "lpips.cuh"
#include "igemm.cuh"
#include "norm.cuh"
#include "policy.cuh"
#include "resnet.cuh"
#include "attention.cuh"
#include "tail.cuh"
#include "wgrad.cuh"
#include "wgrad_halo.cuh"

using namespace rovr;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define ROVR_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return fail(-2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                              \
  } while (0)
#define ROVR_REQUIRE(cond, ...) \
  do {                          \
    if (!(cond)) return fail(-1, __VA_ARGS__); \
  } while (0)
static std::atomic<unsigned long long> g_launches{0};
static int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-3, "launch of %s failed: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
extern "C" unsigned long long rovr_launch_count(void) { return g_launches.load(); }

extern "C" int rovr_abi_version(void) { return 1; }
extern "C" const char* rovr_last_error(void) { return g_err; }

struct DeviceInfo {
  int ok = 0;
  int sms = 0;
  int major = 0, minor = 0;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
};
static DeviceInfo g_dev;
static std::once_flag g_dev_once;
static int g_dev_rc = 0;

static void init_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    g_dev_rc = fail(-4, "no CUDA device: the ROVR B200 path has no CPU fallback");
    return;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    g_dev_rc = fail(-4, "cudaGetDeviceProperties failed");
    return;
  }
  g_dev.sms = prop.multiProcessorCount;
  g_dev.major = prop.major;
  g_dev.minor = prop.minor;
  if (prop.major != 10) {
    g_dev_rc = fail(-4, "device is sm_%d%d; this library is built for sm_100a (B200) only",
                    prop.major, prop.minor);
    return;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      fn == nullptr) {
    g_dev_rc = fail(-4, "cuTensorMapEncodeTiled entry point not found");
    return;
  }
  g_dev.encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  g_dev.ok = 1;
}
// Function attributes (227 KB of dynamic shared memory) belong to a (function, device) pair: they are set
// once on EVERY device a call is made on, not just on the first one (a second GPU used by the same process
// would otherwise fail its launches).
static std::mutex g_attr_mutex;
static unsigned long long g_attr_mask = 0;   // bit d: device d has its attributes
template <typename K>
static void set_max_smem(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
static void init_function_attributes();
static int ensure_device() {
  std::call_once(g_dev_once, init_device);
  if (!g_dev.ok) {
    if (g_err[0] == 0) fail(-4, "device initialisation failed earlier (not sm_100 or no driver)");
    return g_dev_rc ? g_dev_rc : -4;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(-4, "cudaGetDevice failed");
  if (!(__atomic_load_n(&g_attr_mask, __ATOMIC_ACQUIRE) >> dev & 1ull)) {
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    if (!(g_attr_mask >> dev & 1ull)) {
      int major = 0;
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
      if (major != 10) return fail(-4, "device %d is not sm_100: this library is built for sm_100a (B200) only", dev);
      init_function_attributes();
      __atomic_store_n(&g_attr_mask, g_attr_mask | (1ull << dev), __ATOMIC_RELEASE);
    }
  }
  return 0;
}
static void init_function_attributes() {
  set_max_smem(igemm_kernel<true>);
  set_max_smem(igemm_kernel<false>);
  set_max_smem(igemm_kernel<true, true>);
  set_max_smem(igemm_kernel<true, false, true>);
  set_max_smem(igemm_kernel<true, false, false, true>);
  set_max_smem(igemm_kernel<false, false, false, true>);
  set_max_smem(igemm_kernel<true, true, false, true>);
  set_max_smem(igemm_kernel<true, false, true, true>);
  set_max_smem(wgrad_kernel);
  set_max_smem(wgrad_halo_kernel);
}
extern "C" int rovr_device_check(void) { return ensure_device(); }
extern "C" int rovr_hang_code(unsigned int* code) {
  unsigned int v = 0;
  ROVR_CUDA(cudaMemcpyFromSymbol(&v, g_rovr_hang_code, sizeof(v)));
  *code = v;
  if (v != 0) {  // reading clears it so that later launches run normally
    const unsigned int zero = 0;
    ROVR_CUDA(cudaMemcpyToSymbol(g_rovr_hang_code, &zero, sizeof(zero)));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
static CUtensorMapSwizzle swizzle_for(int sw_bytes) {
  return sw_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                         : (sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
// rank-5 bf16 map: dims[0] is the contiguous (channel) dim; strides_elems[i] is the stride of
// dim i+1 in elements.
// elem_strides (optional): traversal stride per dim — the box then covers box[i] global elements and delivers
// every elem_strides[i]-th of them (a stride-2 convolution reads every other input pixel this way).
static int make_map5(CUtensorMap* m, const void* base, const long long dims[5],
                     const long long strides_elems[4], const int box[5], int sw_bytes,
                     const int* elem_strides = nullptr) {
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  if (elem_strides != nullptr)
    for (int i = 0; i < 5; ++i) es[i] = static_cast<cuuint32_t>(elem_strides[i]);
  for (int i = 0; i < 5; ++i) {
    gd[i] = static_cast<cuuint64_t>(dims[i]);
    bx[i] = static_cast<cuuint32_t>(box[i]);
  }
  for (int i = 0; i < 4; ++i) gs[i] = static_cast<cuuint64_t>(strides_elems[i]) * 2ull;
  ROVR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "tensor base %p not 16-byte aligned",
               base);
  for (int i = 0; i < 4; ++i)
    ROVR_REQUIRE((gs[i] & 15u) == 0 && gs[i] > 0, "tensor stride %d (%llu B) not a multiple of 16",
                 i, static_cast<unsigned long long>(gs[i]));
  ROVR_REQUIRE(box[0] * 2 == sw_bytes, "inner box (%d elems) must equal the swizzle span", box[0]);
  // A map over a <= 64-channel SLICE of a wider pixel (e.g. the up-conv half of a concat gradient)
  // only ever touches 128 of every 256+ bytes: promoting its misses to 256 B would drag the unused
  // half through DRAM (ncu: upconv3 dgrad read 502 MB for 301 MB of operands). Everything else is
  // read in full, where 256 B promotion prefetches the neighbouring chunk / pixel.
  const bool narrow_slice = dims[0] * 2 <= 128 && strides_elems[0] > dims[0];
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gd, gs,
                            bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(sw_bytes),
                            narrow_slice ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-5,
                "cuTensorMapEncodeTiled(5d) failed rc=%d dims=[%lld,%lld,%lld,%lld,%lld] "
                "box=[%d,%d,%d,%d,%d]",
                static_cast<int>(r), dims[0], dims[1], dims[2], dims[3], dims[4], box[0], box[1],
                box[2], box[3], box[4]);
  return 0;
}
// rank-2 bf16 map [rows][cols] row-major, box [box_rows][box_cols]
static int make_map2(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                     int box_rows, int box_cols, int sw_bytes) {
  cuuint64_t gd[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gs[1] = {static_cast<cuuint64_t>(ld) * 2ull};
  cuuint32_t bx[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t es[2] = {1, 1};
  ROVR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "weight base not 16-byte aligned");
  ROVR_REQUIRE((gs[0] & 15u) == 0, "weight row stride not a multiple of 16 bytes");
  ROVR_REQUIRE(box_cols * 2 == sw_bytes, "weight inner box must equal the swizzle span");
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gd, gs,
                            bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(sw_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-5, "cuTensorMapEncodeTiled(2d) failed rc=%d rows=%lld cols=%lld box=[%d,%d]",
                static_cast<int>(r), rows, cols, box_rows, box_cols);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// tiling helpers
// ------------------------------------------------------------------------------------------------
static int ceil_div(int a, int b) { return (a + b - 1) / b; }
static int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}
static int pick_bk(int c) { return (c % 64 == 0) ? 64 : ((c % 32 == 0) ? 32 : 16); }

// Choose a (tw, th, nb) patch of <= max_rows pixels that wastes the fewest rows over a
// [B][H][W] map. Patches may span several images only when they cover a whole image.
static void pick_patch(int W, int H, int B, int max_rows, int* tw_o, int* th_o, int* nb_o) {
  double best = -1.0;
  int btw = 1, bth = 1, bnb = 1;
  for (int tw = 1; tw <= std::min(W, std::min(max_rows, 256)); ++tw) {
    for (int th = 1; th <= std::min(H, max_rows / tw); ++th) {
      int nb = 1;
      if (tw == W && th == H) nb = std::max(1, std::min(B, max_rows / (tw * th)));
      const double tiles = static_cast<double>(ceil_div(W, tw)) * ceil_div(H, th) * ceil_div(B, nb);
      const double eff = (static_cast<double>(W) * H * B) / (tiles * max_rows);
      const bool better = eff > best + 1e-9 || (eff > best - 1e-9 && tw > btw);
      if (better) {
        best = eff;
        btw = tw;
        bth = th;
        bnb = nb;
      }
    }
  }
  *tw_o = btw;
  *th_o = bth;
  *nb_o = bnb;
}


// ------------------------------------------------------------------------------------------------
// packing / repacking
// ------------------------------------------------------------------------------------------------
extern "C" int rovr_pack_nchw_to_nhwc(const float* s0, int c0, const float* s1, int c1,
                                      const float* s2, int c2, void* dst, int B, int H, int W,
                                      int cpad, void* stream) {
  if (int rc = ensure_device()) return rc;
  PackSrc src;
  src.nsrc = 0;
  const float* ptrs[3] = {s0, s1, s2};
  const int chs[3] = {c0, c1, c2};
  int tot = 0;
  for (int i = 0; i < 3; ++i)
    if (ptrs[i] && chs[i] > 0) {
      src.ptr[src.nsrc] = ptrs[i];
      src.ch[src.nsrc] = chs[i];
      ++src.nsrc;
      tot += chs[i];
    }
  ROVR_REQUIRE(tot <= cpad, "pack: %d source channels exceed cpad %d", tot, cpad);
  const long long n = 1ll * B * H * W;
  launch_chain(pack_nchw_to_nhwc_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0,
               static_cast<cudaStream_t>(stream), 1, src, static_cast<__nv_bfloat16*>(dst), B, H * W, cpad);
  return launch_check("pack_nchw_to_nhwc");
}
extern "C" int rovr_unpack_nhwc_to_nchw(const void* src, int ld, float* dst, int B, int H, int W,
                                        int C, void* stream) {
  if (int rc = ensure_device()) return rc;
  const long long n = 1ll * B * H * W;
  unpack_nhwc_to_nchw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                               static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), ld, dst, B, H * W, C);
  return launch_check("unpack_nhwc_to_nchw");
}

extern "C" int rovr_u8_to_f32(const void* src, float* dst, long long n, float denom, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0,
               "u8_to_f32: pointers must be 16-byte aligned");
  const long long threads = (n + 15) / 16;
  u8_to_f32_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), dst, n, denom);
  return launch_check("u8_to_f32");
}

extern "C" int rovr_corrupt_frames(const float* clean, const long long* frame_index, float* out, float* mask, int N,
                                   int C, int H, int W, int box_w, int box_h, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && box_w > 0 && box_h > 0, "corrupt_frames: empty problem");
  const long long n = 1ll * N * C * H * W;
  corrupt_frames_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      clean, frame_index, out, mask, N, C, H, W, box_w, box_h);
  return launch_check("corrupt_frames");
}

static int repack(const float* w, void* wk, int d0, int d1, int d2, long long s0, long long s1,
                  long long s2, int v0, int v2, void* stream, long long off = 0) {
  w += off;
  if (int rc = ensure_device()) return rc;
  const long long n = 1ll * d0 * d1 * d2;
  repack_weights_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(wk), d0, d1, d2, s0, s1, s2, v0, v2);
  return launch_check("repack_weights");
}
extern "C" int rovr_repack_conv3x3_fprop(const float* w, void* wk, int Cout, int Cin, int cin_pad,
                                         void* stream) {
  // dst[co][t][ci] = w[co][ci][t]
  return repack(w, wk, Cout, 9, cin_pad, 1ll * Cin * 9, 1, 9, Cout, Cin, stream);
}
extern "C" int rovr_repack_conv3x3_dgrad(const float* w, void* wk, int Cout, int Cin, int cin_pad,
                                         void* stream) {
  // dst[ci][t][co] = w[co][ci][8 - t]: the 180-degree tap flip of the transposed convolution is
  // baked into the packing, so dgrad runs the forward kernel unchanged (same tap geometry)
  return repack(w, wk, cin_pad, 9, Cout, 9, -1, 1ll * Cin * 9, Cin, Cout, stream, 8);
}
extern "C" int rovr_repack_convT2x2_fprop(const float* w, void* wk, int Cin, int Cout, void* stream) {
  // dst[q][co][ci] = w[ci][co][q]
  return repack(w, wk, 4, Cout, Cin, 1, 4, 1ll * Cout * 4, 4, Cin, stream);
}
extern "C" int rovr_repack_convT2x2_dgrad(const float* w, void* wk, int Cin, int Cout, void* stream) {
  // dst[ci][q][co] = w[ci][co][q]
  return repack(w, wk, Cin, 4, Cout, 1ll * Cout * 4, 1, 4, Cin, Cout, stream);
}

// All weight tensors of a network in one launch. kind: 0 = Conv2d 3x3 forward operand (a = Cout, b = Cin,
// c = cin_pad), 1 = Conv2d 3x3 data-gradient operand (same), 2 = ConvTranspose2d 2x2 forward operand
// (a = Cin, b = Cout), 3 = ConvTranspose2d 2x2 data-gradient operand (same); layouts as the single calls above.
extern "C" int rovr_repack_batch(const rovr_repack_item* items, int n, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(n >= 1 && n <= REPACK_MAX, "repack_batch: %d items (max %d)", n, REPACK_MAX);
  RepackTable tab;
  memset(&tab, 0, sizeof(tab));
  tab.n = n;
  long long start = 0;
  for (int i = 0; i < n; ++i) {
    const rovr_repack_item& it = items[i];
    RepackEntry& e = tab.e[i];
    e.src = it.w;
    e.dst = static_cast<__nv_bfloat16*>(it.wk);
    switch (it.kind) {
      case 0: e.d0 = it.a; e.d1 = 9; e.d2 = it.c; e.s0 = 9ll * it.b; e.s1 = 1; e.s2 = 9; e.v0 = it.a; e.v2 = it.b; break;
      case 1: e.d0 = it.c; e.d1 = 9; e.d2 = it.a; e.s0 = 9; e.s1 = -1; e.s2 = 9ll * it.b; e.v0 = it.b; e.v2 = it.a;
              e.src += 8; break;
      case 2: e.d0 = 4; e.d1 = it.b; e.d2 = it.a; e.s0 = 1; e.s1 = 4; e.s2 = 4ll * it.b; e.v0 = 4; e.v2 = it.a; break;
      case 3: e.d0 = it.a; e.d1 = 4; e.d2 = it.b; e.s0 = 4ll * it.b; e.s1 = 1; e.s2 = 4; e.v0 = it.a; e.v2 = it.b; break;
      default: return fail(-2, "repack_batch: unknown kind %d", it.kind);
    }
    e.start = start;
    start += 1ll * e.d0 * e.d1 * e.d2;
  }
  tab.total = start;
  launch_chain(repack_batch_kernel, dim3(static_cast<unsigned>((start + 255) / 256)), dim3(256), 0,
               static_cast<cudaStream_t>(stream), 1, tab);
  return launch_check("repack_batch");
}

// nn.Linear / 1x1-conv weight [N][K] fp32 -> bf16 [n_pad][k_pad] (transpose = 0) or its transpose
// [k_pad][n_pad] (transpose = 1, the operand of the data gradient), zero padded.
extern "C" int rovr_repack_linear(const float* w, void* wk, int N, int K, int n_pad, int k_pad,
                                  int transpose, void* stream) {
  if (transpose) return repack(w, wk, k_pad, 1, n_pad, 1, 0, K, K, N, stream);
  return repack(w, wk, n_pad, 1, k_pad, K, 0, 1, N, K, stream);
}

#include "api_igemm.inc"
#include "api_wgrad.inc"

// ------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------
extern "C" int rovr_maxpool_fwd(const void* x, int x_ld, void* y, int y_ld, int B, int H, int W,
                                int C, int kh, int kw, int sh, int sw, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0, "maxpool_fwd: C and ld must be multiples of 8");
  const int Ho = (H - kh) / sh + 1, Wo = (W - kw) / sw + 1;
  const long long n = 1ll * B * Ho * Wo * (C / 8);
  if (kh == 2 && kw == 2 && sh == 2 && sw == 2 && H % 2 == 0 && W % 2 == 0) {
    maxpool_fwd_2x2_kernel<<<static_cast<unsigned>((n + 511) / 512), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<__nv_bfloat16*>(y), y_ld, B, H, W, C);
    return launch_check("maxpool_fwd_2x2");
  }
  maxpool_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<__nv_bfloat16*>(y), y_ld, B, H, W, C,
      kh, kw, sh, sw, Ho, Wo);
  return launch_check("maxpool_fwd");
}
extern "C" size_t rovr_maxpool_bwd_colsum_workspace(int B, int H, int W, int C, int kh, int kw) {
  const long long n = 1ll * B * (H / kh) * (W / kw) * (C / 8);
  const long long blocks = (n + 255) / 256;
  return static_cast<size_t>(blocks + 256) * C * sizeof(float);
}
extern "C" int rovr_maxpool_bwd(const void* x, int x_ld, const void* gp, int gp_ld,
                                const void* gskip, int gs_ld, void* gx, int gx_ld, int B, int H,
                                int W, int C, int kh, int kw, int sh, int sw, int relu_mask,
                                float* colsum, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tiled = (kh == sh && kw == sw && H % kh == 0 && W % kw == 0 && C % 8 == 0 &&
                      x_ld % 8 == 0 && gp_ld % 8 == 0 && gx_ld % 8 == 0 && (gskip == nullptr || gs_ld % 8 == 0));
  if (tiled) {
    const long long n = 1ll * B * (H / kh) * (W / kw) * (C / 8);
    int blocks = static_cast<int>((n + 255) / 256);
    if (colsum != nullptr) blocks = std::min(blocks, 8 * g_dev.sms);   // grid-stride: few, fat partial rows
    float* partial = nullptr;
    if (colsum != nullptr) {
      ROVR_REQUIRE(C <= 256 && 256 % (C / 8) == 0, "maxpool_bwd: fused column sums need C <= 256 and C/8 a power of two");
      ROVR_REQUIRE(ws != nullptr && ws_bytes >= rovr_maxpool_bwd_colsum_workspace(B, H, W, C, kh, kw),
                   "maxpool_bwd: column-sum workspace too small");
      partial = static_cast<float*>(ws);
    }
    if (kh == 2 && kw == 2 && (C / 8 <= 32) && 256 % (C / 8) == 0) {
      blocks = std::min(blocks, 8 * g_dev.sms);
      launch_chain(maxpool_bwd_2x2_kernel, dim3(blocks), dim3(256), 0, st, 1,
                   static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<const __nv_bfloat16*>(gp), gp_ld,
                   static_cast<const __nv_bfloat16*>(gskip), gs_ld, static_cast<__nv_bfloat16*>(gx), gx_ld, B, H, W, C,
                   relu_mask, partial);
    } else
    maxpool_bwd_tiled_kernel<<<blocks, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<const __nv_bfloat16*>(gp), gp_ld,
        static_cast<const __nv_bfloat16*>(gskip), gs_ld, static_cast<__nv_bfloat16*>(gx), gx_ld, B,
        H, W, C, kh, kw, relu_mask, partial);
    if (int rc = launch_check("maxpool_bwd_tiled")) return rc;
    if (colsum == nullptr) return 0;
    // [blocks][C] -> [chunks][C] -> [C], both stages in a fixed order
    launch_reduce_rows(partial, blocks, C, C, colsum, partial + static_cast<size_t>(blocks) * C, st);
    return launch_check("reduce_rows(pool colsum)");
  }
  ROVR_REQUIRE(gskip == nullptr, "maxpool_bwd: skip gradient only supported for non-overlapping windows");
  ROVR_REQUIRE(colsum == nullptr, "maxpool_bwd: fused column sums only supported for non-overlapping windows");
  const int Ho = (H - kh) / sh + 1, Wo = (W - kw) / sw + 1;
  const long long n = 1ll * B * H * W * C;
  maxpool_bwd_generic_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<const __nv_bfloat16*>(gp), gp_ld,
      static_cast<__nv_bfloat16*>(gx), gx_ld, B, H, W, C, kh, kw, sh, sw, Ho, Wo, relu_mask);
  return launch_check("maxpool_bwd_generic");
}

// ------------------------------------------------------------------------------------------------
// LocalNet tail
// ------------------------------------------------------------------------------------------------
#include "api_tail.inc"


// ------------------------------------------------------------------------------------------------
// normalisation, fp32 linear layers, policy heads, LSTM
// ------------------------------------------------------------------------------------------------
#include "api_policy.inc"
#include "api_fp32x.inc"
#include "api_lpips.inc"
#include "api_resnet.inc"
#include "api_attention.inc"

// ------------------------------------------------------------------------------------------------
// bias gradient
// ------------------------------------------------------------------------------------------------
// wide matrices (any even C): rows split into up to 64 chunks
static const int kColsumWideChunks = 64;
extern "C" size_t rovr_colsum_rows_workspace(int C) { return static_cast<size_t>(kColsumWideChunks) * C * sizeof(float); }
extern "C" int rovr_colsum_rows(const void* g, long long ld, long long M, int C, float* out, void* ws,
                                size_t ws_bytes, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(C % 2 == 0 && ld % 2 == 0 && M > 0, "colsum_rows: C and ld must be even");
  ROVR_REQUIRE(ws_bytes >= rovr_colsum_rows_workspace(C), "colsum_rows: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rpc = static_cast<int>((M + kColsumWideChunks - 1) / kColsumWideChunks);
  const int chunks = static_cast<int>((M + rpc - 1) / rpc);
  colsum_wide_kernel<<<dim3((C + 63) / 64, chunks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), ld, M, C, rpc,
                                                                 static_cast<float*>(ws));
  if (int rc = launch_check("colsum_wide")) return rc;
  reduce_rows_kernel<<<(C + 31) / 32, RR_THREADS, 0, st>>>(static_cast<float*>(ws), chunks, C, C, out, 0);
  return launch_check("reduce_rows(colsum_rows)");
}

static const int kColsumBlocks = 592;  // 4 per SM
extern "C" size_t rovr_colsum_workspace(int C) { return static_cast<size_t>(kColsumBlocks) * C * sizeof(float); }
extern "C" int rovr_colsum(const void* g, int ld, long long npix, int C, float* out, void* ws,
                           size_t ws_bytes, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(C % 2 == 0 && C <= 512 && ld % 2 == 0, "colsum: C must be even and <= 512");
  ROVR_REQUIRE(ws_bytes >= rovr_colsum_workspace(C), "colsum: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256 >= C / 2 ? 256 : C / 2;
  const int plane = threads / (C / 2);
  int grid = static_cast<int>(std::min<long long>(kColsumBlocks, (npix + plane - 1) / plane));
  grid = std::max(grid, 1);
  colsum_partial_kernel<<<grid, threads, threads * 2 * sizeof(float), st>>>(
      static_cast<const __nv_bfloat16*>(g), ld, npix, C, static_cast<float*>(ws));
  if (int rc = launch_check("colsum_partial")) return rc;
  reduce_rows_kernel<<<(C + 31) / 32, RR_THREADS, 0, st>>>(static_cast<float*>(ws), grid, C, C, out, 0);
  return launch_check("reduce_rows(colsum)");
}
