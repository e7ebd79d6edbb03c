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
#include <cstring>
#include <mutex>

#include "elementwise.cuh"
#include "igemm.cuh"
#include "norm.cuh"
#include "wgrad.cuh"

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
  cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  g_dev.ok = 1;
}
static int ensure_device() {
  std::call_once(g_dev_once, init_device);
  if (!g_dev.ok) {
    if (g_err[0] == 0) fail(-4, "device initialisation failed earlier (not sm_100 or no driver)");
    return g_dev_rc ? g_dev_rc : -4;
  }
  return 0;
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
static int make_map5(CUtensorMap* m, const void* base, const long long dims[5],
                     const long long strides_elems[4], const int box[5], int sw_bytes) {
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
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
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gd, gs,
                            bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(sw_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, IgemmParams& p,
                        cudaStream_t st, const char* what) {
  const size_t stage = (128 + static_cast<size_t>(p.n_tile)) * p.bk * 2;
  int stages = static_cast<int>((200 * 1024) / stage);
  stages = std::max(2, std::min(stages, IG_MAX_STAGES));
  p.stages = stages;
  p.tmem_cols = pow2_at_least(2 * p.n_tile);
  ROVR_REQUIRE(p.tmem_cols <= 512, "n_tile %d too large for TMEM double buffering", p.n_tile);
  ROVR_REQUIRE(p.n_total <= 4096, "n_total too large for the bias stage");
  const size_t smem = igemm_smem_bytes(p.bk, p.n_tile, stages, p.n_total);
  ROVR_REQUIRE(smem <= 227 * 1024, "igemm smem %zu exceeds 227 KB", smem);
  const long long m_tiles = 1ll * p.ntile[0] * p.ntile[1] * p.ntile[2] * p.ntile[3];
  const long long tiles = m_tiles * p.n_tiles_n;
  const int grid = static_cast<int>(std::min<long long>(tiles, g_dev.sms));
  igemm_kernel<<<grid, IG_THREADS, smem, st>>>(tmA, tmB, p);
  return launch_check(what);
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
  pack_nchw_to_nhwc_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), B, H * W, cpad);
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

static int repack(const float* w, void* wk, int d0, int d1, int d2, long long s0, long long s1,
                  long long s2, int v0, int v2, void* stream) {
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
  // dst[ci][t][co] = w[co][ci][t]
  return repack(w, wk, cin_pad, 9, Cout, 9, 1, 1ll * Cin * 9, Cin, Cout, stream);
}
extern "C" int rovr_repack_convT2x2_fprop(const float* w, void* wk, int Cin, int Cout, void* stream) {
  // dst[q][co][ci] = w[ci][co][q]
  return repack(w, wk, 4, Cout, Cin, 1, 4, 1ll * Cout * 4, 4, Cin, stream);
}
extern "C" int rovr_repack_convT2x2_dgrad(const float* w, void* wk, int Cin, int Cout, void* stream) {
  // dst[ci][q][co] = w[ci][co][q]
  return repack(w, wk, Cin, 4, Cout, 1ll * Cout * 4, 1, 4, Cin, Cout, stream);
}

// ------------------------------------------------------------------------------------------------
// Conv2d 3x3 pad 1: forward and data gradient share one code path (the tap table differs)
// ------------------------------------------------------------------------------------------------
static int conv3x3_common(const void* x, int x_ld, const void* wk, const float* bias, void* y,
                          int y_ld, const void* mask, int mask_ld, int B, int H, int W, int Cin,
                          int Cout, int relu, bool flip, void* stream, const char* what) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "%s: Cin=%d Cout=%d must be multiples of 16", what,
               Cin, Cout);
  ROVR_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0, "%s: ld must be a multiple of 8", what);
  int tw, th, nb;
  pick_patch(W, H, B, 128, &tw, &th, &nb);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.bk = pick_bk(Cin);
  const int sw = p.bk * 2;
  p.dimM[0] = W; p.dimM[1] = H; p.dimM[2] = B; p.dimM[3] = 1;
  p.boxM[0] = tw; p.boxM[1] = th; p.boxM[2] = nb; p.boxM[3] = 1;
  p.ntile[0] = ceil_div(W, tw); p.ntile[1] = ceil_div(H, th); p.ntile[2] = ceil_div(B, nb); p.ntile[3] = 1;
  p.ostride[0] = y_ld; p.ostride[1] = 1ll * W * y_ld; p.ostride[2] = 1ll * H * W * y_ld; p.ostride[3] = 0;
  p.mstride[0] = mask_ld; p.mstride[1] = 1ll * W * mask_ld; p.mstride[2] = 1ll * H * W * mask_ld; p.mstride[3] = 0;
  p.ntaps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int t = r * 3 + s;
      // forward: out(y,x) reads in(y+r-1, x+s-1); dgrad: dx(y,x) reads dy(y+1-r, x+1-s)
      p.tap_off[t][0] = flip ? (1 - s) : (s - 1);
      p.tap_off[t][1] = flip ? (1 - r) : (r - 1);
      p.tap_off[t][2] = 0;
      p.tap_off[t][3] = 0;
    }
  p.cin = Cin;
  p.n_tile = Cout <= 256 ? Cout : 256;
  ROVR_REQUIRE(Cout % p.n_tile == 0, "%s: Cout=%d not divisible by n_tile", what, Cout);
  p.n_tiles_n = Cout / p.n_tile;
  p.n_total = Cout;
  p.epi_mode = IG_EPI_PLAIN;
  p.relu = relu;
  p.bias = bias;
  p.bias_mod = Cout;
  p.out = static_cast<__nv_bfloat16*>(y);
  p.mask = static_cast<const __nv_bfloat16*>(mask);
  CUtensorMap tmA, tmB;
  const long long dims[5] = {Cin, W, H, B, 1};
  const long long strides[4] = {x_ld, 1ll * W * x_ld, 1ll * H * W * x_ld, 1ll * B * H * W * x_ld};
  const int box[5] = {p.bk, tw, th, nb, 1};
  if (int rc = make_map5(&tmA, x, dims, strides, box, sw)) return rc;
  if (int rc = make_map2(&tmB, wk, Cout, 9ll * Cin, 9ll * Cin, p.n_tile, p.bk, sw)) return rc;
  return launch_igemm(tmA, tmB, p, static_cast<cudaStream_t>(stream), what);
}

extern "C" int rovr_conv3x3_fprop(const void* x, int x_ld, const void* wk, const float* bias,
                                  void* y, int y_ld, int B, int H, int W, int Cin, int Cout,
                                  int relu, void* stream) {
  return conv3x3_common(x, x_ld, wk, bias, y, y_ld, nullptr, 0, B, H, W, Cin, Cout, relu, false,
                        stream, "conv3x3_fprop");
}
extern "C" int rovr_conv3x3_dgrad(const void* dy, int dy_ld, const void* wk_d, void* dx, int dx_ld,
                                  const void* mask, int mask_ld, int B, int H, int W, int Cin,
                                  int Cout, void* stream) {
  // a 3x3 conv over dy with Cout input channels producing Cin channels
  return conv3x3_common(dy, dy_ld, wk_d, nullptr, dx, dx_ld, mask, mask_ld, B, H, W, Cout, Cin, 0,
                        true, stream, "conv3x3_dgrad");
}

// ------------------------------------------------------------------------------------------------
// ConvTranspose2d k2 s2
// ------------------------------------------------------------------------------------------------
extern "C" int rovr_convT2x2_fprop(const void* x, int x_ld, const void* wk, const float* bias,
                                   void* y, int y_ld, int B, int H, int W, int Cin, int Cout,
                                   int relu, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "convT2x2_fprop: Cin=%d Cout=%d must be multiples of 16", Cin, Cout);
  ROVR_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0, "convT2x2_fprop: ld must be a multiple of 8");
  // rows = input pixels; B and H merge into one dim because no halo is needed
  int tw, th, nb;
  pick_patch(W, B * H, 1, 128, &tw, &th, &nb);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.bk = pick_bk(Cin);
  const int sw = p.bk * 2;
  const int BH = B * H;
  p.dimM[0] = W; p.dimM[1] = BH; p.dimM[2] = 1; p.dimM[3] = 1;
  p.boxM[0] = tw; p.boxM[1] = th; p.boxM[2] = 1; p.boxM[3] = 1;
  p.ntile[0] = ceil_div(W, tw); p.ntile[1] = ceil_div(BH, th); p.ntile[2] = 1; p.ntile[3] = 1;
  // output pixel (b, 2y+qy, 2x+qx): merged row index bh = b*H + y -> output row 2*bh (+qy)
  p.ostride[0] = 2ll * y_ld;
  p.ostride[1] = 2ll * (2ll * W) * y_ld;
  p.ntaps = 1;
  p.cin = Cin;
  const int Ntot = 4 * Cout;
  p.n_tile = Ntot <= 256 ? Ntot : 256;
  ROVR_REQUIRE(Ntot % p.n_tile == 0, "convT2x2_fprop: 4*Cout=%d not divisible by n_tile", Ntot);
  p.n_tiles_n = Ntot / p.n_tile;
  p.n_total = Ntot;
  p.epi_mode = IG_EPI_PIXSHUF;
  p.relu = relu;
  p.bias = bias;
  p.bias_mod = Cout;
  p.shuf_cout = Cout;
  p.shuf_sy = 2ll * W * y_ld;
  p.shuf_sx = y_ld;
  p.out = static_cast<__nv_bfloat16*>(y);
  CUtensorMap tmA, tmB;
  const long long dims[5] = {Cin, W, BH, 1, 1};
  const long long strides[4] = {x_ld, 1ll * W * x_ld, 1ll * BH * W * x_ld, 1ll * BH * W * x_ld};
  const int box[5] = {p.bk, tw, th, 1, 1};
  if (int rc = make_map5(&tmA, x, dims, strides, box, sw)) return rc;
  if (int rc = make_map2(&tmB, wk, Ntot, Cin, Cin, p.n_tile, p.bk, sw)) return rc;
  return launch_igemm(tmA, tmB, p, static_cast<cudaStream_t>(stream), "convT2x2_fprop");
}

// 5-D view of a [B][2H][2W][ld] tensor as (C, qx, W, qy, B*H)
static void subpixel_view(int W, int BH, int C, int ld, long long dims[5], long long strides[4]) {
  dims[0] = C; dims[1] = 2; dims[2] = W; dims[3] = 2; dims[4] = BH;
  strides[0] = ld;              // qx
  strides[1] = 2ll * ld;        // W
  strides[2] = 2ll * W * ld;    // qy
  strides[3] = 4ll * W * ld;    // merged (b, y)
}

extern "C" int rovr_convT2x2_dgrad(const void* dy, int dy_ld, const void* wk_d, void* dx, int dx_ld,
                                   const void* mask, int mask_ld, int B, int H, int W, int Cin,
                                   int Cout, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "convT2x2_dgrad: Cin=%d Cout=%d must be multiples of 16", Cin, Cout);
  ROVR_REQUIRE(dy_ld % 8 == 0 && dx_ld % 8 == 0, "convT2x2_dgrad: ld must be a multiple of 8");
  const int BH = B * H;
  int tw, th, nb;
  pick_patch(W, BH, 1, 128, &tw, &th, &nb);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.bk = pick_bk(Cout);
  const int sw = p.bk * 2;
  // M-space dims follow the tensor-map coordinate order (qx, W, qy, BH)
  p.dimM[0] = 1; p.dimM[1] = W; p.dimM[2] = 1; p.dimM[3] = BH;
  p.boxM[0] = 1; p.boxM[1] = tw; p.boxM[2] = 1; p.boxM[3] = th;
  p.ntile[0] = 1; p.ntile[1] = ceil_div(W, tw); p.ntile[2] = 1; p.ntile[3] = ceil_div(BH, th);
  p.ostride[1] = dx_ld; p.ostride[3] = 1ll * W * dx_ld;
  p.mstride[1] = mask_ld; p.mstride[3] = 1ll * W * mask_ld;
  p.ntaps = 4;
  for (int q = 0; q < 4; ++q) {
    p.tap_off[q][0] = q & 1;   // qx
    p.tap_off[q][1] = 0;
    p.tap_off[q][2] = q >> 1;  // qy
    p.tap_off[q][3] = 0;
  }
  p.cin = Cout;
  p.n_tile = Cin <= 256 ? Cin : 256;
  ROVR_REQUIRE(Cin % p.n_tile == 0, "convT2x2_dgrad: Cin=%d not divisible by n_tile", Cin);
  p.n_tiles_n = Cin / p.n_tile;
  p.n_total = Cin;
  p.epi_mode = IG_EPI_PLAIN;
  p.bias_mod = 1;
  p.out = static_cast<__nv_bfloat16*>(dx);
  p.mask = static_cast<const __nv_bfloat16*>(mask);
  CUtensorMap tmA, tmB;
  long long dims[5], strides[4];
  subpixel_view(W, BH, Cout, dy_ld, dims, strides);
  const int box[5] = {p.bk, 1, tw, 1, th};
  if (int rc = make_map5(&tmA, dy, dims, strides, box, sw)) return rc;
  if (int rc = make_map2(&tmB, wk_d, Cin, 4ll * Cout, 4ll * Cout, p.n_tile, p.bk, sw)) return rc;
  return launch_igemm(tmA, tmB, p, static_cast<cudaStream_t>(stream), "convT2x2_dgrad");
}

// ------------------------------------------------------------------------------------------------
// plain GEMM
// ------------------------------------------------------------------------------------------------
extern "C" int rovr_gemm_bf16(const void* x, int x_ld, const void* wk, const float* bias,
                              void* y_bf16, float* y_f32, int y_ld, int M, int N, int K, int relu,
                              void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(K % 16 == 0 && N % 16 == 0, "gemm: K=%d N=%d must be multiples of 16", K, N);
  ROVR_REQUIRE(x_ld % 8 == 0, "gemm: x_ld must be a multiple of 8");
  ROVR_REQUIRE((y_bf16 != nullptr) != (y_f32 != nullptr), "gemm: exactly one output must be given");
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.bk = pick_bk(K);
  const int sw = p.bk * 2;
  p.dimM[0] = M; p.dimM[1] = 1; p.dimM[2] = 1; p.dimM[3] = 1;
  p.boxM[0] = 128; p.boxM[1] = 1; p.boxM[2] = 1; p.boxM[3] = 1;
  p.ntile[0] = ceil_div(M, 128); p.ntile[1] = 1; p.ntile[2] = 1; p.ntile[3] = 1;
  p.ostride[0] = y_ld;
  p.ntaps = 1;
  p.cin = K;
  int n_tile = 256;
  while (N % n_tile != 0) n_tile -= 16;
  p.n_tile = n_tile;
  p.n_tiles_n = N / n_tile;
  p.n_total = N;
  p.epi_mode = IG_EPI_PLAIN;
  p.relu = relu;
  p.bias = bias;
  p.bias_mod = N;
  p.out = static_cast<__nv_bfloat16*>(y_bf16);
  p.out_f32 = y_f32;
  CUtensorMap tmA, tmB;
  const long long dims[5] = {K, M, 1, 1, 1};
  const long long strides[4] = {x_ld, 1ll * M * x_ld, 1ll * M * x_ld, 1ll * M * x_ld};
  const int box[5] = {p.bk, 128, 1, 1, 1};
  if (int rc = make_map5(&tmA, x, dims, strides, box, sw)) return rc;
  if (int rc = make_map2(&tmB, wk, N, K, K, p.n_tile, p.bk, sw)) return rc;
  return launch_igemm(tmA, tmB, p, static_cast<cudaStream_t>(stream), "gemm_bf16");
}

// ------------------------------------------------------------------------------------------------
// weight gradients
// ------------------------------------------------------------------------------------------------
struct WgradPlan {
  WgradParams p;
  int blk;
  size_t ws_bytes;
  int grid;
  size_t smem;
};

// m_total: channels of the un-shifted operand A; c_total: channels of the tapped operand B.
static int plan_wgrad(WgradPlan* pl, int m_total, int c_total, int taps, const int box[4],
                      const int ntile[4], int sms) {
  WgradParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  int blk = 64;
  while (blk > 16 && (m_total % blk != 0 || c_total % blk != 0)) blk >>= 1;
  ROVR_REQUIRE(m_total % blk == 0 && c_total % blk == 0, "wgrad: channel counts %d/%d not multiples of 16", m_total, c_total);
  pl->blk = blk;
  p.blk = blk;
  int rows = 1;
  for (int j = 0; j < 4; ++j) {
    p.boxM[j] = box[j];
    p.ntile[j] = ntile[j];
    rows *= box[j];
  }
  p.kp = (rows + 15) / 16 * 16;
  p.m_total = m_total;
  p.m_chunks = ceil_div(m_total, 128);
  p.a_blocks = std::min(128, m_total) / blk;
  p.taps_total = taps;
  p.c_total = c_total;
  // taps per CTA: all of them if the accumulators fit in 512 TMEM columns, else one kernel row /
  // half of the sub-pixels
  int tpg = taps;
  if (taps * std::min(c_total, blk) > 512 || taps * c_total > 512) tpg = (taps == 9) ? 3 : 2;
  int cb = c_total / blk;
  while (tpg * cb * blk > 512) cb = (cb + 1) / 2;
  ROVR_REQUIRE((c_total / blk) % cb == 0, "wgrad: channel blocks %d not divisible by group %d", c_total / blk, cb);
  p.taps_per_group = tpg;
  p.tap_groups = ceil_div(taps, tpg);
  p.c_blocks_per_group = cb;
  p.c_groups = (c_total / blk) / cb;
  p.k_tiles = ntile[0] * ntile[1] * ntile[2] * ntile[3];
  const int groups = p.m_chunks * p.tap_groups * p.c_groups;
  p.n_slices = std::max(1, std::min(p.k_tiles, sms / std::max(1, groups)));
  const int nblk = tpg * cb;
  p.tmem_cols = pow2_at_least(nblk * blk);
  const size_t stage = (128 / blk + nblk) * static_cast<size_t>(p.kp) * blk * 2;
  int stages = static_cast<int>((200 * 1024) / stage);
  stages = std::max(2, std::min(stages, WG_MAX_STAGES));
  p.stages = stages;
  pl->smem = wgrad_smem_bytes(blk, p.kp, nblk, stages);
  ROVR_REQUIRE(pl->smem <= 227 * 1024, "wgrad smem %zu exceeds 227 KB", pl->smem);
  pl->grid = groups * p.n_slices;
  pl->ws_bytes = static_cast<size_t>(p.n_slices) * m_total * taps * c_total * sizeof(float);
  return 0;
}

static void conv3_patch_for_wgrad(int B, int H, int W, int box[4], int ntile[4]) {
  int tw, th, nb;
  pick_patch(W, H, B, 64, &tw, &th, &nb);
  box[0] = tw; box[1] = th; box[2] = nb; box[3] = 1;
  ntile[0] = ceil_div(W, tw); ntile[1] = ceil_div(H, th); ntile[2] = ceil_div(B, nb); ntile[3] = 1;
}

extern "C" size_t rovr_conv3x3_wgrad_workspace(int B, int H, int W, int Cin, int Cout) {
  if (ensure_device()) return 0;
  int box[4], ntile[4];
  conv3_patch_for_wgrad(B, H, W, box, ntile);
  WgradPlan pl;
  if (plan_wgrad(&pl, Cout, Cin, 9, box, ntile, g_dev.sms)) return 0;
  return pl.ws_bytes;
}

extern "C" int rovr_conv3x3_wgrad(const void* dy, int dy_ld, const void* x, int x_ld, float* dw,
                                  int B, int H, int W, int Cin, int cin_keep, int Cout, void* ws,
                                  size_t ws_bytes, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(dy_ld % 8 == 0 && x_ld % 8 == 0, "conv3x3_wgrad: ld must be a multiple of 8");
  int box[4], ntile[4];
  conv3_patch_for_wgrad(B, H, W, box, ntile);
  WgradPlan pl;
  if (int rc = plan_wgrad(&pl, Cout, Cin, 9, box, ntile, g_dev.sms)) return rc;
  ROVR_REQUIRE(ws_bytes >= pl.ws_bytes, "conv3x3_wgrad: workspace %zu < %zu", ws_bytes, pl.ws_bytes);
  WgradParams& p = pl.p;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      p.tap_off[r * 3 + s][0] = s - 1;
      p.tap_off[r * 3 + s][1] = r - 1;
    }
  p.partial = static_cast<float*>(ws);
  const int sw = pl.blk * 2;
  CUtensorMap tmA, tmB;
  const long long dimsA[5] = {Cout, W, H, B, 1};
  const long long stA[4] = {dy_ld, 1ll * W * dy_ld, 1ll * H * W * dy_ld, 1ll * B * H * W * dy_ld};
  const long long dimsB[5] = {Cin, W, H, B, 1};
  const long long stB[4] = {x_ld, 1ll * W * x_ld, 1ll * H * W * x_ld, 1ll * B * H * W * x_ld};
  const int bx[5] = {pl.blk, box[0], box[1], box[2], 1};
  if (int rc = make_map5(&tmA, dy, dimsA, stA, bx, sw)) return rc;
  if (int rc = make_map5(&tmB, x, dimsB, stB, bx, sw)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  wgrad_kernel<<<pl.grid, WG_THREADS, pl.smem, st>>>(tmA, tmB, p);
  if (int rc = launch_check("conv3x3_wgrad")) return rc;
  // partial [slice][co][t][ci] -> dw[co][ci_keep][t]
  const long long n = 1ll * Cout * 9 * Cin;
  wgrad_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      p.partial, dw, p.n_slices, Cout, 9, Cin, cin_keep, 9ll * cin_keep, 1, 9, 0);
  return launch_check("wgrad_reduce");
}

static void convT_patch_for_wgrad(int BH, int W, int box[4], int ntile[4]) {
  int tw, th, nb;
  pick_patch(W, BH, 1, 64, &tw, &th, &nb);
  box[0] = 1; box[1] = tw; box[2] = 1; box[3] = th;
  ntile[0] = 1; ntile[1] = ceil_div(W, tw); ntile[2] = 1; ntile[3] = ceil_div(BH, th);
}

extern "C" size_t rovr_convT2x2_wgrad_workspace(int B, int H, int W, int Cin, int Cout) {
  if (ensure_device()) return 0;
  int box[4], ntile[4];
  convT_patch_for_wgrad(B * H, W, box, ntile);
  WgradPlan pl;
  if (plan_wgrad(&pl, Cin, Cout, 4, box, ntile, g_dev.sms)) return 0;
  return pl.ws_bytes;
}

extern "C" int rovr_convT2x2_wgrad(const void* dy, int dy_ld, const void* x, int x_ld, float* dw,
                                   int B, int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes,
                                   void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(dy_ld % 8 == 0 && x_ld % 8 == 0, "convT2x2_wgrad: ld must be a multiple of 8");
  const int BH = B * H;
  int box[4], ntile[4];
  convT_patch_for_wgrad(BH, W, box, ntile);
  WgradPlan pl;
  if (int rc = plan_wgrad(&pl, Cin, Cout, 4, box, ntile, g_dev.sms)) return rc;
  ROVR_REQUIRE(ws_bytes >= pl.ws_bytes, "convT2x2_wgrad: workspace %zu < %zu", ws_bytes, pl.ws_bytes);
  WgradParams& p = pl.p;
  for (int q = 0; q < 4; ++q) {
    p.tap_off[q][0] = q & 1;
    p.tap_off[q][2] = q >> 1;
  }
  p.partial = static_cast<float*>(ws);
  const int sw = pl.blk * 2;
  CUtensorMap tmA, tmB;
  // A = x viewed as (Cin, 1, W, 1, BH); B = dy viewed as (Cout, qx, W, qy, BH)
  const long long dimsA[5] = {Cin, 1, W, 1, BH};
  const long long stA[4] = {x_ld, x_ld, 1ll * W * x_ld, 1ll * W * x_ld};
  long long dimsB[5], stB[4];
  subpixel_view(W, BH, Cout, dy_ld, dimsB, stB);
  const int bx[5] = {pl.blk, 1, box[1], 1, box[3]};
  if (int rc = make_map5(&tmA, x, dimsA, stA, bx, sw)) return rc;
  if (int rc = make_map5(&tmB, dy, dimsB, stB, bx, sw)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  wgrad_kernel<<<pl.grid, WG_THREADS, pl.smem, st>>>(tmA, tmB, p);
  if (int rc = launch_check("convT2x2_wgrad")) return rc;
  // partial [slice][ci][q][co] -> dw[ci][co][q]
  const long long n = 1ll * Cin * 4 * Cout;
  wgrad_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      p.partial, dw, p.n_slices, Cin, 4, Cout, Cout, 4ll * Cout, 1, 4, 0);
  return launch_check("wgrad_reduce");
}

// ------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------
extern "C" int rovr_maxpool_fwd(const void* x, int x_ld, void* y, int y_ld, int B, int H, int W,
                                int C, int kh, int kw, int sh, int sw, void* stream) {
  if (int rc = ensure_device()) return rc;
  ROVR_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0, "maxpool_fwd: C and ld must be multiples of 8");
  const int Ho = (H - kh) / sh + 1, Wo = (W - kw) / sw + 1;
  const long long n = 1ll * B * Ho * Wo * (C / 8);
  maxpool_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<__nv_bfloat16*>(y), y_ld, B, H, W, C,
      kh, kw, sh, sw, Ho, Wo);
  return launch_check("maxpool_fwd");
}
extern "C" int rovr_maxpool_bwd(const void* x, int x_ld, const void* gp, int gp_ld,
                                const void* gskip, int gs_ld, void* gx, int gx_ld, int B, int H,
                                int W, int C, int kh, int kw, int sh, int sw, int relu_mask,
                                void* stream) {
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tiled = (kh == sh && kw == sw && H % kh == 0 && W % kw == 0 && C % 8 == 0 &&
                      x_ld % 8 == 0 && gp_ld % 8 == 0 && gx_ld % 8 == 0 && (gskip == nullptr || gs_ld % 8 == 0));
  if (tiled) {
    const long long n = 1ll * B * (H / kh) * (W / kw) * (C / 8);
    maxpool_bwd_tiled_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<const __nv_bfloat16*>(gp), gp_ld,
        static_cast<const __nv_bfloat16*>(gskip), gs_ld, static_cast<__nv_bfloat16*>(gx), gx_ld, B,
        H, W, C, kh, kw, relu_mask);
    return launch_check("maxpool_bwd_tiled");
  }
  ROVR_REQUIRE(gskip == nullptr, "maxpool_bwd: skip gradient only supported for non-overlapping windows");
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
static int tail_bwd_grid(long long npix) {
  const long long per_block = 1ll * TAILB_THREADS * TAILB_PIX_PER_THREAD;
  return static_cast<int>((npix + per_block - 1) / per_block);
}
extern "C" size_t rovr_tail_workspace(int B, int H, int W) {
  const long long npix = 1ll * B * H * W;
  const size_t fwd = static_cast<size_t>((npix + 255) / 256) * sizeof(float);
  const size_t bwd = static_cast<size_t>(tail_bwd_grid(npix)) * (3 * TAIL_C + 3) * sizeof(float);
  return std::max(fwd, bwd) + 256;
}
extern "C" int rovr_tail_fwd(const void* y7, const float* w8, const float* b8, float* out,
                             const float* target, float* loss, void* ws, size_t ws_bytes, int B,
                             int H, int W, void* stream) {
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long npix = 1ll * B * H * W;
  const int grid = static_cast<int>((npix + 255) / 256);
  float* partial = nullptr;
  if (target) {
    ROVR_REQUIRE(loss != nullptr && ws != nullptr && ws_bytes >= grid * sizeof(float),
                 "tail_fwd: loss requested but loss/ws missing or too small");
    partial = static_cast<float*>(ws);
  }
  tail_fwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(y7), w8, b8, out, target,
                                        partial, B, H * W);
  if (int rc = launch_check("tail_fwd")) return rc;
  if (target) {
    sum_partials_kernel<<<1, 1024, 0, st>>>(partial, grid, 1.f / (3.f * static_cast<float>(npix)), loss);
    return launch_check("sum_partials");
  }
  return 0;
}
extern "C" int rovr_tail_bwd(const void* y7, const float* w8, const float* out, const float* gout,
                             const float* target, float mse_scale, const float* gloss, void* g7,
                             float* dw8, float* db8, void* ws, size_t ws_bytes, int B, int H, int W,
                             void* stream) {
  if (int rc = ensure_device()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long npix = 1ll * B * H * W;
  const int grid = tail_bwd_grid(npix);
  const int ncols = 3 * TAIL_C + 3;
  ROVR_REQUIRE(ws_bytes >= static_cast<size_t>(grid) * ncols * sizeof(float), "tail_bwd: workspace too small");
  ROVR_REQUIRE(gout != nullptr || target != nullptr, "tail_bwd: give gout and/or target");
  float* partial = static_cast<float*>(ws);
  tail_bwd_kernel<<<grid, TAILB_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(y7), w8, out, gout,
                                                  target, mse_scale, gloss,
                                                  static_cast<__nv_bfloat16*>(g7), partial, B, H * W);
  if (int rc = launch_check("tail_bwd")) return rc;
  // columns [0, 192) -> dw8, [192, 195) -> db8
  reduce_rows_kernel<<<1, 256, 0, st>>>(partial, grid, ncols, 3 * TAIL_C, dw8, 0);
  if (int rc = launch_check("reduce_rows(dw8)")) return rc;
  // dw8 buffer is [3][64] contiguous; db8 is separate: run a second tiny reduce on the tail columns
  reduce_rows_kernel<<<1, 32, 0, st>>>(partial + 3 * TAIL_C, grid, ncols, 3, db8, 0);
  return launch_check("reduce_rows(db8)");
}

// ------------------------------------------------------------------------------------------------
// bias gradient
// ------------------------------------------------------------------------------------------------
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
  reduce_rows_kernel<<<(C + 255) / 256, 256, 0, st>>>(static_cast<float*>(ws), grid, C, C, out, 0);
  return launch_check("reduce_rows(colsum)");
}
