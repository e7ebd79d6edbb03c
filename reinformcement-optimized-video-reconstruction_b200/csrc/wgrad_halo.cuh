// wgrad_halo.cuh — Conv2d 3x3 weight gradient with the tapped operand loaded ONCE per pixel patch.
//
//     P[slice][m][t][c] = sum_{pixels p in slice}  A[p][m] * B[p + off_t][c]
//
// Same contraction as wgrad.cuh, but the K step is an 8 x 8 pixel patch: A is the plain 64-row
// tile of the un-shifted tensor and B is fetched as its (8+2) x (8+2) halo (100 rows per channel
// block). Tap (dy, dx) is then just an MN-major *view* of that halo tile — start (dy*10 + dx) rows
// in, 8-row K groups (= image rows of the patch) 10 rows apart — so the nine shifted copies that
// wgrad.cuh pulls through the L2 -> SM fabric collapse into one 100-row load (the same trick as
// igemm's halo mode, on the other operand layout; valid because the 128/64/32-byte swizzle is a
// function of the shared-memory address only).
//
// A CTA owns (a 128-row chunk of m) x (a group of taps) x (a chunk of c) and a slice of the pixel
// patches; per patch it issues taps x 4 MMAs of N = c-chunk columns. Partials are reduced in a
// fixed order by wgrad_reduce_kernel.
#pragma once
#include <type_traits>

#include "ptx.cuh"
#include "wgrad.cuh"

namespace rovr {

constexpr int WH_PATCH = 8;                    // patch is 8 x 8 pixels
constexpr int WH_HALO = WH_PATCH + 2;          // 10
constexpr int WH_HALO_ROWS = WH_HALO * WH_HALO;  // 100
__host__ __device__ inline uint32_t wh_halo_slot(int sw) { return (WH_HALO_ROWS * sw + 1023u) / 1024u * 1024u; }

struct WgradHaloParams {
  int ntile[3];            // patches along W, H, B
  int blk_a, blk_b;        // channels per block of A / B (16 / 32 / 64)
  int a_blocks;            // A blocks loaded (M = 128 rows always addressable)
  int m_chunks, m_total;
  int taps_first[3], taps_count[3], tap_groups;  // taps handled by tap group g
  int tap_sign;            // +1: B read at p + (dy, dx);  -1: at p - (dy, dx)   (operand swap)
  int c_total, c_blocks_per_group, c_groups;
  int n_slices, k_tiles, stages, tmem_cols;   // n_slices = max over tap groups (slice stride of `partial`)
  int slices_g[3], slice_base[3], slices_total;   // pixel slices per tap group (groups with more taps per CTA
                                                  // get more CTAs), first CTA index of each group, their sum
  float* partial;          // [n_slices][m_total][9][c_total]
};

// four K=16 MMAs of one tap view: descriptor low words advance by a_step / b_step per K step
__device__ __forceinline__ void umma_mn_x4(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t a_step,
                                           uint32_t b_lo, uint32_t b_hi, uint32_t b_step,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 a0, b0, a1, b1, a2, b2, a3, b3;\n\t"
      ".reg .b32 l1, l2, l3, m1, m2, m3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %8, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
      "add.u32 l1, %1, %3;\n\tadd.u32 l2, l1, %3;\n\tadd.u32 l3, l2, %3;\n\t"
      "add.u32 m1, %4, %6;\n\tadd.u32 m2, m1, %6;\n\tadd.u32 m3, m2, %6;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%4, %5};\n\t"
      "mov.b64 a1, {l1, %2};\n\tmov.b64 b1, {m1, %5};\n\t"
      "mov.b64 a2, {l2, %2};\n\tmov.b64 b2, {m2, %5};\n\t"
      "mov.b64 a3, {l3, %2};\n\tmov.b64 b3, {m3, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a0, b0, %7, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %7, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %7, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %7, pt;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(a_step), "r"(b_lo), "r"(b_hi), "r"(b_step), "r"(idesc),
      "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradHaloParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int sw_a = p.blk_a * 2, sw_b = p.blk_b * 2;
  const uint32_t ablk_bytes = 64u * sw_a;                       // 64 rows per A block
  const uint32_t a_bytes = static_cast<uint32_t>(128 / p.blk_a) * ablk_bytes;  // 16 KB
  const uint32_t bslot = wh_halo_slot(sw_b);
  const uint32_t stage_bytes = a_bytes + static_cast<uint32_t>(p.c_blocks_per_group) * bslot;
  const uint32_t tx_bytes = static_cast<uint32_t>(p.a_blocks) * ablk_bytes +
                            static_cast<uint32_t>(p.c_blocks_per_group) * WH_HALO_ROWS * sw_b;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* done_bar = empty_bar + WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  int g = blockIdx.x;
  const int sidx = g % p.slices_total;
  g /= p.slices_total;
  const int cg = g % p.c_groups;
  g /= p.c_groups;
  const int mc = g;
  int tg = 0;
  while (tg + 1 < p.tap_groups && sidx >= p.slice_base[tg + 1]) ++tg;
  const int slice = sidx - p.slice_base[tg];
  const int nsl = p.slices_g[tg];
  const int kt0 = static_cast<int>((static_cast<long long>(p.k_tiles) * slice) / nsl);
  const int kt1 = static_cast<int>((static_cast<long long>(p.k_tiles) * (slice + 1)) / nsl);
  const int t_first = p.taps_first[tg], t_count = p.taps_count[tg];
  const int ncol = p.c_blocks_per_group * p.blk_b;  // columns per tap
  // whole kernel rows of a single-block operand: the three dx taps of a row run as one MMA (see below)
  const bool merge3 = p.c_blocks_per_group == 1 && (t_first % 3) == 0 && (t_count % 3) == 0 && 3 * ncol <= 256;

  // A blocks that are never loaded (m_total < 128) must read as zero
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = static_cast<int>((static_cast<size_t>(p.stages) * stage_bytes) >> 4);
    for (int i = threadIdx.x; i < n16; i += WG_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer: per patch, A (64 rows x a_blocks) and the B halo (100 rows x c blocks) ----
    int s = 0;
    uint32_t ph = 0;
    for (int kt = kt0; kt < kt1; ++kt) {
      int r = kt;
      const int x0 = (r % p.ntile[0]) * WH_PATCH;
      r /= p.ntile[0];
      const int y0 = (r % p.ntile[1]) * WH_PATCH;
      const int b0 = r / p.ntile[1];
      mbar_wait(&empty_bar[s], ph ^ 1u, 0x500u + s);
      uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
      if (elect_one_sync()) {
        mbar_expect_tx(&full_bar[s], tx_bytes);
        for (int i = 0; i < p.a_blocks; ++i)
          tma_load_5d(&tmA, &full_bar[s], st + static_cast<size_t>(i) * ablk_bytes,
                      mc * 128 + i * p.blk_a, x0, y0, b0, 0);
        for (int j = 0; j < p.c_blocks_per_group; ++j)
          tma_load_5d(&tmB, &full_bar[s], st + a_bytes + static_cast<size_t>(j) * bslot,
                      (cg * p.c_blocks_per_group + j) * p.blk_b, x0 - 1, y0 - 1, b0, 0);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    const uint32_t a_hi = static_cast<uint32_t>(umma_smem_desc(0u, ablk_bytes, 8u * sw_a, sw_a) >> 32);
    const uint32_t a_lbo = (ablk_bytes >> 4) << 16;  // LBO lives in bits [16,30) of the low word
    const uint32_t b_hi =
        static_cast<uint32_t>(umma_smem_desc(0u, bslot, static_cast<uint32_t>(WH_HALO * sw_b), sw_b) >> 32);
    const uint32_t b_lbo = (bslot >> 4) << 16;
    const uint32_t idesc = umma_idesc_bf16(128, ncol, 1, 1);
    const uint32_t a_step = static_cast<uint32_t>(16 * sw_a) >> 4;             // 16 rows
    const uint32_t b_step = static_cast<uint32_t>(2 * WH_HALO * sw_b) >> 4;    // 2 image rows
    const uint32_t row16 = static_cast<uint32_t>(sw_b) >> 4;
    const uint32_t idesc3 = umma_idesc_bf16(128, 3 * ncol, 1, 1);
    const uint32_t b_lbo_row = row16 << 16;   // LBO = one halo pixel row
    // The tap loop is unrolled at compile time (TC taps per CTA) with the per-tap operand offsets
    // precomputed: with 16-channel operands an MMA retires faster than a generic loop iteration
    // issues, and this single warp's instruction stream is what bounds the kernel.
    auto run = [&](auto tc_tag) {
      constexpr int TC = decltype(tc_tag)::value;
      uint32_t tap_b[TC];
#pragma unroll
      for (int tl = 0; tl < TC; ++tl) {
        const int t = t_first + tl;
        const int dy = p.tap_sign * (t / 3 - 1) + 1, dx = p.tap_sign * (t % 3 - 1) + 1;  // halo-relative
        tap_b[tl] = b_lbo | (static_cast<uint32_t>(dy * WH_HALO + dx) * row16);
      }
      int s = 0;
      uint32_t ph = 0;
      uint32_t accum = 0;
      const uint32_t stage16 = stage_bytes >> 4;
      const uint32_t base16 = smem_u32(smem) >> 4;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(&full_bar[s], ph, 0x600u + s);
        __syncwarp();
        tc_fence_after();
        const uint32_t a16 = base16 + static_cast<uint32_t>(s) * stage16;
        const uint32_t b16 = a16 + (a_bytes >> 4);
        if (TC % 3 == 0 && merge3) {
          // the three dx taps of one kernel row are the same halo view shifted by one pixel row each:
          // with the leading-dimension offset set to ONE ROW they become three N blocks of a single
          // MMA (N = 3 * ncol) — a third of the MMA instructions for the thin (16 / 32-channel) operands
#pragma unroll
          // (operand swap reads the halo at p - (dy, dx): the tap with dx = 0 is then the LAST of its
          // row, and the N blocks come out in reversed tap order — the epilogue undoes that)
          for (int r = 0; r < TC / 3; ++r)
            umma_mn_x4(tmem_base + static_cast<uint32_t>(3 * r * ncol), a_lbo | a16, a_hi, a_step,
                       (tap_b[p.tap_sign == 1 ? 3 * r : 3 * r + 2] & 0xFFFFu) + b_lbo_row + b16, b_hi, b_step, idesc3,
                       accum);
        } else {
#pragma unroll
          for (int tl = 0; tl < TC; ++tl)
            umma_mn_x4(tmem_base + static_cast<uint32_t>(tl * ncol), a_lbo | a16, a_hi, a_step, tap_b[tl] + b16, b_hi,
                       b_step, idesc, accum);
        }
        accum = 1u;
        umma_commit_elect(&empty_bar[s]);
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    };
    switch (t_count) {
      case 9: run(std::integral_constant<int, 9>{}); break;
      case 6: run(std::integral_constant<int, 6>{}); break;
      case 5: run(std::integral_constant<int, 5>{}); break;
      case 4: run(std::integral_constant<int, 4>{}); break;
      default: run(std::integral_constant<int, 3>{}); break;
    }
    umma_commit_elect(done_bar);
    __syncwarp();
  } else {
    // ---- epilogue: TMEM -> fp32 partials ----
    const int quarter = warp & 3;
    const int m = mc * 128 + quarter * 32 + lane;
    const bool have_work = kt1 > kt0;
    if (have_work) {
      mbar_wait(done_bar, 0u, 0x700u);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int n_cols = t_count * ncol;
    for (int c = 0; c < (n_cols >> 4); ++c) {
      uint32_t v[16];
      if (have_work) {
        tmem_ld16(t_row + static_cast<uint32_t>(c * 16), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      const int col = c * 16;
      int tl = col / ncol;
      const int cc = cg * ncol + (col - tl * ncol);
      if (merge3 && p.tap_sign != 1) tl = 3 * (tl / 3) + (2 - tl % 3);
      const int t = t_first + tl;
      if (m < p.m_total && cc < p.c_total) {
        float4* dst = reinterpret_cast<float4*>(
            p.partial + ((static_cast<size_t>(slice) * p.m_total + m) * 9 + t) * p.c_total + cc);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

inline size_t wgrad_halo_stage_bytes(int blk_b, int c_blocks) {
  return 16384 + static_cast<size_t>(c_blocks) * wh_halo_slot(blk_b * 2);
}
inline size_t wgrad_halo_smem_bytes(int blk_b, int c_blocks, int stages) {
  return 1024 + stages * wgrad_halo_stage_bytes(blk_b, c_blocks) + (2 * WG_MAX_STAGES + 1) * 8 + 16 + 64;
}

}  // namespace rovr
