// wgrad.cuh — weight-gradient contractions on tcgen05 with MN-major operands.
//
//     P[slice][m][t][c] = sum_{pixels p in slice}  A[p][m] * B_t[p][c]
//
// A is an un-shifted NHWC bf16 tensor, B_t the other tensor read at tap offset t. For Conv2d 3x3
// either (A = dy, B_t = x shifted by the tap) or, when Cout is narrower than the 128-row MMA,
// (A = x, B_t = dy shifted by minus the tap) so that all 128 rows do useful work; for
// ConvTranspose2d A = x and B_t = dy at the 2x2 sub-pixel t. The reduction index is the pixel,
// which in NHWC is the *strided* dimension, so both operands are fed to the tensor core as
// MN-major tiles: a TMA box of [kp pixels][blk channels] lands in shared memory as kp rows of
// `2*blk` bytes and is described to tcgen05.mma with the MN-major flag — no transposed copy of any
// activation is ever written to HBM. A and B choose their channel-block width (16/32/64)
// independently, so a 16-channel operand does not force 16-channel boxes on the other side.
//
// Work split: a CTA owns one accumulator group (128 rows of m, a tap group, a channel chunk of
// c — at most 512 TMEM columns) and a contiguous slice of pixel tiles (split-K). Partial sums go
// to an fp32 workspace and are combined in a fixed order by wgrad_reduce_kernel, so the gradient
// is bitwise reproducible run to run.
//
// Reference semantics: autograd of nn.Conv2d / nn.ConvTranspose2d weights,
// rovr/local_net.py:12-39 (layers) as driven by rovr/train_local_net_unet.py:115 (backward()).
#pragma once
#include "ptx.cuh"

namespace rovr {

constexpr int WG_MAX_STAGES = 6;
constexpr int WG_THREADS = 192;

struct WgradParams {
  int boxM[4];   // pixel tile extents (rows per K-step = prod <= kp)
  int ntile[4];  // pixel tiles per dim
  int kp;        // rows reserved per block (>= prod(boxM), multiple of 16, <= 128)
  int blk_a;     // channels per A block (16/32/64)
  int blk_b;     // channels per B block (16/32/64)
  int a_blocks;  // A blocks actually loaded per group (M = 128 rows are always addressable)
  int m_chunks;  // number of 128-row chunks of m
  int m_total;   // valid m
  int taps_total;          // T
  int taps_per_group;      // taps handled by one CTA
  int tap_groups;
  int tap_off[9][4];
  int c_total;             // valid c (multiple of blk_b)
  int c_blocks_per_group;  // B channel blocks per tap in one CTA
  int c_groups;
  int n_slices;
  int k_tiles;             // total pixel tiles
  int stages;
  int tmem_cols;
  float* partial;          // [n_slices][m_total][taps_total][c_total]
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const WgradParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int sw_a = p.blk_a * 2, sw_b = p.blk_b * 2;
  const uint32_t ablk_bytes = static_cast<uint32_t>(p.kp) * sw_a;
  const uint32_t bblk_bytes = static_cast<uint32_t>(p.kp) * sw_b;
  const int nblk = p.taps_per_group * p.c_blocks_per_group;
  const uint32_t a_bytes = static_cast<uint32_t>(128 / p.blk_a) * ablk_bytes;  // = 256 * kp
  const uint32_t stage_bytes = a_bytes + static_cast<uint32_t>(nblk) * bblk_bytes;
  const int rows = p.boxM[0] * p.boxM[1] * p.boxM[2] * p.boxM[3];
  const uint32_t tx_bytes = static_cast<uint32_t>(p.a_blocks * sw_a + nblk * sw_b) * rows;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* done_bar = empty_bar + WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  // group / slice decode
  int g = blockIdx.x;
  const int slice = g % p.n_slices;
  g /= p.n_slices;
  const int cg = g % p.c_groups;
  g /= p.c_groups;
  const int tg = g % p.tap_groups;
  g /= p.tap_groups;
  const int mc = g;  // m chunk
  const int kt0 = static_cast<int>((static_cast<long long>(p.k_tiles) * slice) / p.n_slices);
  const int kt1 = static_cast<int>((static_cast<long long>(p.k_tiles) * (slice + 1)) / p.n_slices);
  const int t_first = tg * p.taps_per_group;
  const int n_cols = nblk * p.blk_b;

  // Rows of a block that TMA never writes (kp > rows) and A blocks that are never loaded must
  // read as zero / be harmless: zero the whole staging area once.
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
    int s = 0;
    uint32_t ph = 0;
    for (int kt = kt0; kt < kt1; ++kt) {
      int org[4];
      int r = kt;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        org[j] = (r % p.ntile[j]) * p.boxM[j];
        r /= p.ntile[j];
      }
      mbar_wait(&empty_bar[s], ph ^ 1u, 0x500u + s);
      uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
      if (elect_one_sync()) {
        mbar_expect_tx(&full_bar[s], tx_bytes);
        for (int i = 0; i < p.a_blocks; ++i)
          tma_load_5d(&tmA, &full_bar[s], st + static_cast<size_t>(i) * ablk_bytes,
                      mc * 128 + i * p.blk_a, org[0], org[1], org[2], org[3]);
        for (int tl = 0; tl < p.taps_per_group; ++tl) {
          const int t = t_first + tl;
          for (int j = 0; j < p.c_blocks_per_group; ++j) {
            const int c0 = (cg * p.c_blocks_per_group + j) * p.blk_b;
            tma_load_5d(&tmB, &full_bar[s],
                        st + a_bytes + static_cast<size_t>(tl * p.c_blocks_per_group + j) * bblk_bytes,
                        c0, org[0] + p.tap_off[t][0], org[1] + p.tap_off[t][1],
                        org[2] + p.tap_off[t][2], org[3] + p.tap_off[t][3]);
          }
        }
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // Warp-uniform walk (uniform registers for addresses/descriptors); lane 0 issues.
    // The N columns of a group are covered by at most 3 MMAs per 16-pixel K step.
    const uint64_t dhi_a = umma_smem_desc(0u, ablk_bytes, 8u * sw_a, sw_a);
    const uint64_t dhi_b = umma_smem_desc(0u, bblk_bytes, 8u * sw_b, sw_b);
    const int blocks_per_mma = 256 / p.blk_b < nblk ? 256 / p.blk_b : nblk;
    const int n_mma = (nblk + blocks_per_mma - 1) / blocks_per_mma;  // <= 3 (<= 512 columns)
    const int nb_last = nblk - (n_mma - 1) * blocks_per_mma;
    const uint32_t idesc_full = umma_idesc_bf16(128, blocks_per_mma * p.blk_b, 1, 1);
    const uint32_t idesc_last = umma_idesc_bf16(128, nb_last * p.blk_b, 1, 1);
    const uint32_t kstep_a16 = static_cast<uint32_t>(16 * sw_a) >> 4;  // 16 pixel rows, in 16 B units
    const uint32_t kstep_b16 = static_cast<uint32_t>(16 * sw_b) >> 4;
    const uint32_t mma_b16 = (static_cast<uint32_t>(blocks_per_mma) * bblk_bytes) >> 4;
    const uint32_t mma_cols = static_cast<uint32_t>(blocks_per_mma * p.blk_b);
    const int ksteps = p.kp >> 4;
    int s = 0;
    uint32_t ph = 0;
    uint32_t accum = 0;
    for (int kt = kt0; kt < kt1; ++kt) {
      mbar_wait(&full_bar[s], ph, 0x600u + s);
      tc_fence_after();
      const uint32_t a16 = smem_u32(smem + static_cast<size_t>(s) * stage_bytes) >> 4;
      const uint32_t b16 = a16 + (a_bytes >> 4);
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t ad = dhi_a | static_cast<uint64_t>(a16 + k * kstep_a16);
        const uint32_t bk16 = b16 + k * kstep_b16;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j < n_mma) {
            const uint64_t bd = dhi_b | static_cast<uint64_t>(bk16 + j * mma_b16);
            const uint32_t id = (j == n_mma - 1) ? idesc_last : idesc_full;
            umma_bf16_elect(tmem_base + j * mma_cols, ad, bd, id, accum);
          }
        }
        accum = 1u;
      }
      umma_commit_elect(&empty_bar[s]);
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    umma_commit_elect(done_bar);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int m_local = quarter * 32 + lane;
    const int m = mc * 128 + m_local;
    const bool have_work = kt1 > kt0;
    if (have_work) {
      mbar_wait(done_bar, 0u, 0x700u);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
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
      const int b = col / p.blk_b;               // block within the group
      const int tl = b / p.c_blocks_per_group;   // local tap
      const int cj = b - tl * p.c_blocks_per_group;
      const int cc = (cg * p.c_blocks_per_group + cj) * p.blk_b + (col - b * p.blk_b);
      const int t = t_first + tl;
      if (m < p.m_total && cc < p.c_total && t < p.taps_total) {
        float4* dst = reinterpret_cast<float4*>(
            p.partial +
            ((static_cast<size_t>(slice) * p.m_total + m) * p.taps_total + t) * p.c_total + cc);
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

// final[m*fs_m + t*fs_t + c*fs_c] = sum_slices P[slice][m][t][c], for m < m_keep, c < c_keep.
// One thread per element: right when there are few slices and many elements (the wide layers).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad,
                                    int n_slices, int m_total, int taps, int c_total, int m_keep,
                                    int c_keep, long long fs_m, long long fs_t, long long fs_c,
                                    int accumulate, int tap_split, int n_slices_hi) {
  pdl_trigger();
  pdl_wait();
  const long long n = static_cast<long long>(m_total) * taps * c_total;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % c_total);
  const int t = static_cast<int>((i / c_total) % taps);
  const int m = static_cast<int>(i / (static_cast<long long>(c_total) * taps));
  if (c >= c_keep || m >= m_keep) return;
  if (t >= tap_split) n_slices = n_slices_hi;   // tap groups with their own number of pixel slices (6 + 3 taps)
  // four independent partial sums keep four loads in flight (slices are ~100s of KB apart); the
  // combination order is fixed, so the result stays bitwise reproducible
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = 0;
  for (; k + 3 < n_slices; k += 4) {
    s0 += __ldg(partial + static_cast<long long>(k) * n + i);
    s1 += __ldg(partial + static_cast<long long>(k + 1) * n + i);
    s2 += __ldg(partial + static_cast<long long>(k + 2) * n + i);
    s3 += __ldg(partial + static_cast<long long>(k + 3) * n + i);
  }
  for (; k < n_slices; ++k) s0 += __ldg(partial + static_cast<long long>(k) * n + i);
  const float s = (s0 + s1) + (s2 + s3);
  float* dst = grad + m * fs_m + t * fs_t + c * fs_c;
  *dst = accumulate ? (*dst + s) : s;
}
// Many slices, few elements (thin layers: up to 148 slices of a few thousand weights):
// Block = 32 consecutive elements (x) by NY slice lanes (y): lane y adds slices y, y + NY, ... (up to
// four loads in flight), the NY partial sums meet in shared memory and are added in the order
// y = 0 .. NY-1 — a fixed order, so the gradient is bitwise reproducible. With one thread per element
// (the first version) a thin layer's 148 slices were 37 dependent round trips to L2 on 36 CTAs.
constexpr int WGR_MAX_NY = 32;
__global__ void wgrad_reduce_tall_kernel(const float* __restrict__ partial, float* __restrict__ grad,
                                    int n_slices, int m_total, int taps, int c_total, int m_keep,
                                    int c_keep, long long fs_m, long long fs_t, long long fs_c,
                                    int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[WGR_MAX_NY][33];
  const long long n = static_cast<long long>(m_total) * taps * c_total;
  const long long i = static_cast<long long>(blockIdx.x) * 32 + threadIdx.x;
  const int ny = blockDim.y, y = threadIdx.y;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    int k = y;
    for (; k + 3 * ny < n_slices; k += 4 * ny) {
      s0 += __ldg(partial + static_cast<long long>(k) * n + i);
      s1 += __ldg(partial + static_cast<long long>(k + ny) * n + i);
      s2 += __ldg(partial + static_cast<long long>(k + 2 * ny) * n + i);
      s3 += __ldg(partial + static_cast<long long>(k + 3 * ny) * n + i);
    }
    for (; k < n_slices; k += ny) s0 += __ldg(partial + static_cast<long long>(k) * n + i);
  }
  part[y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (y != 0 || i >= n) return;
  float s = 0.f;
  for (int j = 0; j < ny; ++j) s += part[j][threadIdx.x];
  const int c = static_cast<int>(i % c_total);
  const int t = static_cast<int>((i / c_total) % taps);
  const int m = static_cast<int>(i / (static_cast<long long>(c_total) * taps));
  if (c >= c_keep || m >= m_keep) return;
  float* dst = grad + m * fs_m + t * fs_t + c * fs_c;
  *dst = accumulate ? (*dst + s) : s;
}
// Layout-changing case [slice][m][t][c] -> grad[m][c][t] (Conv2d / ConvTranspose2d weights: taps innermost):
// one block per (m, 32 channels). Thread (t, c) adds its slices with coalesced 128-byte reads, the 32 x T tile
// turns in shared memory, and the block writes 32 * T consecutive floats. One thread per element (above) wrote
// with a stride of T floats — nine partially filled sectors per warp store (conv4: 23 us for 33 MB).
constexpr int WGR_MAX_TAPS = 9;
__global__ void __launch_bounds__(32 * WGR_MAX_TAPS)
wgrad_reduce_tiled_kernel(const float* __restrict__ partial, float* __restrict__ grad, int n_slices, int m_total,
                          int taps, int c_total, int m_keep, int c_keep, long long fs_m, int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][WGR_MAX_TAPS + 1];
  const int m = blockIdx.y, c0 = blockIdx.x * 32;
  const int cl = threadIdx.x & 31, t = threadIdx.x >> 5;
  const long long n = static_cast<long long>(m_total) * taps * c_total;
  if (m >= m_keep) return;
  float s0 = 0.f, s1 = 0.f;
  if (c0 + cl < c_total) {
    const float* src = partial + (static_cast<long long>(m) * taps + t) * c_total + c0 + cl;
    int k = 0;
    for (; k + 1 < n_slices; k += 2) {
      s0 += __ldg(src + static_cast<long long>(k) * n);
      s1 += __ldg(src + static_cast<long long>(k + 1) * n);
    }
    if (k < n_slices) s0 += __ldg(src + static_cast<long long>(k) * n);
  }
  tile[cl][t] = s0 + s1;
  __syncthreads();
  // output element j of this block: channel c0 + j / taps, tap j % taps
  const int j = threadIdx.x;
  const int jc = j / taps, jt = j - jc * taps;
  if (c0 + jc < c_keep) {
    float* dst = grad + m * fs_m + static_cast<long long>(c0) * taps + j;
    const float v = tile[jc][jt];
    *dst = accumulate ? (*dst + v) : v;
  }
}

inline void launch_wgrad_reduce(const float* partial, float* grad, int n_slices, int m_total, int taps, int c_total,
                                int m_keep, int c_keep, long long fs_m, long long fs_t, long long fs_c, int accumulate,
                                cudaStream_t st, int tap_split = 1 << 30, int n_slices_hi = 0) {
  const long long n = static_cast<long long>(m_total) * taps * c_total;
  if (tap_split < taps) {   // per-tap-group slice counts: only the element-per-thread kernel knows them
    launch_chain(wgrad_reduce_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1, partial, grad,
                 n_slices, m_total, taps, c_total, m_keep, c_keep, fs_m, fs_t, fs_c, accumulate, tap_split, n_slices_hi);
    return;
  }
  if (fs_t == 1 && fs_c == taps && taps > 1 && taps <= WGR_MAX_TAPS && n_slices <= 32) {   // (more slices: measured slower)
    launch_chain(wgrad_reduce_tiled_kernel, dim3((c_total + 31) / 32, m_total), dim3(32 * taps), 0, st, 1, partial, grad,
                 n_slices, m_total, taps, c_total, m_keep, c_keep, fs_m, accumulate);
    return;
  }
  // measured: the slice-parallel kernel only wins when the element count is too small to fill the GPU
  if (n_slices >= 24 && n <= 32768) {
    launch_chain(wgrad_reduce_tall_kernel, dim3(static_cast<unsigned>((n + 31) / 32)), dim3(32, WGR_MAX_NY), 0, st, 1,
        partial, grad, n_slices, m_total, taps, c_total, m_keep, c_keep, fs_m, fs_t, fs_c, accumulate);
  } else {
    launch_chain(wgrad_reduce_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1,
        partial, grad, n_slices, m_total, taps, c_total, m_keep, c_keep, fs_m, fs_t, fs_c, accumulate, 1 << 30, 0);
  }
}

inline size_t wgrad_stage_bytes(int kp, int blk_b, int nblk) {
  return 256 * static_cast<size_t>(kp) + static_cast<size_t>(nblk) * kp * blk_b * 2;
}
inline size_t wgrad_smem_bytes(int kp, int blk_b, int nblk, int stages) {
  return 1024 + stages * wgrad_stage_bytes(kp, blk_b, nblk) + (2 * WG_MAX_STAGES + 1) * 8 + 16 + 64;
}

}  // namespace rovr
