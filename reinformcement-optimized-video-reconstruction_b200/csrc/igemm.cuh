// igemm.cuh — the tensor-core engine of the ROVR hot path on B200.
//
// One persistent, warp-specialised kernel computes
//
//     D[m][n] = sum_{t < ntaps} sum_{c < cin}  A_t[m][c] * Wk[n][t*cin + c]       (fp32 accumulate)
//
// where m runs over output pixels, A_t is the NHWC bf16 activation tensor read at the pixel
// shifted by tap t, and Wk is a bf16 weight matrix packed K-major per output channel. With the
// right tap table this is
//   * Conv2d 3x3 pad 1 forward          (reference rovr/local_net.py:12-36,52-68)      9 taps
//   * its data gradient (dgrad)         (autograd of the same lines)                   9 taps, flipped
//   * ConvTranspose2d k2 s2 forward     (rovr/local_net.py:24,29,34,58,62,66)          1 tap, N = 4*Cout,
//                                       pixel-shuffle store into the concat buffer slice
//   * its data gradient                 4 taps gathered through a 5-D strided view of dy
//   * 1x1 convolutions / plain GEMMs    1 tap
//
// Data movement: activations are fetched by TMA (cp.async.bulk.tensor.5d) as a [rows<=128][bk]
// box of a (C, d1, d2, d3, d4) tensor map; out-of-bounds coordinates (the 3x3 halo, ragged edge
// tiles, padded batch) are zero-filled by the TMA unit, which is exactly the conv zero padding.
// Both operands land in shared memory in the 128/64/32-byte swizzled K-major layout that
// tcgen05.mma reads directly. Accumulators live in TMEM (2 stages x n_tile fp32 columns) so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/ReLU/mask -> bf16 -> global).
#pragma once
#include "ptx.cuh"

namespace rovr {

constexpr int IG_MAX_STAGES = 8;
constexpr int IG_MAX_TAPS = 9;
constexpr int IG_THREADS = 192;

enum : int { IG_EPI_PLAIN = 0, IG_EPI_PIXSHUF = 1 };

struct IgemmParams {
  int dimM[4];            // logical output extent along each tiled dim (for edge masking)
  int boxM[4];            // tile extent along each dim; rows per tile = prod(boxM) <= 128
  int ntile[4];           // tiles per dim
  long long ostride[4];   // output element stride per dim
  long long mstride[4];   // mask element stride per dim
  int ntaps;
  int tap_off[IG_MAX_TAPS][4];
  int cin;                // K extent per tap (multiple of bk)
  int bk;                 // K elements per pipeline stage: 16 / 32 / 64  (swizzle 32/64/128 B)
  int n_tile;             // N per tile, multiple of 16, <= 256
  int n_tiles_n;
  int n_total;            // valid N (multiple of 16)
  int stages;
  int tmem_cols;          // power of two >= 2*n_tile, >= 32
  int epi_mode;
  int relu;
  int bias_mod;           // bias index = n % bias_mod
  int shuf_cout;          // pixel-shuffle: channels per quadrant
  long long shuf_sy, shuf_sx;
  const float* bias;      // may be null
  __nv_bfloat16* out;
  const __nv_bfloat16* mask;  // may be null: out *= (mask > 0)
  float* out_f32;         // optional fp32 output instead of bf16 (plain mode only)
};

__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is what SWIZZLE_128B tiles need.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int sw = p.bk * 2;  // bytes per operand row == swizzle span
  int rows = p.boxM[0] * p.boxM[1] * p.boxM[2] * p.boxM[3];
  const uint32_t a_bytes = 128u * sw;  // always reserve the full 128-row tile
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_tile) * sw;
  const uint32_t stage_bytes = a_bytes + b_bytes;  // multiples of 1024 when sw=128
  const uint32_t tx_bytes = static_cast<uint32_t>(rows) * sw + b_bytes;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + IG_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + IG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);

  const int m_tiles = p.ntile[0] * p.ntile[1] * p.ntile[2] * p.ntile[3];
  const int total_tiles = m_tiles * p.n_tiles_n;
  const int kchunks = p.cin / p.bk;
  const int k_iters = p.ntaps * kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  for (int i = threadIdx.x; i < p.n_total; i += IG_THREADS)
    sbias[i] = p.bias ? p.bias[i % p.bias_mod] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        int org[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          org[j] = (mt % p.ntile[j]) * p.boxM[j];
          mt /= p.ntile[j];
        }
        for (int t = 0; t < p.ntaps; ++t) {
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[s], ph ^ 1u, 0x100u + s);
            uint8_t* a_dst = smem + static_cast<size_t>(s) * stage_bytes;
            uint8_t* b_dst = a_dst + a_bytes;
            mbar_expect_tx(&full_bar[s], tx_bytes);
            tma_load_5d(&tmA, &full_bar[s], a_dst, kc * p.bk, org[0] + p.tap_off[t][0],
                        org[1] + p.tap_off[t][1], org[2] + p.tap_off[t][2],
                        org[3] + p.tap_off[t][3]);
            tma_load_2d(&tmB, &full_bar[s], b_dst, t * p.cin + kc * p.bk, nt * p.n_tile);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile, 0, 0);
      const uint32_t sbo = 8u * sw;
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], aph ^ 1u, 0x200u + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.n_tile);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[s], ph, 0x300u + s);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
          const int ksteps = p.bk >> 4;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = umma_smem_desc(a_addr + k * 32u, 0u, sbo, sw);
            const uint64_t bd = umma_smem_desc(b_addr + k * 32u, 0u, sbo, sw);
            umma_bf16(d_tmem, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; aph ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int m = quarter * 32 + lane;
    int acc = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      bool valid = m < rows;
      long long obase = 0, mbase = 0;
      {
        int mm = m;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int org = (mt % p.ntile[j]) * p.boxM[j];
          mt /= p.ntile[j];
          const int pj = org + (mm % p.boxM[j]);
          mm /= p.boxM[j];
          valid = valid && (pj < p.dimM[j]);
          obase += static_cast<long long>(pj) * p.ostride[j];
          mbase += static_cast<long long>(pj) * p.mstride[j];
        }
      }
      mbar_wait(&tfull_bar[acc], aph, 0x400u + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * p.n_tile);
      const int nchunks = p.n_tile >> 4;
      for (int c = 0; c < nchunks; ++c) {
        const int n0 = nt * p.n_tile + c * 16;
        if (n0 >= p.n_total) break;  // uniform across the CTA
        uint32_t v[16];
        tmem_ld16(t_row + static_cast<uint32_t>(c * 16), v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            f[j] = __uint_as_float(v[j]) + sbias[n0 + j];
            if (p.relu) f[j] = fmaxf(f[j], 0.f);
          }
          long long off;
          if (p.epi_mode == IG_EPI_PIXSHUF) {
            const int q = n0 / p.shuf_cout;
            off = obase + (q >> 1) * p.shuf_sy + (q & 1) * p.shuf_sx + (n0 - q * p.shuf_cout);
          } else {
            off = obase + n0;
          }
          if (p.mask) {
            const uint4* mp = reinterpret_cast<const uint4*>(p.mask + mbase + n0);
            const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (!(bf16_lo(mw[j]) > 0.f)) f[2 * j] = 0.f;
              if (!(bf16_hi(mw[j]) > 0.f)) f[2 * j + 1] = 0.f;
            }
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]);
            o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]);
            o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]);
            o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]);
            o1.w = pack_bf16x2(f[14], f[15]);
            uint4* op = reinterpret_cast<uint4*>(p.out + off);
            op[0] = o0;
            op[1] = o1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; aph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// Shared memory the kernel needs for a given configuration (host side).
inline size_t igemm_smem_bytes(int bk, int n_tile, int stages, int n_total) {
  const size_t sw = static_cast<size_t>(bk) * 2;
  const size_t stage = 128 * sw + static_cast<size_t>(n_tile) * sw;
  return 1024 + stages * stage + (2 * IG_MAX_STAGES + 4) * 8 + 16 + static_cast<size_t>(n_total) * 4 + 64;
}

}  // namespace rovr
