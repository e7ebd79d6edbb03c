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
// Data movement
//   loads   : TMA (cp.async.bulk.tensor.5d) fetches a [rows<=128][bk] box of a (C, d1..d4) tensor
//             map per tap; out-of-bounds coordinates (the 3x3 halo, ragged edge tiles, padded
//             batch) are zero-filled by the TMA unit — exactly the conv zero padding. Operands
//             land in the 128/64/32-byte swizzled K-major layout tcgen05.mma reads directly. A
//             pipeline stage holds `tps` (tap, k-chunk) sub-tiles so that thin layers (Cin = 16)
//             still move tens of KB per barrier round trip.
//   compute : accumulators live in TMEM (2 or 4 stages x n_tile fp32 columns): the epilogues of
//             tiles i and i+1 overlap the main loops of the tiles after them.
//   stores  : the epilogue converts 64-channel column blocks to bf16 into a swizzled smem staging
//             tile and a TMA store (cp.async.bulk.tensor...global.shared::cta) writes it out, so
//             global writes are full 128-byte lines and ragged tiles are clipped by the TMA unit.
//             The ReLU mask of dgrad (the forward activation at the same coordinates) is read
//             from global memory by the thread that owns the row, one column block ahead.
//   pairs   : optionally two CTAs of a cluster (one TPC) run ONE 256-row cta_group::2 MMA per
//             step, each staging its own 128 pixels and half of the weight tile (kPair).
//
// Warp roles (320 threads): warps 0..7 = two epilogue groups, warp 8 = TMA producer,
// warp 9 = TMEM owner + MMA issuer.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace rovr {

constexpr int IG_MAX_STAGES = 8;
constexpr int IG_MAX_TAPS = 9;
constexpr int IG_THREADS = 320;
constexpr int IG_EPI_THREADS = 128;
// Warp roles. The four epilogue warps must be warps whose id % 4 covers the four TMEM lane
// quarters, so each SM sub-partition hosts one of them; the MMA issuer gets the HIGHEST warp id
// because the sub-partition arbiter favours higher warp ids, and the single-thread MMA issue
// stream is the latency-critical path of the kernel.
constexpr int IG_WARP_TMA = 8, IG_WARP_MMA = 9;
// Halo mode (3x3 convolutions with Cin % 64 == 0): the M tile is an 8 x 16 pixel patch and its
// (8+2) x (16+2) halo is fetched ONCE per 64-channel chunk (180 rows x 128 B). The nine taps are
// nine K-major views of that one tile: tap (dy, dx) starts (dy*10 + dx) rows further on, each
// 8-row core-matrix group is one image row (8 consecutive smem rows) and consecutive groups are
// one halo row apart (SBO = 10 * 128 B). The 128-byte swizzle is a function of the shared-memory
// address, so TMA's write pattern and the tensor core's read pattern agree for any such view.
// This cuts the L2 -> SM traffic of the activation operand 6.4x, which is what bounds an
// implicit-GEMM conv on B200 (the fabric delivers ~42 B/clk/SM; a 128x128 tile wants 128).
constexpr int IG_HALO_TW = 8, IG_HALO_TH = 16;
constexpr int IG_HALO_ROWS = (IG_HALO_TW + 2) * (IG_HALO_TH + 2);     // 180 rows of `sw` bytes
constexpr int IG_MAX_ASLOTS = 4;
__host__ __device__ inline uint32_t ig_halo_bytes(int sw) { return IG_HALO_ROWS * sw; }
__host__ __device__ inline uint32_t ig_halo_slot(int sw) { return (ig_halo_bytes(sw) + 1023u) / 1024u * 1024u; }
// Resident-weight mode (halo convs whose whole packed weight matrix fits in shared memory, e.g.
// conv7 of LocalNet: 64 x 1152 bf16 = 144 KB): every (tap, k-chunk) weight tile is loaded once
// per CTA and the pipeline only streams input patches.

enum : int { IG_EPI_PLAIN = 0, IG_EPI_PIXSHUF = 1 };

struct IgemmParams {
  int dimM[4];            // logical output extent along each tiled dim (direct fp32 epilogue only)
  int boxM[4];            // tile extent along each dim; rows per tile = prod(boxM) <= 128
  int ntile[4];           // tiles per dim
  long long ostride[4];   // output element stride per dim (direct fp32 epilogue only)
  int ntaps;
  int tap_off[IG_MAX_TAPS][4];
  int a_mul[4];           // A-operand coordinate of a tile = org[j] * a_mul[j] + tap_off[t][j] (2 for the pixel dims of a
                          // stride-2 convolution, whose A map samples every other input pixel; 1 otherwise)
  int cin;                // K extent per tap (multiple of bk)
  int bk;                 // K elements per sub-tile: 16 / 32 / 64  (swizzle 32/64/128 B)
  int tps;                // (tap, k-chunk) sub-tiles per pipeline stage
  int n_tile;             // N per tile, multiple of 16, <= 256
  int n_tiles_n;
  int n_total;            // valid N (multiple of cw)
  int stages;
  int tmem_cols;          // power of two >= 2*n_tile, >= 32
  int epi_mode;
  int cw;                 // epilogue column-block width: 16 / 32 / 64 channels
  int relu;
  int pool2;              // 1 (halo mode, plain epilogue): also emit the 2x2 max-pooled tile (4 x 8 pixels) through
                          //    the fourth tensor map — nn.MaxPool2d(2, 2) of the encoder fused into the conv
  // Fused LocalNet tail (conv7 forward only: n_tile == cw == 64): out[b][k][pixel] = sigmoid(b8[k] +
  // sum_c w8[k][c] * y[pixel][c]) for k < 3, computed by the thread that owns the pixel row from the
  // bf16-rounded values it stores, plus (optional) per-(tile, warp) partial sums of (out - target)^2.
  const float* tail_w;    // [3][64] fp32 (Conv2d 1x1 weight), null = no tail
  const float* tail_b;    // [3]
  float* tail_out;        // NCHW fp32 [B][3][H][W]
  const float* tail_target;   // may be null
  float* tail_loss_partial;   // [m_tiles * 4], may be null
  long long tail_plane;   // H * W
  int b_batched;          // 1: the B operand is a rank-5 map (K, N, d2, d3, d4) whose trailing coordinates are
                          //    the tile's M-space origin org[1..3] (batched GEMM: one B matrix per head / batch)
  int halo;               // 1: Conv2d 3x3 with the input patch loaded once per k-chunk (see below)
  int a_slots;            // halo mode: depth of the A (halo tile) ring
  int resident_b;         // halo mode: all weight tiles stay in shared memory for the whole kernel
  unsigned stage_begin_mask;  // halo mode, bit t: tap t is the first sub-tile of a weight stage
  unsigned stage_end_mask;    // halo mode, bit t: tap t is the last sub-tile of a weight stage
  int bias_mod;           // bias index = n % bias_mod
  int shuf_cout;          // pixel-shuffle: channels per quadrant
  const float* bias;      // may be null
  const __nv_bfloat16* mask;  // non-null: out *= (mask > 0); same pixel space as the output
  int mask_cols;          // the mask applies to output columns < mask_cols only (the skip half of a concat
                          // gradient is masked later by the pool backward that consumes it)
  int cs_cols;            // column sums are wanted for columns < cs_cols only
  long long mstride[4];   // mask element stride per tiled dim
  float* out_f32;         // non-null: direct fp32 epilogue (dimM / ostride address the rows; with IG_EPI_PIXSHUF the
                          // quadrant offsets are ostride[0] (qx) and ostride[2] (qy)) instead of the TMA store
  float* colsum_partial;  // non-null: per-(M tile, lane quarter) column sums of the bf16 output,
                          // [m_tiles * 4][n_total] fp32 — the bias gradient of the layer that
                          // consumes this gradient tensor, reduced afterwards in a fixed order
  // division by the run-time tile counts without the ~25-instruction integer-division sequence (ncu: the
  // per-tile index arithmetic was a quarter of the epilogue's instruction stream on thin layers):
  // x / d == (umulhi(x, fd_mul) + x) >> fd_shr for x < 2^31. Index 0..3 = ntile[j], 4 = n_tiles_n,
  // 5 = shuf_cout (filled by igemm_dispatch).
  uint32_t fd_mul[6], fd_shr[6];
  int pair;               // 1: CTA-pair launch (cluster of 2, cta_group::2 MMAs of 256 rows; each CTA stages its own
                          //    128-row A tile and half of the B tile) — halves the weight traffic L2 -> SM
  int acc_stages;         // TMEM accumulator stages: 4 when 4 * n_tile <= 512 columns, else 2
  int stg_bufs;           // staging tiles per epilogue group: 2 when shared memory allows (the TMA store of a
                          // column block drains while the next block is converted), else 1
};

inline void ig_fastdiv_make(uint32_t d, uint32_t* mul, uint32_t* shr) {
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  *mul = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  *shr = l;
}
__device__ __forceinline__ int ig_fastdiv(int x, uint32_t mul, uint32_t shr) {
  return static_cast<int>((__umulhi(static_cast<uint32_t>(x), mul) + static_cast<uint32_t>(x)) >> shr);
}

// ---- TMA store / bulk-group helpers ------------------------------------------------------------
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1,
                                             int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(IG_EPI_THREADS) : "memory");
}
// byte offset of 16-byte chunk `j` of row `m` in a TMA-swizzled tile whose rows are `rowb` bytes
__device__ __forceinline__ uint32_t swz_off(int m, int j, int rowb) {
  const uint32_t a = static_cast<uint32_t>(m) * rowb + static_cast<uint32_t>(j) * 16u;
  const uint32_t msk = (rowb == 128) ? 7u : ((rowb == 64) ? 3u : 1u);
  return a ^ (((a >> 7) & msk) << 4);
}

// kPlainEpi: the epilogue has neither a ReLU mask nor fused column sums (every forward launch):
// their register state disappears and all four 16-column TMEM chunks of a block are fetched at once.
// kTail: the fused LocalNet tail (conv7 forward only); kPool: the fused 2x2 max-pool (encoder convs).
// Each is its own instance so that the plain one keeps its 130-register epilogue — growing it to 141
// (pool) or 167 (tail) registers measurably slows every thin layer.
// kPair: CTA-pair variant (see IgemmParams::pair); launched as clusters of two CTAs.
template <bool kPlainEpi, bool kTail = false, bool kPool = false, bool kPair = false>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmMask,
             const IgemmParams p) {
  pdl_trigger();   // the next launch may start filling SMs as this grid's CTAs retire
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is what SWIZZLE_128B tiles need.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t crank = kPair ? cluster_ctarank() : 0u;   // rank in the CTA pair; 0 = leader (issues the MMAs)
  const int sw = p.bk * 2;  // bytes per operand row == swizzle span
  const int rows = p.boxM[0] * p.boxM[1] * p.boxM[2] * p.boxM[3];
  const bool halo = p.halo != 0;
  const uint32_t a_bytes = halo ? 0u : 128u * sw;  // always reserve the full 128-row tile
  const int nb_rows = kPair ? (p.n_tile >> 1) : p.n_tile;   // B rows staged by this CTA
  const uint32_t b_bytes = static_cast<uint32_t>(nb_rows) * sw;
  const uint32_t sub_bytes = a_bytes + b_bytes;
  const uint32_t stage_bytes = sub_bytes * p.tps;
  const uint32_t sub_tx = (halo ? 0u : static_cast<uint32_t>(rows) * sw) + b_bytes;
  const int epi_rowb = p.cw * 2;
  const uint32_t stg_bytes = 128u * epi_rowb;

  const uint32_t halo_bytes = ig_halo_bytes(sw), halo_slot = ig_halo_slot(sw);
  // [pipeline stages | resident weights] [halo ring] [staging] [mask] [barriers] [bias]
  const size_t front_bytes = p.resident_b ? static_cast<size_t>(p.ntaps) * (p.cin / p.bk) * b_bytes
                                          : static_cast<size_t>(p.stages) * stage_bytes;
  uint8_t* aring = smem + front_bytes;                                           // halo tiles
  aring += (1024u - (smem_u32(aring) & 1023u)) & 1023u;                          // swizzle alignment
  uint8_t* stg_base = aring + (halo ? p.a_slots * halo_slot : 0);                // 2 or 4 staging tiles
  const uint32_t pstg_bytes = p.pool2 ? 32u * epi_rowb : 0u;                     // pooled tile: 32 rows
  uint8_t* pstg_base = stg_base + 2 * p.stg_bufs * stg_bytes;                    // 2 pooled staging tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(pstg_base + 2 * p.stg_bufs * pstg_bytes);
  uint64_t* empty_bar = full_bar + IG_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + IG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 4;
  uint64_t* afull_bar = tempty_bar + 4;
  uint64_t* aempty_bar = afull_bar + IG_MAX_ASLOTS;
  uint64_t* wfull_bar = aempty_bar + IG_MAX_ASLOTS;
  // 33 barriers + one pad = 272 bytes: everything after stays 16-byte aligned by construction (pointer
  // arithmetic only — an integer round trip would make ptxas emit generic instead of shared loads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 2);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);
  float4* stail = reinterpret_cast<float4*>(sbias + ((p.n_total + 3) & ~3));

  const int m_tiles = p.ntile[0] * p.ntile[1] * p.ntile[2] * p.ntile[3];
  const int total_tiles = m_tiles * p.n_tiles_n;
  const int kchunks = p.cin / p.bk;
  const int k_iters = p.ntaps * kchunks;
  const int s_iters = (k_iters + p.tps - 1) / p.tps;

  if (warp == IG_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (kPool) tma_prefetch_desc(&tmMask);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kPair ? 8 : 4);   // the leader's barrier also collects the peer's four warps
    }
    for (int a = 0; a < IG_MAX_ASLOTS; ++a) {
      mbar_init(&afull_bar[a], 1);
      mbar_init(&aempty_bar[a], 1);
    }
    mbar_init(wfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == IG_WARP_MMA) {
    if constexpr (kPair) tmem_alloc_pair(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    else tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  }
  pdl_wait();      // everything above is private to this CTA; from here on global memory is read
  for (int i = threadIdx.x; i < p.n_total; i += IG_THREADS)
    sbias[i] = p.bias ? p.bias[i % p.bias_mod] : 0.f;
  if (kTail && threadIdx.x < 64)   // w8 transposed to [c][k] so that one 16-byte read serves a column
    stail[threadIdx.x] = make_float4(p.tail_w[threadIdx.x], p.tail_w[64 + threadIdx.x], p.tail_w[128 + threadIdx.x], 0.f);
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Work items: a CTA walks tiles blockIdx.x, +gridDim.x, ...; a CTA pair walks PAIRS of M tiles
  // (2 pm + rank, nt) with the same stride in pair units.
  const int it0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int it_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_its = kPair ? (total_tiles >> 1) : total_tiles;
  auto tile_of = [&](int it) {
    if constexpr (!kPair) {
      return it;
    } else {
      const int pm = ig_fastdiv(it, p.fd_mul[4], p.fd_shr[4]);
      return ((pm << 1) + static_cast<int>(crank)) * p.n_tiles_n + (it - pm * p.n_tiles_n);
    }
  };
  // loads: in a pair the bytes of both CTAs are counted on the LEADER's barrier
  auto expect_tx = [&](uint64_t* bar, uint32_t bytes) {
    if constexpr (!kPair) mbar_expect_tx(bar, bytes);
    else if (crank == 0) mbar_expect_tx(bar, 2u * bytes);
  };
  auto load_b2 = [&](uint64_t* bar, void* dst, int c0, int c1) {
    if constexpr (kPair) tma_load_2d_pair(&tmB, leader_addr(bar), dst, c0, c1);
    else tma_load_2d(&tmB, bar, dst, c0, c1);
  };
  auto load_a5 = [&](uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
    if constexpr (kPair) tma_load_5d_pair(&tmA, leader_addr(bar), dst, c0, c1, c2, c3, c4);
    else tma_load_5d(&tmA, bar, dst, c0, c1, c2, c3, c4);
  };
  const int b_row0 = static_cast<int>(crank) * nb_rows;   // this CTA's half of a B tile

  if (warp == IG_WARP_TMA) {
    // ================================ TMA producer ================================
    {
      if (p.resident_b) {  // all weight tiles, once: tile index = t * kchunks + kc
        if (elect_one_sync()) {
          expect_tx(wfull_bar, static_cast<uint32_t>(k_iters) * b_bytes);
          for (int it = 0; it < k_iters; ++it)
            load_b2(wfull_bar, smem + static_cast<size_t>(it) * b_bytes, it * p.bk, b_row0);
        }
        __syncwarp();
      }
      int s = 0, sa = 0;
      uint32_t ph = 0, pha = 0;
      for (int it = it0; it < total_its; it += it_step) {
        const int tile = tile_of(it);
        int mt = ig_fastdiv(tile, p.fd_mul[4], p.fd_shr[4]);
        const int nt = tile - mt * p.n_tiles_n;
        int org[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = ig_fastdiv(mt, p.fd_mul[j], p.fd_shr[j]);
          org[j] = (mt - q * p.ntile[j]) * p.boxM[j];
          mt = q;
        }
        if (halo) {
          // per k-chunk: one halo tile of the input, then (unless resident) the nine weight tiles
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait_warp(&aempty_bar[sa], pha ^ 1u, 0x900u + sa);
            if (elect_one_sync()) {
              expect_tx(&afull_bar[sa], halo_bytes);
              load_a5(&afull_bar[sa], aring + sa * halo_slot, kc * p.bk, org[0] - 1, org[1] - 1, org[2], org[3]);
            }
            __syncwarp();
            if (++sa == p.a_slots) { sa = 0; pha ^= 1u; }
            for (int t0 = 0; t0 < (p.resident_b ? 0 : 9); t0 += p.tps) {
              const int nsub = (9 - t0) < p.tps ? (9 - t0) : p.tps;
              mbar_wait_warp(&empty_bar[s], ph ^ 1u, 0x100u + s);
              uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
              if (elect_one_sync()) {
                expect_tx(&full_bar[s], b_bytes * nsub);
                for (int u = 0; u < nsub; ++u)
                  load_b2(&full_bar[s], st + static_cast<size_t>(u) * b_bytes, (t0 + u) * p.cin + kc * p.bk,
                          nt * p.n_tile + b_row0);
              }
              __syncwarp();
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
          }
          continue;
        }
        for (int si = 0; si < s_iters; ++si) {
          const int it0 = si * p.tps;
          const int nsub = (k_iters - it0) < p.tps ? (k_iters - it0) : p.tps;
          mbar_wait_warp(&empty_bar[s], ph ^ 1u, 0x100u + s);
          uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
          if (elect_one_sync()) expect_tx(&full_bar[s], sub_tx * nsub);
          for (int u = 0; u < nsub; ++u) {
            const int it = it0 + u;
            const int t = it / kchunks;
            const int kc = it - t * kchunks;
            uint8_t* a_dst = st + static_cast<size_t>(u) * sub_bytes;
            if (elect_one_sync()) {
              load_a5(&full_bar[s], a_dst, kc * p.bk, org[0] * p.a_mul[0] + p.tap_off[t][0],
                      org[1] * p.a_mul[1] + p.tap_off[t][1], org[2] * p.a_mul[2] + p.tap_off[t][2],
                      org[3] * p.a_mul[3] + p.tap_off[t][3]);
              if (p.b_batched)   // (never combined with a pair launch)
                tma_load_5d(&tmB, &full_bar[s], a_dst + a_bytes, t * p.cin + kc * p.bk, nt * p.n_tile, org[1],
                            org[2], org[3]);
              else
                load_b2(&full_bar[s], a_dst + a_bytes, t * p.cin + kc * p.bk, nt * p.n_tile + b_row0);
            }
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == IG_WARP_MMA && (!kPair || crank == 0)) {
    // ================================ MMA issuer (the leader's, in a pair) ==================================
    // The whole warp walks the pipeline with warp-uniform control flow; waits are done by one
    // lane + __syncwarp, and every (tap, k-chunk) sub-tile is issued by one asm block
    // (umma_tap<KS>) whose elected lane fires KS tcgen05.mma with 32-byte descriptor steps.
    const uint32_t idesc = umma_idesc_bf16(kPair ? 256 : 128, p.n_tile, 0, 0);
    auto commit = [&](uint64_t* bar) {
      if constexpr (kPair) umma_commit_elect_pair(bar);
      else umma_commit_elect(bar);
    };
    const uint64_t dkm = umma_smem_desc(0u, 0u, 8u * sw, sw);  // K-major, dense rows
    const uint32_t b_hi = static_cast<uint32_t>(dkm >> 32);
    const uint32_t a_hi_dense = b_hi;
    const uint32_t a_hi_halo =
        static_cast<uint32_t>(umma_smem_desc(0u, 0u, static_cast<uint32_t>((IG_HALO_TW + 2) * sw), sw) >> 32);
    const int ksteps = p.bk >> 4;
    const uint32_t sub16 = sub_bytes >> 4, a16 = a_bytes >> 4, b16 = b_bytes >> 4;
    const uint32_t row16 = static_cast<uint32_t>(sw) >> 4;
    int s = 0, sa = 0;
    uint32_t ph = 0, pha = 0;
    int acc = 0;
    uint32_t aph = 0;
    if (p.resident_b) {
      mbar_wait_warp(wfull_bar, 0u, 0xb00u);
      tc_fence_after();
    }
    // The issuing warp's own instruction stream is what bounds thin layers (ncu: with N <= 64 or
    // 16-channel K chunks an MMA retires faster than ~70 SASS instructions of generic per-tap control
    // flow), so the tap loop is specialised at compile time: KS = K steps per tap, and for the halo
    // mode TPS = taps per weight stage (9 / 3 / 1) or 0 = weights resident in shared memory. Inside,
    // every tap is straight-line code: two adds and one asm block.
    auto tile_loop = [&](auto ks_tag, auto tps_tag) {
      constexpr int KS = decltype(ks_tag)::value;
      constexpr int TPS = decltype(tps_tag)::value;
      uint32_t tap_a[9];  // halo-tile row offset of tap t = (dy, dx): (dy * 10 + dx) rows, in 16-byte units
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_a[t] = static_cast<uint32_t>((t / 3) * (IG_HALO_TW + 2) + (t % 3)) * row16;
      const uint32_t w16 = smem_u32(smem) >> 4;
      const uint32_t stage16 = stage_bytes >> 4;
      auto tap = [&](uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t bh, uint32_t id, uint32_t ac) {
        if constexpr (kPair) umma_tap_pair<KS>(d, a_lo, a_hi, b_lo, bh, id, ac);
        else umma_tap<KS>(d, a_lo, a_hi, b_lo, bh, id, ac);
      };
      for (int it = it0; it < total_its; it += it_step) {
        mbar_wait_warp(&tempty_bar[acc], aph ^ 1u, 0x200u + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.n_tile);
        uint32_t accum = 0;
        if (halo) {
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait_warp(&afull_bar[sa], pha, 0xa00u + sa);
            tc_fence_after();
            const uint32_t a_slot16 = smem_u32(aring + sa * halo_slot) >> 4;
            if constexpr (TPS == 0) {
              // resident weights: tile (t, kc) sits at (t * kchunks + kc) * b16
              const uint32_t bd0 = w16 + static_cast<uint32_t>(kc) * b16;
              const uint32_t bstep = static_cast<uint32_t>(kchunks) * b16;
#pragma unroll
              for (int t = 0; t < 9; ++t)
                tap(d_tmem, a_slot16 + tap_a[t], a_hi_halo, bd0 + static_cast<uint32_t>(t) * bstep, b_hi, idesc,
                             t == 0 ? accum : 1u);
            } else {
#pragma unroll
              for (int t0 = 0; t0 < 9; t0 += TPS) {
                mbar_wait_warp(&full_bar[s], ph, 0x300u + s);
                tc_fence_after();
                const uint32_t bd0 = w16 + static_cast<uint32_t>(s) * stage16;
#pragma unroll
                for (int u = 0; u < TPS; ++u)
                  tap(d_tmem, a_slot16 + tap_a[t0 + u], a_hi_halo, bd0 + static_cast<uint32_t>(u) * b16, b_hi,
                               idesc, (t0 + u) == 0 ? accum : 1u);
                commit(&empty_bar[s]);
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
              }
            }
            accum = 1u;
            commit(&aempty_bar[sa]);  // halo tile consumed by all nine taps
            __syncwarp();
            if (++sa == p.a_slots) { sa = 0; pha ^= 1u; }
          }
        } else {
          for (int si = 0; si < s_iters; ++si) {
            const int it0 = si * p.tps;
            const int nsub = (k_iters - it0) < p.tps ? (k_iters - it0) : p.tps;
            mbar_wait_warp(&full_bar[s], ph, 0x300u + s);
            tc_fence_after();
            const uint32_t st16 = w16 + static_cast<uint32_t>(s) * stage16;
            for (int u = 0; u < nsub; ++u) {
              const uint32_t ad = st16 + static_cast<uint32_t>(u) * sub16;
              tap(d_tmem, ad, a_hi_dense, ad + a16, b_hi, idesc, accum);
              accum = 1u;
            }
            commit(&empty_bar[s]);  // frees the slot once these MMAs retire
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
        commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        __syncwarp();
        if (++acc == p.acc_stages) { acc = 0; aph ^= 1u; }
      }
    };
    auto with_ks = [&](auto tps_tag) {
      if (ksteps == 4) tile_loop(std::integral_constant<int, 4>{}, tps_tag);
      else if (ksteps == 2) tile_loop(std::integral_constant<int, 2>{}, tps_tag);
      else tile_loop(std::integral_constant<int, 1>{}, tps_tag);
    };
    if (!halo || p.resident_b) with_ks(std::integral_constant<int, 0>{});
    else if (p.tps == 9) with_ks(std::integral_constant<int, 9>{});
    else if (p.tps == 3) with_ks(std::integral_constant<int, 3>{});
    else with_ks(std::integral_constant<int, 1>{});
  } else if (warp < 8) {
    // ================================ epilogue ====================================
    // Two groups of four warps. Group g drains every other tile of this CTA, so two tiles are in
    // their epilogue at once (thin-K layers are epilogue-bound). Per 64-channel column block:
    // TMEM -> registers (thread = output row) -> bias / ReLU / ReLU mask -> bf16 -> swizzled
    // staging tile -> one TMA store (ragged tiles are clipped by the TMA unit); the fused bias
    // gradient reads its column sums back from the staging tile.
    const int grp = warp >> 2;
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const bool elected = (quarter == 0 && lane == 0);
    const int nblk = p.n_tile / p.cw;
    const int chunks = p.cw >> 4;
    uint8_t* const stg0 = stg_base + grp * p.stg_bufs * stg_bytes;
    uint8_t* const pstg0 = pstg_base + grp * p.stg_bufs * pstg_bytes;
    // pooled-tile row of this lane's 2x2 window: the tile is 8 (x) by 16 (y) pixels, row m = x + 8 y,
    // so the window partners are lanes l ^ 1 (x) and l ^ 8 (y) of the same warp
    const int pm = ((lane & 7) >> 1) + 4 * (quarter * 2 + (lane >> 4));
    // accumulator stages: tile j of this CTA lives in stage j % acc_stages and is drained by group j % 2, so
    // group g alternates between stages g and g + 2 when four stages fit (N tile <= 128): the MMA warp can
    // then run two tiles ahead of each group and a group never waits for "its" next accumulator
    int acc = grp;
    uint32_t aph = 0;
    auto next_acc = [&]() {
      acc += 2;
      if (acc >= p.acc_stages) { acc = grp; aph ^= 1u; }
    };
    const int cpr = epi_rowb >> 4;        // 16-byte chunks per output row of a block: 8 / 4 / 2

    if (p.out_f32 != nullptr) {
      // ---- direct fp32 epilogue (small GEMMs whose consumer wants fp32) ----
      for (int tile = blockIdx.x + grp * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        bool valid = m < rows;
        long long obase = 0;
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
          }
        }
        mbar_wait(&tfull_bar[acc], aph, 0x400u + acc);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * p.n_tile);
        for (int c = 0; c < (p.n_tile >> 4); ++c) {
          const int n0 = nt * p.n_tile + c * 16;
          uint32_t v[16];
          tmem_ld16(t_row + static_cast<uint32_t>(c * 16), v);
          tmem_ld_wait();
          if (valid) {
            // pixel-shuffle (ConvTranspose2d k2 s2): column n = q * Cout + c lands in sub-pixel q = (qy, qx) of
            // the 2x2 block of this input pixel; a 16-column chunk never straddles two quadrants (Cout % 16 == 0)
            long long ocol = n0;
            if (p.epi_mode == IG_EPI_PIXSHUF) {
              const int q = n0 / p.shuf_cout;
              ocol = (n0 - q * p.shuf_cout) + (q & 1) * p.ostride[0] + (q >> 1) * p.ostride[2];
            }
            float4* op = reinterpret_cast<float4*>(p.out_f32 + obase + ocol);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[e] = __uint_as_float(v[4 * j + e]) + sbias[n0 + 4 * j + e];
                if (p.relu) f[e] = fmaxf(f[e], 0.f);
              }
              op[j] = make_float4(f[0], f[1], f[2], f[3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        next_acc();
      }
    } else {
      // ---- bf16 epilogue: TMEM -> registers (thread = output row) -> bias / ReLU -> bf16 ->
      // [ReLU mask of dgrad] -> swizzled smem -> TMA store, one 64-channel block at a time.
      // The mask is read straight from global memory by the thread that owns the row (16-byte
      // loads of its own 128-byte channel segment, issued one block ahead so their latency hides
      // behind the previous block) and applied on the packed bf16 pairs.
      int lrow[4];
      {
        int mm = m;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          lrow[j] = mm % p.boxM[j];
          mm /= p.boxM[j];
        }
      }
      const bool want_cs = !kPlainEpi && p.colsum_partial != nullptr;
      const bool want_mask = !kPlainEpi && p.mask != nullptr;
      const bool has_bias = p.bias != nullptr;
      constexpr bool want_tail = kPlainEpi && kTail;
      float tail_bias[3] = {0.f, 0.f, 0.f};
      if (want_tail) {
#pragma unroll
        for (int k = 0; k < 3; ++k) tail_bias[k] = __ldg(p.tail_b + k);
      }
      // tile geometry: origin, validity of this thread's row, its mask row offset
      auto tile_setup = [&](int tile, int* org, bool& valid, long long& moff) {
        int mt = ig_fastdiv(tile, p.fd_mul[4], p.fd_shr[4]);
        valid = m < rows;
        moff = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = ig_fastdiv(mt, p.fd_mul[j], p.fd_shr[j]);
          org[j] = (mt - q * p.ntile[j]) * p.boxM[j];
          mt = q;
          const int pj = org[j] + lrow[j];
          valid = valid && (pj < p.dimM[j]);
          moff += static_cast<long long>(pj) * p.mstride[j];
        }
      };
      auto mask_fetch = [&](uint4* mreg, bool valid, long long moff, int nglb) {
        const uint4* src = reinterpret_cast<const uint4*>(p.mask + moff + nglb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mreg[i] = make_uint4(0, 0, 0, 0);
          if (i < cpr && valid) mreg[i] = __ldg(src + i);
        }
      };
      const int istep = 2 * it_step;
      int it = it0 + grp * it_step;
      int tile = it < total_its ? tile_of(it) : 0;
      int org[4] = {0, 0, 0, 0};
      bool valid = false;
      long long moff = 0;
      uint4 mreg[8];
      auto tile_nt = [&](int t) { return t - ig_fastdiv(t, p.fd_mul[4], p.fd_shr[4]) * p.n_tiles_n; };
      if (it < total_its) {
        tile_setup(tile, org, valid, moff);
        if (want_mask && tile_nt(tile) * p.n_tile < p.mask_cols)
          mask_fetch(mreg, valid, moff, tile_nt(tile) * p.n_tile);
      }
      int sbuf = 0;   // staging tile of this group used by the current column block
      while (it < total_its) {
        const int tile_m = ig_fastdiv(tile, p.fd_mul[4], p.fd_shr[4]);
        const int nt = tile - tile_m * p.n_tiles_n;
        const int it_next = it + istep;
        const bool has_next = it_next < total_its;
        const int ntile_next = has_next ? tile_of(it_next) : 0;
        int org_n[4] = {0, 0, 0, 0};
        bool valid_n = false;
        long long moff_n = 0;
        // fused tail: this pixel's three target values, requested before the accumulator wait so that
        // their DRAM latency hides behind it and the conversion
        float ttgt[3] = {0.f, 0.f, 0.f};
        if (want_tail && p.tail_target != nullptr && valid) {
#pragma unroll
          for (int k = 0; k < 3; ++k) ttgt[k] = __ldg(p.tail_target + moff + k * p.tail_plane);
        }
        mbar_wait(&tfull_bar[acc], aph, 0x400u + acc);
        tc_fence_after();
        for (int cb = 0; cb < nblk; ++cb) {
          const int nloc = cb * p.cw;
          const int nglb = nt * p.n_tile + nloc;
          // staging tile free again? (with two tiles per group: the store issued two blocks ago has read it)
          uint8_t* const stg = stg0 + sbuf * stg_bytes;
          uint8_t* const pstg = pstg0 + sbuf * pstg_bytes;
          if (elected) {
            if (p.stg_bufs == 2) bulk_wait_read<1>();
            else bulk_wait_read<0>();
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (cb == 0 && has_next) tile_setup(ntile_next, org_n, valid_n, moff_n);
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(acc * p.n_tile + nloc);
          const bool blk_mask = want_mask && nglb < p.mask_cols;   // warp-uniform
          const bool blk_cs = want_cs && nglb < p.cs_cols;
          float* cs_dst = blk_cs ? p.colsum_partial + (static_cast<long long>(tile_m) * 4 + quarter) * p.n_total + nglb
                                  : nullptr;
          // G = 16-column chunks fetched from TMEM per wait: all four when registers allow (plain
          // forward epilogue), two when the ReLU mask / column sums keep more state live
          float tz[3] = {0.f, 0.f, 0.f};   // fused tail: conv8 pre-activations of this thread's pixel
          auto convert = [&](auto g_tag) {
            constexpr int G = decltype(g_tag)::value;
#pragma unroll
            for (int h = 0; h < 4 / G; ++h) {
              if (G * h < chunks) {
                uint32_t v[G][16];
#pragma unroll
                for (int c2 = 0; c2 < G; ++c2)
                  if (G * h + c2 < chunks) tmem_ld16(t_row + static_cast<uint32_t>((G * h + c2) * 16), v[c2]);
                tmem_ld_wait();
#pragma unroll
                for (int c2 = 0; c2 < G; ++c2) {
                  const int ch = G * h + c2;
                  if (ch < chunks) {
                    uint32_t pk[8];
                    float bv[16];
                    if (has_bias) {   // nglb and sbias are 16-byte aligned: four 128-bit shared loads
                      const float4* b4 = reinterpret_cast<const float4*>(sbias + nglb + ch * 16);
#pragma unroll
                      for (int j = 0; j < 4; ++j) {
                        const float4 t = b4[j];
                        bv[4 * j] = t.x; bv[4 * j + 1] = t.y; bv[4 * j + 2] = t.z; bv[4 * j + 3] = t.w;
                      }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      float f0 = __uint_as_float(v[c2][2 * j]), f1 = __uint_as_float(v[c2][2 * j + 1]);
                      if (has_bias) {
                        f0 += bv[2 * j];
                        f1 += bv[2 * j + 1];
                      }
                      pk[j] = p.relu ? pack_bf16x2_relu(f0, f1) : pack_bf16x2(f0, f1);
                    }
                    if (blk_mask) {
                      const uint32_t mw[8] = {mreg[2 * ch].x, mreg[2 * ch].y, mreg[2 * ch].z, mreg[2 * ch].w,
                                              mreg[2 * ch + 1].x, mreg[2 * ch + 1].y, mreg[2 * ch + 1].z, mreg[2 * ch + 1].w};
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                        __nv_bfloat162 mv;
                        memcpy(&mv, &mw[j], 4);
                        pk[j] &= __hgt2_mask(mv, __float2bfloat162_rn(0.f));
                      }
                    }
                    *reinterpret_cast<uint4*>(stg + swz_off(m, ch * 2, epi_rowb)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(stg + swz_off(m, ch * 2 + 1, epi_rowb)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    if constexpr (want_tail) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                        const float4 w0 = stail[ch * 16 + 2 * j], w1 = stail[ch * 16 + 2 * j + 1];
                        const float v0 = bf16_lo(pk[j]), v1 = bf16_hi(pk[j]);
                        tz[0] = fmaf(v0, w0.x, tz[0]); tz[1] = fmaf(v0, w0.y, tz[1]); tz[2] = fmaf(v0, w0.z, tz[2]);
                        tz[0] = fmaf(v1, w1.x, tz[0]); tz[1] = fmaf(v1, w1.y, tz[1]); tz[2] = fmaf(v1, w1.z, tz[2]);
                      }
                    }
                    if constexpr (kPlainEpi && kPool) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                        __nv_bfloat162 a, b;
                        uint32_t o = __shfl_xor_sync(0xffffffffu, pk[j], 1);
                        memcpy(&a, &pk[j], 4);
                        memcpy(&b, &o, 4);
                        a = __hmax2(a, b);
                        memcpy(&pk[j], &a, 4);
                        o = __shfl_xor_sync(0xffffffffu, pk[j], 8);
                        memcpy(&b, &o, 4);
                        a = __hmax2(a, b);
                        memcpy(&pk[j], &a, 4);
                      }
                      if ((lane & 9) == 0) {
                        *reinterpret_cast<uint4*>(pstg + swz_off(pm, ch * 2, epi_rowb)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(pstg + swz_off(pm, ch * 2 + 1, epi_rowb)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                      }
                    }
                    if (blk_cs && !blk_mask && !valid) {   // rows outside the image must not reach the column sums
                      *reinterpret_cast<uint4*>(stg + swz_off(m, ch * 2, epi_rowb)) = make_uint4(0, 0, 0, 0);
                      *reinterpret_cast<uint4*>(stg + swz_off(m, ch * 2 + 1, epi_rowb)) = make_uint4(0, 0, 0, 0);
                    }
                  }
                }
              }
            }
          };
          convert(std::integral_constant<int, kPlainEpi ? 4 : 2>{});
          if (want_tail) {
            // moff = b * 3 * plane + y * W + x (mstride set by the host); three planes per image
            float se = 0.f;
            if (valid) {
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const float z = tz[k] + tail_bias[k];
                const float yv = 1.f / (1.f + expf(-z));
                const long long o = moff + k * p.tail_plane;
                p.tail_out[o] = yv;
                if (p.tail_target != nullptr) {
                  const float d = yv - ttgt[k];
                  se = fmaf(d, d, se);
                }
              }
            }
            if (p.tail_loss_partial != nullptr) {
#pragma unroll
              for (int o2 = 16; o2 > 0; o2 >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o2);
              if (lane == 0) p.tail_loss_partial[static_cast<long long>(tile_m) * 4 + quarter] = se;
            }
          }
          // mask of the NEXT MASKED column block, straight into the registers this block has just finished
          // with: the next block of this tile if it is masked, else — as soon as the tile's last masked block
          // is converted, not at the end of the tile — the first block of the next tile. With only the up-conv
          // half of a concat gradient masked (conv7 dgrad) the loads get a whole unmasked block more to land.
          if (want_mask) {
            const bool more_here = cb + 1 < nblk && nglb + p.cw < p.mask_cols;
            const bool last_masked_here = blk_mask ? !more_here : (cb == 0);   // cb == 0: tile without masked blocks
            if (more_here) {
              mask_fetch(mreg, valid, moff, nglb + p.cw);
            } else if (last_masked_here && has_next) {
              const int n_next = tile_nt(ntile_next) * p.n_tile;
              if (n_next < p.mask_cols) mask_fetch(mreg, valid_n, moff_n, n_next);
            }
          }
          if (cb == nblk - 1) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (kPair) mbar_arrive_cluster(leader_addr(&tempty_bar[acc]));   // the leader's MMA warp waits
              else mbar_arrive(&tempty_bar[acc]);
            }
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (elected) {
            int c[5];
            c[1] = org[0]; c[2] = org[1]; c[3] = org[2]; c[4] = org[3];
            if (p.epi_mode == IG_EPI_PIXSHUF) {
              const int q = ig_fastdiv(nglb, p.fd_mul[5], p.fd_shr[5]);
              c[0] = nglb - q * p.shuf_cout;
              c[1] += q & 1;
              c[3] += q >> 1;
            } else {
              c[0] = nglb;
            }
            tma_store_5d(&tmOut, stg, c[0], c[1], c[2], c[3], c[4]);
            if constexpr (kPlainEpi && kPool) tma_store_5d(&tmMask, pstg, c[0], c[1] >> 1, c[2] >> 1, c[3], c[4]);
            bulk_commit();
          }
          if (blk_cs) {
            // Fused bias gradient: column sums of the STORED bf16 values over this warp's 32 rows, read
            // back from the staging tile — one channel pair per lane, one row per instruction (a row is
            // 32 distinct words, so the reads are conflict-free under any swizzle), rows added in a
            // fixed order. Narrow blocks (32 / 16 channels) fold 2 / 4 rows into one instruction and
            // combine the row subsets with shuffles. The tile is only overwritten after the group's
            // next barrier, which every thread reaches after these reads.
            const int npair = p.cw >> 1;                 // 32 / 16 / 8
            const int pair = lane & (npair - 1), rs = lane / npair, rstep = 32 / npair;
            float s0 = 0.f, s1 = 0.f, u0 = 0.f, u1 = 0.f;
            const uint32_t wofs = static_cast<uint32_t>(pair & 3) * 4u;
#pragma unroll 8
            for (int r = 0; r < 32; r += 2 * rstep) {
              const uint32_t wa = *reinterpret_cast<const uint32_t*>(stg + swz_off(quarter * 32 + r + rs, pair >> 2, epi_rowb) + wofs);
              const uint32_t wb = *reinterpret_cast<const uint32_t*>(stg + swz_off(quarter * 32 + r + rstep + rs, pair >> 2, epi_rowb) + wofs);
              s0 += bf16_lo(wa); s1 += bf16_hi(wa);
              u0 += bf16_lo(wb); u1 += bf16_hi(wb);
            }
            s0 += u0; s1 += u1;
            for (int off = npair; off < 32; off <<= 1) {
              s0 += __shfl_xor_sync(0xffffffffu, s0, off);
              s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            }
            if (rs == 0) *reinterpret_cast<float2*>(cs_dst + 2 * pair) = make_float2(s0, s1);
          }
          sbuf = (p.stg_bufs == 2) ? (sbuf ^ 1) : 0;
        }
        next_acc();
        tile = ntile_next;
        it = it_next;
#pragma unroll
        for (int j = 0; j < 4; ++j) org[j] = org_n[j];
        valid = valid_n;
        moff = moff_n;
      }
      if (elected) bulk_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // the peer's shared memory and TMEM are in use until the leader is done
  if (warp == IG_WARP_MMA) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// Fixed (non-pipeline) shared memory of a configuration (host side).
inline size_t igemm_fixed_smem(int cw, int pool2, int n_total, int a_slots = 0, int sw = 128, int stg_bufs = 1) {
  const size_t stg = (128 * static_cast<size_t>(cw) * 2 + (pool2 ? 32 * static_cast<size_t>(cw) * 2 : 0)) * stg_bufs;
  return 2048 + 16 + static_cast<size_t>(a_slots) * ig_halo_slot(sw) + 2 * stg +
         (2 * IG_MAX_STAGES + 2 * IG_MAX_ASLOTS + 10) * 8 + 16 +
         static_cast<size_t>(n_total) * 4 + 64 + 64 * 16 + 16;
}
inline size_t igemm_stage_bytes(int bk, int n_tile, int tps) {
  return (128 + static_cast<size_t>(n_tile)) * bk * 2 * tps;
}

}  // namespace rovr
