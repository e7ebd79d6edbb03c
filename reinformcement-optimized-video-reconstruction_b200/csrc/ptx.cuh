// ptx.cuh — thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences) and UMMA shared-memory + instruction descriptors.
// Everything here is hand-written for B200; nothing is taken from the reference (which has no
// native code at all — SURVEY.md §2).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace rovr {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
// One lane of a converged warp (always the same one for a full mask). Guarding single-thread
// instructions (tcgen05.mma, TMA issue) with elect.sync rather than `lane == 0` lets ptxas keep
// their operands in uniform registers instead of emitting a per-thread waterfall loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// A spin that never hangs the box: after ~2^31 SM cycles (about a second) the waiter records
// which barrier it was stuck on and gives up (as does every later waiter), so a protocol bug
// becomes a kernel that finishes with wrong results plus a non-zero rovr_hang_code(), instead of
// a GPU that has to be reset.
__device__ unsigned int g_rovr_hang_code = 0;

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch. A kernel of the LocalNet step calls pdl_trigger() first thing, which lets
// the NEXT launch on the stream (made with launch_chain() below) place its CTAs on SMs as soon as they
// free up, and run its prologue (barrier init, TMEM allocation, descriptor prefetch) while this grid is
// still draining. pdl_wait() returns once every grid launched before has completed and flushed; nothing
// before it may touch global memory. Both are no-ops in a launch without the attribute.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ROVR_PDL=1 enables programmatic stream serialisation; the default is plain stream-ordered launches. Round 1 enabled
// it for the eager path when that was bound by launch latency (3.55 -> 3.40 ms); with today's step (back-to-back
// persistent kernels that fill every SM, so a dependent grid has nowhere to start early) an alternating A/B on one
// B200 measured it 1.5-3 % SLOWER for eager launches and graph replays alike (3.17 / 3.19 / 3.17 vs 3.21 / 3.23 /
// 3.25 ms eager; 3.17 / 3.19 / 3.20 vs 3.31 / 3.23 / 3.27 ms replay; gpurun_out/r2_pdl_ab.log -> profiles/).
inline int pdl_mode() {
  static const int v = [] { const char* e = getenv("ROVR_PDL"); return e ? atoi(e) : 0; }();
  return v;
}
// Launch `kernel` behind the previous launch of the stream with programmatic stream serialisation, optionally
// as clusters of `cluster_x` CTAs. Argument types are converted to the kernel's parameter types.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_mode()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned code) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  unsigned polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++polls) & 255u) == 0u) {
      // once any waiter has timed out, every later wait gives up at once so the grid drains
      if (*reinterpret_cast<volatile unsigned int*>(&g_rovr_hang_code) != 0u) return;
      if ((clock64() - t0) > (1ll << 31)) {
        atomicCAS(&g_rovr_hang_code, 0u, code | 0x80000000u);
        return;
      }
    }
  }
}

// Wait used by the single-warp producer / MMA roles. Every lane polls: measured on B200, letting
// one lane poll and parking the other 31 at __syncwarp() made the MMA issue stream ~1.8x slower
// (the lone poller is slow to observe the phase flip), so all 32 lanes spin and then reconverge.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, unsigned code) {
  mbar_wait(bar, parity, code);
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// TMA: tiled tensor loads, global -> shared, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: tensor memory + 5th-gen tensor core MMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate, issued by the elected lane of a converged
// warp: elect.sync and the predicated MMA sit in
// one asm block so that no branch is generated and ptxas can keep the (warp-uniform) operands in
// uniform registers. All 32 lanes must execute this with identical arguments.
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One (tap, k-chunk) sub-tile: KS (1, 2 or 4) K=16 MMAs whose descriptors advance by 32 bytes.
// Descriptors are passed as 32-bit halves (the high word is loop-invariant, the low word is
// "address >> 4" and never carries out of its 14-bit field), which lets ptxas do the per-MMA
// arithmetic with a handful of integer adds and register moves instead of 64-bit chains.
template <int KS>
__device__ __forceinline__ void umma_tap(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate);
template <>
__device__ __forceinline__ void umma_tap<1>(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                            uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t.reg .b64 a0, b0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%3, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a0, b0, %5, pa;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_tap<2>(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                            uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 a0, b0, a1, b1;\n\t.reg .b32 l1, m1;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
      "add.u32 l1, %1, 2;\n\tadd.u32 m1, %3, 2;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%3, %4};\n\t"
      "mov.b64 a1, {l1, %2};\n\tmov.b64 b1, {m1, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a0, b0, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %5, pt;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_tap<4>(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                            uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 a0, b0, a1, b1, a2, b2, a3, b3;\n\t"
      ".reg .b32 l1, l2, l3, m1, m2, m3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
      "add.u32 l1, %1, 2;\n\tadd.u32 l2, %1, 4;\n\tadd.u32 l3, %1, 6;\n\t"
      "add.u32 m1, %3, 2;\n\tadd.u32 m2, %3, 4;\n\tadd.u32 m3, %3, 6;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%3, %4};\n\t"
      "mov.b64 a1, {l1, %2};\n\tmov.b64 b1, {m1, %4};\n\t"
      "mov.b64 a2, {l2, %2};\n\tmov.b64 b2, {m2, %4};\n\t"
      "mov.b64 a3, {l3, %2};\n\tmov.b64 b3, {m3, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a0, b0, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %5, pt;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int KS>
__device__ __forceinline__ void umma_tap_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate);
template <>
__device__ __forceinline__ void umma_tap_pair<1>(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                            uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t.reg .b64 a0, b0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%3, %4};\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a0, b0, %5, pa;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_tap_pair<2>(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                            uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 a0, b0, a1, b1;\n\t.reg .b32 l1, m1;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
      "add.u32 l1, %1, 2;\n\tadd.u32 m1, %3, 2;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%3, %4};\n\t"
      "mov.b64 a1, {l1, %2};\n\tmov.b64 b1, {m1, %4};\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a0, b0, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %5, pt;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_tap_pair<4>(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                            uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 a0, b0, a1, b1, a2, b2, a3, b3;\n\t"
      ".reg .b32 l1, l2, l3, m1, m2, m3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
      "add.u32 l1, %1, 2;\n\tadd.u32 l2, %1, 4;\n\tadd.u32 l3, %1, 6;\n\t"
      "add.u32 m1, %3, 2;\n\tadd.u32 m2, %3, 4;\n\tadd.u32 m3, %3, 6;\n\t"
      "mov.b64 a0, {%1, %2};\n\tmov.b64 b0, {%3, %4};\n\t"
      "mov.b64 a1, {l1, %2};\n\tmov.b64 b1, {m1, %4};\n\t"
      "mov.b64 a2, {l2, %2};\n\tmov.b64 b2, {m2, %4};\n\t"
      "mov.b64 a3, {l3, %2};\n\tmov.b64 b3, {m3, %4};\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a0, b0, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %5, pt;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has retired (elected lane of a warp)
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread t <-> lane base+t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster on one TPC execute ONE 256-row MMA. Each CTA
// stages its own 128 rows of A and HALF of B (N/2 rows); the leader (cluster rank 0) issues the
// MMA and owns the "data ready" barriers, commits are multicast to both CTAs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in the cluster's rank-0 CTA
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair loads: data into THIS CTA's shared memory, bytes counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_elect_pair(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t.reg .b16 msk;\n\t"
      "mov.b16 msk, 3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], msk;\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version = 1 on sm_100
//   [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
// Operand tiles here are always "rows of `sw` bytes" written by TMA with swizzle mode `sw`:
//   K-major  (row = one M/N index, the row's bytes run along K): 8-row groups are 8*sw bytes
//            apart -> SBO = 8*sw; LBO unused.
//   MN-major (row = one K index, the row's bytes run along M/N): a 16-byte-unit column block is
//            `sw` bytes wide; blocks of sw/2 elements along M/N are LBO apart, 8-row K groups are
//            SBO = 8*sw apart.
__host__ __device__ __forceinline__ uint32_t umma_layout_code(int sw_bytes) {
  return sw_bytes == 128 ? 2u : (sw_bytes == 64 ? 4u : (sw_bytes == 32 ? 6u : 0u));
}
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, int sw_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(umma_layout_code(sw_bytes)) << 61;
  return d;
}
// Instruction descriptor (32 bit) for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt  [15] A MN-major  [16] B MN-major
//   [17,23) N >> 3         [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major,
                                                             int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------
// small numeric helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// round-to-nearest-even pack with the ReLU folded into the conversion (cvt.rn.relu.bf16x2.f32):
// one instruction instead of two FMNMX + one F2FP per pair in the forward epilogues
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

}  // namespace rovr
