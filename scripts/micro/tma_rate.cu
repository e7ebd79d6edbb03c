// TMA load-rate micro-benchmark: one CTA per SM issues `iters` box loads round-robin over `depth`
// smem slots and waits on mbarriers; reports aggregate GB/s for several box shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_rate tma_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../../reinformcement-optimized-video-reconstruction_b200/csrc/ptx.cuh"
using namespace rovr;

__global__ void __launch_bounds__(64, 1)
tma_kernel(const __grid_constant__ CUtensorMap tm, int iters, int depth, uint32_t bytes, int slot_bytes,
           int c1max, int c2max, int c3max, int st1, int st2, int rank5, int* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + depth * slot_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ph[16] = {0};
    int issued = 0;
    int c1 = (blockIdx.x * 7) % c1max, c2 = (blockIdx.x * 3) % c2max, c3 = blockIdx.x % c3max;
    for (int it = 0; it < iters + depth; ++it) {
      const int s = it % depth;
      if (it >= depth) {  // wait for the load issued `depth` iterations ago
        mbar_wait(&bar[s], ph[s], 0x1);
        ph[s] ^= 1u;
      }
      if (it < iters) {
        mbar_expect_tx(&bar[s], bytes);
        if (rank5) tma_load_5d(&tm, &bar[s], smem + s * slot_bytes, 0, c1 - 1, c2 - 1, c3, 0);
        else tma_load_2d(&tm, &bar[s], smem + s * slot_bytes, 0, c1);
        ++issued;
        c1 += st1; if (c1 >= c1max) { c1 -= c1max; c2 += st2; if (c2 >= c2max) { c2 -= c2max; c3 = (c3 + 1) % c3max; } }
      }
    }
    if (issued == -1) *sink = 1;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 enc;
int main() {
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int B = 24, H = 256, W = 256;
  void* buf; cudaMalloc(&buf, (size_t)B * H * W * 128 * 2); cudaMemset(buf, 0, (size_t)B * H * W * 128 * 2);
  int* sink; cudaMalloc(&sink, 4);
  struct Case { const char* name; int C; int ld; int bx, by; int rank5; int sw; };
  // C = channels in the box (inner), ld = pixel stride of the tensor
  Case cases[] = {
    {"halo 10x18 x 64ch of ld128 (conv7 fprop A)", 64, 128, 10, 18, 1, 128},
    {"halo 10x18 x 64ch of ld64  (dense pixels)  ", 64, 64, 10, 18, 1, 128},
    {"halo 10x18 x 16ch of ld16  (conv1 fprop A) ", 16, 16, 10, 18, 1, 32},
    {"tile 8x16  x 64ch of ld128 (no halo)       ", 64, 128, 8, 16, 1, 128},
    {"2-D 128 rows x 64ch of ld128               ", 64, 128, 128, 1, 0, 128},
    {"2-D 128 rows x 64ch of ld64 (contiguous)   ", 64, 64, 128, 1, 0, 128},
    {"2-D 180 rows x 64ch of ld128               ", 64, 128, 180, 1, 0, 128},
  };
  for (auto& c : cases) {
    CUtensorMap tm;
    int rows;
    if (c.rank5) {
      cuuint64_t gd[5] = {(cuuint64_t)c.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 1};
      cuuint64_t gs[4] = {(cuuint64_t)c.ld * 2, (cuuint64_t)W * c.ld * 2, (cuuint64_t)H * W * c.ld * 2, (cuuint64_t)B * H * W * c.ld * 2};
      cuuint32_t bx[5] = {(cuuint32_t)c.C, (cuuint32_t)c.bx, (cuuint32_t)c.by, 1, 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       c.sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode failed %d\n", (int)r); continue; }
      rows = c.bx * c.by;
    } else {
      cuuint64_t gd[2] = {(cuuint64_t)c.C, (cuuint64_t)B * H * W};
      cuuint64_t gs[1] = {(cuuint64_t)c.ld * 2};
      cuuint32_t bx[2] = {(cuuint32_t)c.C, (cuuint32_t)c.bx};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode failed %d\n", (int)r); continue; }
      rows = c.bx;
    }
    const uint32_t bytes = rows * c.C * 2;
    const int slot = (bytes + 1023) / 1024 * 1024;
    for (int depth : {1, 2, 4, 8}) {
      if (depth * slot + 1024 + 256 > 227 * 1024) continue;
      const int iters = 2000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const size_t smem = depth * slot + 2048;
      auto launch = [&]() {
        if (c.rank5) tma_kernel<<<148, 64, smem>>>(tm, iters, depth, bytes, slot, W - c.bx, H - c.by, B, c.bx - 2, c.by - 2, 1, sink);
        else tma_kernel<<<148, 64, smem>>>(tm, iters, depth, bytes, slot, B * H * W - c.bx, 1, 1, c.bx, 0, 0, sink);
      };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      const double gb = (double)bytes * iters * 148 / 1e9;
      printf("%s depth %d: %7.1f GB/s  %6.1f ns/box  %5.1f cycles/row  (%s)\n", c.name, depth, gb / (ms * 1e-3),
             ms * 1e6 / iters, ms * 1e6 / iters * 1.965 / rows, cudaGetErrorString(err));
    }
  }
  return 0;
}
