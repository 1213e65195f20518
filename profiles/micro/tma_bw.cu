// Micro-benchmark (profiles/ tooling): how many bytes per second one SM, and all 148 together, can pull into shared memory with
// TMA tile loads -- the ceiling under the attention kernels' operand traffic.
//   tma_bw <mode> <box_rows> <depth> [grid] [producer warps] [64-column chunks per box]     mode 0: every CTA re-reads its own 256 KB window (L2 hits), 1: streams through HBM
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) {
  asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(b), "r"(par) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(m), "r"(bar), "r"(x), "r"(y), "r"(0) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int mode, int box_rows, int depth, int iters,
                                           int rows_total, int chunks, unsigned long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[64];
  const int pw = threadIdx.x >> 5, bytes = box_rows * 128 * chunks;
  const uint32_t base = smem_u32(sm) + pw * depth * bytes, bar = smem_u32(bars) + pw * 128;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(bar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int win_rows = 2048;                         // 256 KB window per CTA in mode 0
    const unsigned r0 = (mode == 0 ? blockIdx.x * win_rows : blockIdx.x * box_rows) + pw * 977 * box_rows;
    const unsigned stride = mode == 0 ? box_rows : gridDim.x * box_rows;
    // no divisions in the issue loop (a 64-bit modulo costs more than the TMA instruction)
    auto row_of = [&](int i) { return (int)((mode == 0 ? r0 + (((unsigned)i * box_rows) & (win_rows - 1)) : r0 + (unsigned)i * stride) & (rows_total - 1)); };
    const unsigned long long t0 = clock64();
    for (int i = 0; i < depth; ++i) { mbar_expect(bar + 8 * i, bytes); tma2d(base + i * bytes, &map, bar + 8 * i, 0, row_of(i)); }
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      mbar_wait(bar + 8 * s, (i / depth) & 1);
      if (i + depth < iters) { mbar_expect(bar + 8 * s, bytes); tma2d(base + s * bytes, &map, bar + 8 * s, 0, row_of(i + depth)); }
    }
    if (pw == 0) cyc[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, box_rows = argc > 2 ? atoi(argv[2]) : 64, depth = argc > 3 ? atoi(argv[3]) : 8;
  const int grid = argc > 4 ? atoi(argv[4]) : 148, iters = 4096;
  const int nprod = argc > 5 ? atoi(argv[5]) : 1, chunks = argc > 6 ? atoi(argv[6]) : 1;
  const int rows_total = 1 << 20;                      // 1 Mi rows x 1536 B = 1.5 GiB
  void* buf; cudaMalloc(&buf, (size_t)rows_total * 1536); cudaMemset(buf, 1, (size_t)rows_total * 1536);
  unsigned long long* cyc; cudaMalloc(&cyc, grid * 8);
  CUtensorMap map;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows_total, 12}, strides[2] = {1536, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)chunks}, es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
  const int smem = nprod * depth * box_rows * 128 * chunks + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<<<grid, 32 * nprod, smem>>>(map, mode, box_rows, depth, iters, rows_total, chunks, cyc);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)grid * iters * box_rows * 128 * chunks * nprod;
    printf("mode %d box_rows %d depth %d grid %d nprod %d chunks %d: %.1f us, %.2f TB/s total, %.1f GB/s per SM\n", mode, box_rows, depth, grid, nprod, chunks, ms * 1e3,
           bytes / ms / 1e9, bytes / grid / ms / 1e6);
  }
  return 0;
}
