// Micro-benchmark (r2): cp.async.bulk.tensor.2d.tile::gather4 -- four arbitrary 32-byte table rows
// per instruction, L2 -> shared memory, bypassing the LSU / L1 fill path -- against the 31.4 B/clk/SM that scattered
// 32-byte ld.global.nc loads get on B200 (profiles/r1e_l2gather.csv).
// Every lane of a warp issues its own gather4 (4 rows = 128 bytes) per trip, a warp-wide mbarrier counts the 4 KiB.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmagather4 scripts/microbench/tmagather4.cu
// Run:   build/tmagather4 [box_rows=1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kWarps = 16;
constexpr uint32_t kRows = (48u << 20) / 32;  // 48 MiB table of 32-byte rows (L2-resident)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kWarps * 32) gather4(const __grid_constant__ CUtensorMap tmap, int iters, int readback, uint32_t* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* mine = smem + ((size_t)warp * 32 + lane) * 128;
  const uint32_t b = smem_u32(&bar[warp]);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t s = hash32(blockIdx.x * 1024u + threadIdx.x), acc = 0, phase = 0;
  for (int it = 0; it < iters; ++it) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4096u) : "memory");
    __syncwarp();
    int r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { s = s * 1664525u + 1013904223u; r[k] = (int)((s >> 4) % kRows); }
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(mine)), "l"(&tmap), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(b) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(phase) : "memory");
    phase ^= 1;
    if (readback) {
      const uint4* q = reinterpret_cast<const uint4*>(mine);
#pragma unroll
      for (int k = 0; k < 8; ++k) { const uint4 v = q[k]; acc ^= v.x ^ v.w; }
    }
    __syncwarp();
  }
  if (acc == 0x12345678u) *sink = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  uint8_t* buf; uint32_t* sink;
  cudaMalloc(&buf, (size_t)kRows * 32); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, (size_t)kRows * 32);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr) != cudaSuccess || !encode) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {8, kRows};
  const cuuint64_t strides[1] = {32};
  const cuuint32_t box[2] = {8, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 1; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = (size_t)kWarps * 32 * 128;  // 64 KiB
  cudaFuncSetAttribute(gather4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("box_rows,ctas_per_sm,readback,GBps,bytes_per_clk_per_SM,rows_per_clk_per_SM\n");
  const int iters = 100;
  for (int occ : {1, 2, 3})
    for (int readback : {0, 1}) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        gather4<<<sms * occ, kWarps * 32, smem>>>(tmap, iters, readback, sink);
        cudaEventRecord(e1);
        const cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      const double rows = (double)iters * 4 * kWarps * 32 * sms * occ, clk = best * 1e-3 * 1.965e9;
      printf("%d,%d,%d,%.0f,%.2f,%.3f\n", box_rows, occ, readback, rows * 32 / (best * 1e-3) / 1e9, rows * 32 / clk / sms, rows / clk / sms);
    }
  return 0;
}
