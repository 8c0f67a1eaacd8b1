// Micro-benchmark: warp-wide scattered loads that MISS L1 and HIT L2 (sm_100a): how many bytes per clock per SM come back
// as a function of the contiguous bytes each lane asks for (16, 32, 64, 128).  The LUT stages sit on this path
// (one 16-byte cell or 32-byte max-tap block per lookup, L1 hit rate ~2 % on high-entropy input).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l2gather scripts/microbench/l2gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr size_t kRegion = 48ull << 20;  // 48 MiB: the size of the six stage-2 tables; L2-resident, far above L1

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint32_t ld32B(const uint8_t* p) {
  uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(p));
  return a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
__device__ __forceinline__ uint32_t ld16B(const uint8_t* p) {
  uint32_t a0, a1, a2, a3;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "l"(p));
  return a0 ^ a1 ^ a2 ^ a3;
}

template <int W>
__global__ void gather(const uint8_t* __restrict__ buf, int iters, uint32_t* sink) {
  uint32_t acc = 0, s = hash32(blockIdx.x * 1024u + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const size_t off = ((size_t)(s >> 4) % (kRegion / W)) * W;
    const uint8_t* p = buf + off;
    if (W == 16) acc ^= ld16B(p);
    if (W >= 32) acc ^= ld32B(p);
    if (W >= 64) acc ^= ld32B(p + 32);
    if (W >= 128) acc ^= ld32B(p + 64) ^ ld32B(p + 96);
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  uint8_t* buf; uint32_t* sink;
  cudaMalloc(&buf, kRegion); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, kRegion);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 400, blocks = sms * 8, threads = 256;  // 64 warps per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("bytes_per_lane,GBps,bytes_per_clk_per_SM,lane_lookups_per_clk_per_SM\n");
  for (int w : {16, 32, 64, 128}) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (w == 16) gather<16><<<blocks, threads>>>(buf, iters, sink);
      if (w == 32) gather<32><<<blocks, threads>>>(buf, iters, sink);
      if (w == 64) gather<64><<<blocks, threads>>>(buf, iters, sink);
      if (w == 128) gather<128><<<blocks, threads>>>(buf, iters, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double lookups = (double)iters * threads * blocks, bytes = lookups * w, clk = best * 1e-3 * 1.965e9;
    printf("%d,%.0f,%.1f,%.2f\n", w, bytes / (best * 1e-3) / 1e9, bytes / clk / sms, lookups / clk / sms);
  }
  return 0;
}
