// Micro-benchmark (r2, replaces the instruction-bound l1gather.cu of r1): cost of warp-wide scattered loads that HIT in L1
// (sm_100a) by load width and lane -> address pattern, with every address computed BEFORE the timed loop.
// Each thread keeps NOFF byte offsets in registers; the timed loop is NOFF loads + NOFF xors per trip, fully unrolled,
// so a coalesced 4-byte load issues in ~2 instructions and the LSU / L1 data stage is what is measured.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l1gather2 scripts/microbench/l1gather2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kRegion = 48 * 1024;  // pattern window; with the per-trip rotation every SM re-reads 96 KiB (fits L1)
constexpr int NOFF = 8;

__host__ __device__ inline uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// PAT: 0 broadcast, 1 coalesced, 2 32 lines same slot, 3 32 lines slot = lane % 8, 5 random W-byte chunks,
//      6 random 16-byte chunks inside 24 lines (a natural-image-like reuse set), 7 random W-byte chunks, 8 lanes share each
template <int PAT, int W>
__device__ __forceinline__ uint32_t pattern_off(int lane, int warp, int i) {
  uint32_t o;
  const uint32_t h = hash32((uint32_t)lane * 7919u + (uint32_t)i * 104729u + (uint32_t)warp * 1299709u);
  if (PAT == 0) o = (uint32_t)i * 4096u;
  else if (PAT == 1) o = (uint32_t)lane * W + (uint32_t)i * 4096u;
  else if (PAT == 2) o = (uint32_t)lane * 128u + (uint32_t)i * 4096u;
  else if (PAT == 3) o = (uint32_t)lane * (128u + W) + (uint32_t)i * 4096u;
  else if (PAT == 5) o = (h % (kRegion / W)) * W;
  else if (PAT == 6) o = ((h % 24u) * 128u + ((h >> 8) % (128u / W)) * W + (uint32_t)i * 4096u);
  else o = (hash32((uint32_t)(lane >> 3) * 7919u + (uint32_t)i * 104729u + (uint32_t)warp * 1299709u) % (kRegion / W)) * W;
  return o % kRegion;
}

template <int W>
__device__ __forceinline__ uint32_t ld(const uint8_t* p) {
  uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
  if (W == 4) { asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a0) : "l"(p)); return a0; }
  if (W == 8) { asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a0), "=r"(a1) : "l"(p)); return a0 ^ a1; }
  // every component is consumed: ptxas narrows a vector load whose lanes are dead
  if (W == 16) { asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "l"(p)); return (a0 ^ a1) ^ (a2 ^ a3); }
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(p));
  return ((a0 ^ a1) ^ (a2 ^ a3)) ^ ((a4 ^ a5) ^ (a6 ^ a7));
}

template <int PAT, int W>
__global__ void __launch_bounds__(512) gather(const uint8_t* __restrict__ buf, int iters, uint32_t* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* p[NOFF];
#pragma unroll
  for (int i = 0; i < NOFF; ++i) p[i] = buf + (pattern_off<PAT, W>(lane, warp, i) & ~(uint32_t)(W - 1));
  uint32_t acc = 0;
  // the whole pattern moves by 4 KiB per trip inside a second copy of the region (ld.global.nc of a loop-invariant
  // address would be hoisted by ptxas): one shared offset per trip, the lane -> address structure stays
  uint32_t rot = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NOFF; ++i) acc ^= ld<W>(p[i] + rot);
    rot = rot + 4096u >= (uint32_t)kRegion ? 0u : rot + 4096u;
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int PAT, int W>
static void run(const uint8_t* buf, uint32_t* sink, int sms) {
  const int iters = 500, blocks = sms * 4, threads = 512;  // 64 warps per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    gather<PAT, W><<<blocks, threads>>>(buf, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double loads_per_sm = (double)iters * NOFF * (threads / 32) * (blocks / sms);
  const double cyc = best * 1e-3 * 1.965e9 / loads_per_sm;  // at the 1965 MHz boost clock
  printf("%d,%d,%.2f,%.1f\n", PAT, W, cyc, 32.0 * W / cyc);
}

template <int W>
static void run_w(const uint8_t* buf, uint32_t* sink, int sms) {
  run<0, W>(buf, sink, sms); run<1, W>(buf, sink, sms); run<2, W>(buf, sink, sms); run<3, W>(buf, sink, sms);
  run<5, W>(buf, sink, sms); run<6, W>(buf, sink, sms); run<7, W>(buf, sink, sms);
}

int main() {
  uint8_t* buf; uint32_t* sink;
  cudaMalloc(&buf, 2 * kRegion); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, 2 * kRegion);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("pattern,width,cycles_per_warp_load_per_SM,bytes_per_clk_per_SM\n");
  run_w<4>(buf, sink, sms); run_w<8>(buf, sink, sms); run_w<16>(buf, sink, sms); run_w<32>(buf, sink, sms);
  return 0;
}
