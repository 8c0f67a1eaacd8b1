// Micro-benchmark: cost of warp-wide scattered loads that HIT in L1 (sm_100a), as a function of load width and of the
// lane -> address pattern.  Used to choose the LUT table layouts.  Build: nvcc -arch=sm_100a -O3.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kRegion = 48 * 1024;  // bytes every SM keeps re-reading (fits L1)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// pattern -> byte offset (multiple of 16) inside the region for (lane, iteration)
__device__ __forceinline__ uint32_t pattern_off(int pat, int lane, int it, int w) {
  const uint32_t rot = (uint32_t)it * 4096u;  // moves the whole pattern, keeps its structure
  uint32_t o;
  switch (pat) {
    case 0: o = 0; break;                                         // broadcast
    case 1: o = lane * w; break;                                  // coalesced
    case 2: o = lane * 128; break;                                // 32 lines, same slot
    case 3: o = lane * (128 + 16); break;                         // 32 lines, slot = lane % 8
    case 4: o = lane * 128 + (lane >> 3) * 16; break;             // 32 lines, slot = quarter
    case 5: o = (hash32(lane * 7919u + it * 104729u) % (kRegion / 16)) * 16; break;  // random 16-B chunks
    case 6: o = (lane & 7) * 128 + (lane >> 3) * 16; break;       // 8 lines; a quarter = 8 lines same slot
    case 7: o = (lane >> 2) * 128; break;                         // 8 lines, 4 lanes share one chunk
    case 8: o = (lane >> 2) * 128 + (lane & 3) * 16; break;       // 8 lines, 4 lanes = 4 chunks of the line
    case 9: o = (lane & 7) * (128 + 16) + (lane >> 3) * 1024 * 4; break;  // quarter: 8 lines 8 slots
    case 10: o = (hash32(lane * 7919u + it * 104729u) % 24) * (128 + 16); break;  // 24 distinct random of a small set
    case 11: o = (hash32(lane * 7919u + it * 104729u) % (kRegion / 32)) * 32; break;  // random 32-B sectors
    case 12: o = (hash32(lane * 7919u + it * 104729u) % (kRegion / 4)) * 4; break;  // random words (w=4 only)
    default: o = 0;
  }
  return (o + rot) % kRegion;
}

template <int W>
__global__ void gather(const uint8_t* __restrict__ buf, int pat, int iters, uint32_t* sink) {
  const int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t off = pattern_off(pat, lane, it + (threadIdx.x >> 5) * 3, W);
    off &= ~(uint32_t)(W - 1);
    if (W == 4) acc ^= __ldg(reinterpret_cast<const uint32_t*>(buf + off));
    if (W == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(buf + off)); acc ^= v.x ^ v.y; }
    if (W == 16) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + off)); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (W == 32) {
      uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
      asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(buf + off));
      acc ^= a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  uint8_t* buf; uint32_t* sink;
  cudaMalloc(&buf, kRegion); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, kRegion);
  int sms = 148, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 2000, blocks = sms * 4, threads = 512;  // 64 warps per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("pattern,width,cycles_per_warp_load_per_SM\n");
  for (int w : {4, 8, 16, 32})
    for (int pat = 0; pat <= 12; ++pat) {
      if (pat == 12 && w != 4) continue;
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (w == 4) gather<4><<<blocks, threads>>>(buf, pat, iters, sink);
        if (w == 8) gather<8><<<blocks, threads>>>(buf, pat, iters, sink);
        if (w == 16) gather<16><<<blocks, threads>>>(buf, pat, iters, sink);
        if (w == 32) gather<32><<<blocks, threads>>>(buf, pat, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      const double loads_per_sm = (double)iters * (threads / 32) * (blocks / sms);
      const double cyc = best * 1e-3 * 1.965e9 / loads_per_sm;  // at the 1965 MHz boost clock
      printf("%d,%d,%.2f\n", pat, w, cyc);
    }
  return 0;
}
