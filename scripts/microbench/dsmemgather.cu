// Micro-benchmark (r2): random 16- / 32-byte gathers from a table spread over the shared memory of
// a thread-block cluster (ld.shared::cluster through mapa addresses), against local shared memory and against the
// 31.4 B/clk/SM of scattered L2-hit loads (profiles/r1e_l2gather.csv).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dsmemgather scripts/microbench/dsmemgather.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cg = cooperative_groups;
constexpr int kSlice = 128 * 1024;  // bytes of table per CTA

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int W>
__global__ void __launch_bounds__(512) gather(int csize, int iters, int local_only, uint32_t* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < kSlice / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  cluster.sync();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t my_rank = cluster.block_rank();
  uint32_t s = hash32(blockIdx.x * 1024u + threadIdx.x), acc = 0;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const uint32_t r = s >> 4;
    const uint32_t rank = local_only ? my_rank : r % (uint32_t)csize;
    const uint32_t off = ((r / 16u) % (kSlice / W)) * W;
    uint32_t addr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(base + off), "r"(rank));
    uint32_t a0, a1, a2, a3;
    asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
    acc ^= a0 ^ a3;
    if (W == 32) {
      asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr + 16));
      acc ^= a0 ^ a3;
    }
  }
  cluster.sync();
  if (acc == 0x12345678u) *sink = acc;
}

template <int W>
static void run(int csize, int local_only, uint32_t* sink, int sms) {
  const int iters = 2000, threads = 512;
  int blocks = (sms / csize) * csize;
  cudaFuncSetAttribute(gather<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlice);
  if (csize > 8) cudaFuncSetAttribute(gather<W>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = kSlice;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, gather<W>, csize, iters, local_only, sink);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("%d,%d,%d,launch failed: %s\n", csize, W, local_only, cudaGetErrorString(e)); cudaGetLastError(); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double lookups = (double)iters * threads * blocks, clk = best * 1e-3 * 1.965e9;
  printf("%d,%d,%d,%d,%.2f,%.3f\n", csize, W, local_only, blocks, lookups * W / clk / blocks, lookups / clk / blocks);
}

int main() {
  uint32_t* sink; cudaMalloc(&sink, 4);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("cluster_size,bytes_per_lane,local_only,ctas,bytes_per_clk_per_SM,lane_lookups_per_clk_per_SM\n");
  for (int cs : {1, 2, 4, 8, 16}) {
    run<16>(cs, 0, sink, sms);
    run<32>(cs, 0, sink, sms);
  }
  run<16>(8, 1, sink, sms);
  run<32>(8, 1, sink, sms);
  return 0;
}
