// Micro-benchmark: scattered 32-byte loads that miss L1 and hit L2, as a function of the ADDRESS SPAN they are spread
// over (sm_100a).  The distinct bytes touched stay at 32 MiB (L2-resident); only the number of 2 MiB pages grows.
// Question (r2a): is the paired-window stage-2 kernel (6 tables x 24 planes of 2 MiB) held back by TLB misses?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tlbgather scripts/microbench/tlbgather.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

// touched set: `chunks` chunks of `chunk_bytes` each, chunk c at byte offset c * stride
__global__ void gather(const uint8_t* __restrict__ buf, uint32_t chunks, uint32_t chunk_sectors, size_t stride, int iters,
                       uint32_t* sink) {
  uint32_t acc = 0, s = hash32(blockIdx.x * 1024u + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const uint32_t r = s >> 4;
    const uint32_t c = r % chunks, w = (r / chunks) % chunk_sectors;
    acc ^= ld32B(buf + (size_t)c * stride + (size_t)w * 32);
  }
  if (acc == 0x12345678u) *sink = acc;
}

// argv[1] = "vmm": back the span with ONE cuMemCreate allocation mapped at a 1 GiB-aligned address (does the driver then
// use larger page-table entries than the 2 MiB of cudaMalloc?)
static uint8_t* vmm_alloc(size_t bytes) {
  typedef CUresult (*F_gran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  typedef CUresult (*F_create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  typedef CUresult (*F_reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  typedef CUresult (*F_map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  typedef CUresult (*F_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  F_gran gran; F_create create; F_reserve reserve; F_map map; F_access access;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuMemGetAllocationGranularity", (void**)&gran, cudaEnableDefault, &qr) != cudaSuccess) return nullptr;
  cudaGetDriverEntryPoint("cuMemCreate", (void**)&create, cudaEnableDefault, &qr);
  cudaGetDriverEntryPoint("cuMemAddressReserve", (void**)&reserve, cudaEnableDefault, &qr);
  cudaGetDriverEntryPoint("cuMemMap", (void**)&map, cudaEnableDefault, &qr);
  cudaGetDriverEntryPoint("cuMemSetAccess", (void**)&access, cudaEnableDefault, &qr);
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = 0;
  size_t gmin = 0, grec = 0;
  gran(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
  gran(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  printf("# vmm: granularity minimum %zu, recommended %zu bytes\n", gmin, grec);
  CUmemGenericAllocationHandle h;
  if (create(&h, bytes, &prop, 0) != CUDA_SUCCESS) return nullptr;
  CUdeviceptr va = 0;
  if (reserve(&va, bytes, 1ull << 30, 0, 0) != CUDA_SUCCESS) return nullptr;
  if (map(va, bytes, 0, h, 0) != CUDA_SUCCESS) return nullptr;
  CUmemAccessDesc ad = {};
  ad.location = prop.location;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (access(va, bytes, &ad, 1) != CUDA_SUCCESS) return nullptr;
  return (uint8_t*)va;
}

int main(int argc, char** argv) {
  const size_t span_max = 3072ull << 20;
  uint8_t* buf; uint32_t* sink;
  cudaFree(0);
  if (argc > 1) {
    buf = vmm_alloc(span_max);
    if (!buf) { printf("vmm allocation failed\n"); return 1; }
  } else if (cudaMalloc(&buf, span_max) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, span_max);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 400, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("touched_MiB,chunk_KiB,pages_2MiB,span_MiB,GBps,bytes_per_clk_per_SM\n");
  const size_t touched = 32ull << 20;
  // (chunk size, stride): chunk = contiguous touched bytes; stride >= chunk spreads the chunks over more pages
  struct Cfg { size_t chunk, stride; } cfgs[] = {
    {32ull << 20, 32ull << 20},                       // 16 pages, dense
    {2ull << 20, 2ull << 20},                         // same thing as 16 chunks
    {1ull << 20, 2ull << 20},                         // 32 pages half used
    {512ull << 10, 2ull << 20},                       // 64 pages
    {256ull << 10, 2ull << 20},                       // 128 pages
    {128ull << 10, 2ull << 20},                       // 256 pages
    {64ull << 10, 2ull << 20},                        // 512 pages
    {32ull << 10, 2ull << 20},                        // 1024 pages
    {21845ull * 32, 2ull << 20},                      // 48 pages of 682 KiB (like one PW table: 24 planes... x2)
  };
  for (const Cfg& c : cfgs) {
    const uint32_t chunks = (uint32_t)(touched / c.chunk);
    const size_t span = (size_t)chunks * c.stride;
    if (span > span_max) continue;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      gather<<<blocks, threads>>>(buf, chunks, (uint32_t)(c.chunk / 32), c.stride, iters, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double bytes = (double)iters * threads * blocks * 32, clk = best * 1e-3 * 1.965e9;
    printf("%zu,%zu,%u,%zu,%.0f,%.1f\n", touched >> 20, c.chunk >> 10, chunks, span >> 20, bytes / (best * 1e-3) / 1e9, bytes / clk / sms);
  }
  return 0;
}
