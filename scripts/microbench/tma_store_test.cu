// Stand-alone probe: which cp.async.bulk.tensor (TMA) store configurations work for a float32 [P][H][W] tensor on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_store_test scripts/microbench/tma_store_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
__global__ void store_kernel(const __grid_constant__ CUtensorMap tmap, int bx, int by, int x0, int y0, int z0) {
  extern __shared__ __align__(128) float tile[];
  for (int i = threadIdx.x; i < bx * by; i += blockDim.x) tile[i] = (float)(1000 * (i / bx) + (i % bx));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tmap), "r"(x0), "r"(y0),
                   "r"(z0), "r"(src)
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap), "r"(x0), "r"(y0), "r"(src)
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int P = 3, H = 400, W = 400;
  float* out;
  cudaMalloc(&out, sizeof(float) * P * H * W);
  float* host = (float*)malloc(sizeof(float) * P * H * W);
  struct Case { int rank, bx, by, x0, y0, z0; };
  const Case cases[] = {{2, 32, 8, 64, 16, 0},   {3, 32, 8, 64, 16, 1},   {3, 64, 16, 64, 16, 1}, {3, 128, 32, 128, 32, 1},
                        {3, 128, 32, 126, 30, 1}, {3, 128, 32, -2, -2, 0}, {3, 128, 32, 382, 382, 2}};
  for (const Case& c : cases) {
    cudaMemset(out, 0, sizeof(float) * P * H * W);
    CUtensorMap tm;
    const cuuint64_t dims3[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}, str3[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint64_t dims2[2] = {(cuuint64_t)W, (cuuint64_t)H * P}, str2[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.rank, out, c.rank == 3 ? dims3 : dims2, c.rank == 3 ? str3 : str2, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const size_t smem = sizeof(float) * c.bx * c.by;
    if (c.rank == 3) {
      cudaFuncSetAttribute(store_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      store_kernel<3><<<1, 256, smem>>>(tm, c.bx, c.by, c.x0, c.y0, c.z0);
    } else {
      cudaFuncSetAttribute(store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      store_kernel<2><<<1, 256, smem>>>(tm, c.bx, c.by, c.x0, c.y0, c.z0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long written = 0, wrong = 0;
    if (e == cudaSuccess) {
      cudaMemcpy(host, out, sizeof(float) * P * H * W, cudaMemcpyDeviceToHost);
      for (int z = 0; z < P; ++z)
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            const float v = host[((size_t)z * H + y) * W + x];
            const int i = y - c.y0, j = x - c.x0;
            const bool in = (c.rank == 2 || z == c.z0) && i >= 0 && i < c.by && j >= 0 && j < c.bx;
            const float want = in ? (float)(1000 * i + j) : 0.0f;
            if (v != 0.0f) ++written;
            if (v != want && !(in && i == 0 && j == 0)) ++wrong;
          }
    }
    printf("rank %d box %dx%d at (%d,%d,%d): encode %d, run %s, nonzero %lld, wrong %lld\n", c.rank, c.bx, c.by, c.x0, c.y0, c.z0, (int)r,
           cudaGetErrorString(e), written, wrong);
    if (e != cudaSuccess) {
      cudaDeviceReset();
      cudaMalloc(&out, sizeof(float) * P * H * W);
    }
  }
  return 0;
}
