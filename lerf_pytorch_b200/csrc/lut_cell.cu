// Rotation-ensembled LUT stages on cell-packed tables (sm_100a): table build + plain launches.
// Kernel body: lut_cell_body.cuh; lookup primitive: lut_cell.cuh.
#include <string.h>

#include <vector>

#include "lut_cell_body.cuh"
#ifdef LERF_EXPERIMENTS
#include "lut_mix.cuh"
#include "lut_mt.cuh"
#endif

namespace lerf {

using namespace cellk;

namespace {

// PAIRED (oC = 1): two lookups per 16x2 sorting network (cell::simplex_pair_of) -- production; false = one sort per lookup.
template <int STAGE, int OC, int MINB, bool PAIRED = false>
__global__ void __launch_bounds__(kTX* kTY, MINB)
    lut_stage_cell_kernel(CellTables tabs, const uint8_t* __restrict__ in, InAddr ia, int H, int W, int y0, int y1,
                          uint8_t* __restrict__ out) {
  __shared__ uint32_t tile[kTileWords];
  lut_stage_cell_body<STAGE, OC, PAIRED>(tabs, in, ia, H, W, y0, y1, out, blockIdx.x, blockIdx.y, blockIdx.z, tile);
}

#ifdef LERF_EXPERIMENTS
template <unsigned MTMASK, int MINB>
__global__ void __launch_bounds__(kTX* kTY, MINB)
    lut_stage2_mix_kernel(mix::MixTables t, const uint8_t* __restrict__ feat, int H, int W, int y0, int y1,
                          uint8_t* __restrict__ out) {
  __shared__ __align__(16) unsigned char smem[mix::kSmemBytes];
  mix::lut_stage2_mix_body<MTMASK>(t, feat, H, W, y0, y1, out, blockIdx.x, blockIdx.y, blockIdx.z, smem);
}

template <int NJ, int MINB, int LD, unsigned TABMASK = 0x3Fu, typename PX = uint2>
__global__ void __launch_bounds__(256, MINB)
    lut_stage2_mt_kernel(mt::MtTables t, const uint8_t* __restrict__ feat, int H, int W, int y0, int y1,
                         uint8_t* __restrict__ out) {
  __shared__ PX tile[(8 * NJ + 2 * mt::kHalo) * mt::kPitch];
  mt::lut_stage2_mt_body<NJ, LD, TABMASK, PX>(t, feat, H, W, y0, y1, out, blockIdx.x, blockIdx.y, blockIdx.z, tile);
}
#endif

}  // namespace

#ifdef LERF_EXPERIMENTS
// Stage 2, oC = 3, max-tap block tables (lut_mt.cuh).  variant: tile height / register budget.
int launch_stage2_mt(const lerf_luts_impl* L, const uint8_t* feat, int planes, int H, int W, int y0, int y1, uint8_t* out,
                     int variant, cudaStream_t st) {
  if (!L->mt2[0]) return fail(LERF_EUNSUPPORTED, "max-tap block tables were not built for this LUT set");
  mt::MtTables t;
  for (int i = 0; i < 6; ++i) t.t[i] = L->mt2[i];
#define LERF_GO(NJ, B, LD)                                                                  \
  {                                                                                         \
    dim3 grid((W + mt::kTX - 1) / mt::kTX, (y1 - y0 + 8 * NJ - 1) / (8 * NJ), planes);      \
    lut_stage2_mt_kernel<NJ, B, LD><<<grid, 256, 0, st>>>(t, feat, H, W, y0, y1, out);      \
  }
  switch (variant) {
    case 0: LERF_GO(1, 4, 0) break;
    case 1: LERF_GO(1, 4, 1) break;
    case 2: LERF_GO(1, 4, 2) break;
    case 3: LERF_GO(1, 4, 3) break;
    case 4: LERF_GO(1, 5, 1) break;
    case 5: LERF_GO(1, 3, 1) break;
    case 6: LERF_GO(2, 4, 1) break;
    case 7: LERF_GO(4, 3, 1) break;
    case 8: LERF_GO(4, 3, 0) break;
#define LERF_GO1(NJ, B, LD)                                                                                  \
  {                                                                                                          \
    dim3 grid((W + mt::kTX - 1) / mt::kTX, (y1 - y0 + 8 * NJ - 1) / (8 * NJ), planes);                       \
    lut_stage2_mt_kernel<NJ, B, LD, 0x3Fu, uint32_t><<<grid, 256, 0, st>>>(t, feat, H, W, y0, y1, out);     \
  }
    case 9: LERF_GO1(1, 4, 1) break;   // single-word taps (prepare1)
    case 10: LERF_GO1(1, 5, 1) break;
    case 11: LERF_GO1(1, 6, 1) break;
    case 12: LERF_GO1(2, 4, 1) break;
#undef LERF_GO1
    default: return fail(LERF_EINVAL, "unknown stage-2 max-tap variant %d", variant);
  }
#undef LERF_GO
  LERF_LAUNCHED();
  return LERF_OK;
}

// Stage 2, oC = 3, table-format mix (lut_mix.cuh).  variant selects the pass subset that uses the max-tap blocks.
int launch_stage2_mix(const lerf_luts_impl* L, const uint8_t* feat, int planes, int H, int W, int y0, int y1, uint8_t* out,
                      int variant, cudaStream_t st) {
  if (!L->mt2[0]) return fail(LERF_EUNSUPPORTED, "max-tap block tables were not built for this LUT set");
  mix::MixTables t;
  for (int i = 0; i < 6; ++i) { t.r[i] = L->s2[i]; t.m[i] = L->mt2[i]; }
  dim3 block(kTX * kTY), grid((W + kTX - 1) / kTX, (y1 - y0 + kTY - 1) / kTY, planes);
#define LERF_GO(MASK, B) lut_stage2_mix_kernel<MASK, B><<<grid, block, 0, st>>>(t, feat, H, W, y0, y1, out)
  switch (variant) {  // mask bit = mode * 4 + rotation; rotations 0,2 share table r0, rotations 1,3 share table r1
    case 0: LERF_GO(0x00Fu, 4); break;   // mode s                      (4 passes on max-tap blocks)
    case 1: LERF_GO(0x05Fu, 4); break;   // s + c rotations 0,2         (6)
    case 2: LERF_GO(0x0FFu, 4); break;   // s + c                       (8)
    case 3: LERF_GO(0x5FFu, 4); break;   // s + c + t rotations 0,2     (10)
    case 4: LERF_GO(0x005u, 4); break;   // s rotations 0,2             (2)
    case 5: LERF_GO(0x05Fu, 3); break;
    case 6: LERF_GO(0x0FFu, 3); break;
    case 7: LERF_GO(0x05Fu, 5); break;
    case 8: LERF_GO(0x0FFu, 5); break;
    case 9: LERF_GO(0xFF0u, 4); break;   // c + t                       (8)
    default: return fail(LERF_EINVAL, "unknown stage-2 mix variant %d", variant);
  }
#undef LERF_GO
  LERF_LAUNCHED();
  return LERF_OK;
}

#endif  // LERF_EXPERIMENTS

// Builds the cell-packed copies of the nine tables inside one device allocation (called by lerf_luts_create).
int build_cell_tables(lerf_luts_impl* L, const int8_t* const host_tables[9]) {
  const int oC = L->oC2;
  const size_t s1_bytes = (size_t)65536 * 16;
  const size_t s2_stride = oC == 3 ? 48 : 16;
  const cell::Hash h{(uint32_t)L->cell_hash[0], (uint32_t)L->cell_hash[1], (uint32_t)L->cell_hash[2]};
  const size_t s2_bytes = (size_t)65536 * s2_stride;
  const size_t total = 3 * s1_bytes + 6 * s2_bytes;
  std::vector<uint8_t> host(total, 0);
  const int ident[4] = {0, 1, 2, 3};
  for (int i = 0; i < 3; ++i) cell::repack_cells(host_tables[i], 1, ident, h, host.data() + i * s1_bytes, 16, 0);
  for (int i = 0; i < 6; ++i) cell::repack_cells(host_tables[3 + i], oC, ident, h, host.data() + 3 * s1_bytes + i * s2_bytes, s2_stride, 0);
  cudaError_t e = cudaMalloc(&L->cell_block, total);
  if (e != cudaSuccess) return fail(LERF_ENOMEM, "cudaMalloc(%zu) for the cell-packed LUT block failed: %s", total, cudaGetErrorString(e));
  e = cudaMemcpy(L->cell_block, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(LERF_ECUDA, "cell-packed LUT upload failed: %s", cudaGetErrorString(e));
  L->cell_block_bytes = total;
#ifdef LERF_EXPERIMENTS
  if (oC == 3) {  // max-tap block tables for stage 2 (lut_mt.cuh)
    std::vector<uint8_t> mtab(mt::kTableBytes);
    e = cudaMalloc(&L->mt_block, 6 * mt::kTableBytes);
    if (e != cudaSuccess) return fail(LERF_ENOMEM, "cudaMalloc for the max-tap LUT block failed: %s", cudaGetErrorString(e));
    for (int i = 0; i < 6; ++i) {
      mt::repack_maxtap(host_tables[3 + i], mtab.data());
      e = cudaMemcpy((uint8_t*)L->mt_block + i * mt::kTableBytes, mtab.data(), mt::kTableBytes, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return fail(LERF_ECUDA, "max-tap LUT upload failed: %s", cudaGetErrorString(e));
      L->mt2[i] = (const uint8_t*)L->mt_block + i * mt::kTableBytes;
    }
  }
#endif
  for (int i = 0; i < 3; ++i) L->c1[i] = (const uint8_t*)L->cell_block + i * s1_bytes;
  for (int i = 0; i < 6; ++i) L->c2[i] = (const uint8_t*)L->cell_block + 3 * s1_bytes + i * s2_bytes;
  return LERF_OK;
}

int launch_stage_cell(const lerf_luts_impl* L, int stage, const uint8_t* in, const InAddr& ia, int planes, int H, int W,
                      int y0, int y1, uint8_t* out, int variant, cudaStream_t st) {
  CellTables t;
  for (int i = 0; i < 6; ++i) t.t[i] = stage == 1 ? (i < 3 ? L->c1[i] : nullptr) : L->c2[i];
  t.h = cell::Hash{(uint32_t)L->cell_hash[0], (uint32_t)L->cell_hash[1], (uint32_t)L->cell_hash[2]};
  dim3 block(kTX * kTY), grid((W + kTX - 1) / kTX, (y1 - y0 + kTY - 1) / kTY, planes);
  if (g_dbg.carveout >= 0) {  // A/B hook; production leaves the driver's choice (the smallest carve-out that fits the blocks)
    cudaFuncSetAttribute(lut_stage_cell_kernel<1, 1, 5, true>, cudaFuncAttributePreferredSharedMemoryCarveout, g_dbg.carveout);
    cudaFuncSetAttribute(lut_stage_cell_kernel<2, 1, 4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, g_dbg.carveout);
  }
#define LERF_GO(S, O, B) lut_stage_cell_kernel<S, O, B><<<grid, block, 0, st>>>(t, in, ia, H, W, y0, y1, out)
#define LERF_GP(S, B) lut_stage_cell_kernel<S, 1, B, true><<<grid, block, 0, st>>>(t, in, ia, H, W, y0, y1, out)
  if (variant == 7 && (stage == 1 || L->oC2 == 1)) {  // one sort per lookup (the r1 form): second implementation for the tests
    if (stage == 1) LERF_GO(1, 1, 6);
    else LERF_GO(2, 1, 4);
    LERF_LAUNCHED();
    return LERF_OK;
  }
#ifdef LERF_EXPERIMENTS
  if (stage == 1) {
    switch (variant) {
      case 2: LERF_GO(1, 1, 2); break;
      case 3: LERF_GO(1, 1, 3); break;
      case 5: LERF_GO(1, 1, 5); break;
      case 4: LERF_GO(1, 1, 4); break;
      case 8: LERF_GP(1, 6); break;
      case 9: LERF_GP(1, 4); break;
      default: LERF_GP(1, 5);
    }
  } else if (L->oC2 == 3) {
    switch (variant) {
      case 2: LERF_GO(2, 3, 2); break;
      case 3: LERF_GO(2, 3, 3); break;
      case 5: LERF_GO(2, 3, 5); break;
      default: LERF_GO(2, 3, 4);
    }
  } else {
    LERF_GP(2, 4);
  }
#else
  if (stage == 1) LERF_GP(1, 5);
  else if (L->oC2 == 3) LERF_GO(2, 3, 4);
  else LERF_GP(2, 4);
#endif
#undef LERF_GP
#undef LERF_GO
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // namespace lerf
