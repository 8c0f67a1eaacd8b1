// Rotation-ensembled LUT stages on cell-packed tables (sm_100a).  See lut_cell.cuh for the lookup primitive.
//
// Reference being replaced (ddlee-cn/LeRF-PyTorch):
//   FourSimplexInterpFaster            resample/eval_lut_sr.py:24-470
//   stage-1 / stage-2 ensembling loops resample/eval_lut_sr.py:541-628 (= eval_lut_warp.py:104-191)
//
// Kernel `lut_stage_cell_kernel`: one thread = one sample; the tile (+3 halo) is kept in shared memory as
// pre-split words (lsb | msb, cell::split_px); each of the 12 passes is one 5-compare-exchange sort, one
// 128-bit table load (oC = 1) or one 256-bit + one 128-bit load (oC = 3), two PRMT + two DP4A per channel.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "lut_cell.cuh"

namespace lerf {

using cell::Simplex;

namespace {

constexpr int kHalo = 3;
constexpr int kTX = 32, kTY = 8;
constexpr int kPitch = 40;  // words; 40 mod 32 = 8: the four 8-word rows of a warp's 8x4 patch hit disjoint banks

template <int MODE, int R, int K>
struct Tap {  // mode pattern (eval_lut_sr.py:30-81) composed with the rotation
  static constexpr int di = MODE == 0 ? (K >> 1) : (MODE == 1 ? 0 : K);
  static constexpr int dj = MODE == 0 ? (K & 1) : K;
  static constexpr int dy = R == 0 ? di : (R == 1 ? dj : (R == 2 ? -di : -dj));
  static constexpr int dx = R == 0 ? dj : (R == 1 ? -di : (R == 2 ? -dj : di));
};

template <int MODE, int R>
__device__ __forceinline__ Simplex simplex_at(const uint32_t* c) {
  return cell::simplex_of(c[Tap<MODE, R, 0>::dy * kPitch + Tap<MODE, R, 0>::dx],
                          c[Tap<MODE, R, 1>::dy * kPitch + Tap<MODE, R, 1>::dx],
                          c[Tap<MODE, R, 2>::dy * kPitch + Tap<MODE, R, 2>::dx],
                          c[Tap<MODE, R, 3>::dy * kPitch + Tap<MODE, R, 3>::dx]);
}

__device__ __forceinline__ int lookup1(const uint8_t* __restrict__ tab, const Simplex& s) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(tab) + s.cell);
  return cell::blend(q.x, q.y, q.z, q.w, s);
}

// oC = 3: 64-byte cells, channel k at bytes [16k, 16k+16)
__device__ __forceinline__ void lookup3(const uint8_t* __restrict__ tab, const Simplex& s, int& n0, int& n1, int& n2) {
  const uint8_t* p = tab + (size_t)s.cell * 64;
  uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
      : "l"(p));
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(p + 32));
  n0 += cell::blend(a0, a1, a2, a3, s);
  n1 += cell::blend(b0, b1, b2, b3, s);
  n2 += cell::blend(c.x, c.y, c.z, c.w, s);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ int rhe_div(int num, int den) {  // round_half_even(num / den), num > 0, den even
  const int t = num + den / 2;
  int q = t / den;
  if (t - q * den == 0 && (q & 1)) --q;
  return q;
}

struct CellTables {
  const uint8_t* t[6];
};

template <int STAGE, int OC, int MINB>
__global__ void __launch_bounds__(kTX* kTY, MINB)
    lut_stage_cell_kernel(CellTables tabs, const uint8_t* __restrict__ in, InAddr ia, int H, int W, int y0, int y1,
                          uint8_t* __restrict__ out) {
  __shared__ uint32_t tile[(kTY + 2 * kHalo) * kPitch];
  const int p = blockIdx.z;
  const int bx = blockIdx.x * kTX, by = y0 + blockIdx.y * kTY;
  const uint8_t* src = in + (long long)(p / ia.channels) * ia.batch_stride + (long long)(p % ia.channels) * ia.chan_stride;
  const int tid = threadIdx.y * kTX + threadIdx.x;
  for (int i = tid; i < (kTY + 2 * kHalo) * (kTX + 2 * kHalo); i += kTX * kTY) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = clampi(by + r - kHalo, 0, H - 1), gx = clampi(bx + c - kHalo, 0, W - 1);
    tile[r * kPitch + c] = cell::split_px(__ldg(src + (long long)gy * ia.row_stride + (long long)gx * ia.pix_stride));
  }
  __syncthreads();
  // a warp covers an 8x4 pixel patch: 2-D neighbours have closer values than the ends of a 32-pixel row, so the 32
  // cells of one load fall into fewer cache lines
  const int lane = tid & 31, wrp = tid >> 5;
  const int tx = (wrp & 3) * 8 + (lane & 7), ty = (wrp >> 2) * 4 + (lane >> 3);
  const int x = bx + tx, y = by + ty;
  if (x >= W || y >= y1) return;
  const uint32_t* c = tile + (ty + kHalo) * kPitch + tx + kHalo;

  if (OC == 1) {
    int n = 0;
#define LERF_L1(M, R) n += lookup1(tabs.t[STAGE == 1 ? M : 2 * M + (R & 1)], simplex_at<M, R>(c));
    LERF_L1(0, 0) LERF_L1(0, 1) LERF_L1(0, 2) LERF_L1(0, 3)
    LERF_L1(1, 0) LERF_L1(1, 1) LERF_L1(1, 2) LERF_L1(1, 3)
    LERF_L1(2, 0) LERF_L1(2, 1) LERF_L1(2, 2) LERF_L1(2, 3)
#undef LERF_L1
    int v;
    if (STAGE == 1) {
      v = n <= 0 ? 0 : min(rhe_div(n, 48), 255);
    } else {
      const int t = n + 127 * 192;
      v = t <= 0 ? 0 : min(rhe_div(t, 192), 255);
    }
    out[((long long)p * H + y) * W + x] = (uint8_t)v;
  } else {
    int n0 = 0, n1 = 0, n2 = 0;
#define LERF_L3(M, R) lookup3(tabs.t[2 * M + (R & 1)], simplex_at<M, R>(c), n0, n1, n2);
    LERF_L3(0, 0) LERF_L3(0, 1) LERF_L3(0, 2) LERF_L3(0, 3)
    LERF_L3(1, 0) LERF_L3(1, 1) LERF_L3(1, 2) LERF_L3(1, 3)
    LERF_L3(2, 0) LERF_L3(2, 1) LERF_L3(2, 2) LERF_L3(2, 3)
#undef LERF_L3
    const long long o = ((long long)p * 3 * H + y) * W + x, ps = (long long)H * W;
    const int t0 = n0 + 127 * 192, t1 = n1 + 127 * 192, t2 = n2 + 127 * 192;
    out[o] = (uint8_t)(t0 <= 0 ? 0 : min(rhe_div(t0, 192), 255));
    out[o + ps] = (uint8_t)(t1 <= 0 ? 0 : min(rhe_div(t1, 192), 255));
    out[o + 2 * ps] = (uint8_t)(t2 <= 0 ? 0 : min(rhe_div(t2, 192), 255));
  }
}

}  // namespace

// Builds the cell-packed copies of the nine tables inside one device allocation (called by lerf_luts_create).
int build_cell_tables(lerf_luts_impl* L, const int8_t* const host_tables[9]) {
  const int oC = L->oC2;
  const size_t s1_bytes = (size_t)65536 * 16;
  const size_t s2_stride = oC == 3 ? 64 : 16;
  const size_t s2_bytes = (size_t)65536 * s2_stride;
  const size_t total = 3 * s1_bytes + 6 * s2_bytes;
  std::vector<uint8_t> host(total, 0);
  const int ident[4] = {0, 1, 2, 3};
  for (int i = 0; i < 3; ++i) cell::repack_cells(host_tables[i], 1, ident, host.data() + i * s1_bytes, 16, 0);
  for (int i = 0; i < 6; ++i) cell::repack_cells(host_tables[3 + i], oC, ident, host.data() + 3 * s1_bytes + i * s2_bytes, s2_stride, 0);
  cudaError_t e = cudaMalloc(&L->cell_block, total);
  if (e != cudaSuccess) return fail(LERF_ENOMEM, "cudaMalloc(%zu) for the cell-packed LUT block failed: %s", total, cudaGetErrorString(e));
  e = cudaMemcpy(L->cell_block, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(LERF_ECUDA, "cell-packed LUT upload failed: %s", cudaGetErrorString(e));
  L->cell_block_bytes = total;
  for (int i = 0; i < 3; ++i) L->c1[i] = (const uint8_t*)L->cell_block + i * s1_bytes;
  for (int i = 0; i < 6; ++i) L->c2[i] = (const uint8_t*)L->cell_block + 3 * s1_bytes + i * s2_bytes;
  return LERF_OK;
}

int launch_stage_cell(const lerf_luts_impl* L, int stage, const uint8_t* in, const InAddr& ia, int planes, int H, int W,
                      int y0, int y1, uint8_t* out, int variant, cudaStream_t st) {
  CellTables t;
  for (int i = 0; i < 6; ++i) t.t[i] = stage == 1 ? (i < 3 ? L->c1[i] : nullptr) : L->c2[i];
  dim3 block(kTX, kTY), grid((W + kTX - 1) / kTX, (y1 - y0 + kTY - 1) / kTY, planes);
#define LERF_GO(S, O, B) lut_stage_cell_kernel<S, O, B><<<grid, block, 0, st>>>(t, in, ia, H, W, y0, y1, out)
  if (stage == 1) {
    switch (variant) {
      case 2: LERF_GO(1, 1, 2); break;
      case 3: LERF_GO(1, 1, 3); break;
      case 5: LERF_GO(1, 1, 5); break;
      case 6: LERF_GO(1, 1, 6); break;
      default: LERF_GO(1, 1, 4);
    }
  } else if (L->oC2 == 3) {
    switch (variant) {
      case 2: LERF_GO(2, 3, 2); break;
      case 3: LERF_GO(2, 3, 3); break;
      case 5: LERF_GO(2, 3, 5); break;
      default: LERF_GO(2, 3, 4);
    }
  } else {
    LERF_GO(2, 1, 4);
  }
#undef LERF_GO
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // namespace lerf
