// LUT half of the LeRF hot path for sm_100a: 4-simplex interpolation over the 17^4 sampling grid
// with rotation ensembling.  All arithmetic is exact integer.
//
// Reference being replaced (ddlee-cn/LeRF-PyTorch):
//   FourSimplexInterpFaster            resample/eval_lut_sr.py:24-470
//   stage-1 / stage-2 ensembling loops resample/eval_lut_sr.py:541-628 (= eval_lut_warp.py:104-191)
//   LUT loader                         resample/eval_lut_sr.py:750-775
//
// Design (not a translation): the reference materialises 16 corner gathers and 24 boolean masks
// per pass and runs 24 passes over rotated copies.  Here one thread owns one sample, the four
// rotations become clamped constant offsets into a shared-memory tile (no rotated copies), the
// 24-way branch becomes a 5-compare-exchange sort of (lsb<<13 | stride) keys that walks the
// simplex from p0000 to p1111, and all 12 passes of a stage accumulate in registers, so the only
// HBM traffic is one uint8 read and one (stage 1) or oC (stage 2) uint8 writes per sample.
#include <stdarg.h>

#include <vector>

#include "common.cuh"

namespace lerf {

thread_local std::string g_last_error;
thread_local long long g_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// ---------------------------------------------------------------------------------------------
// simplex walk
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap_desc(int& a, int& b) {
  const int hi = max(a, b);
  b = min(a, b);
  a = hi;
}

// Sorted-simplex form of the 24 cases of eval_lut_sr.py:218-462: sort taps by LSB,
// largest first; vertex k adds the stride of the k-th sorted tap; weights are the LSB gaps.  Ties
// have zero weight, so their order is irrelevant.
struct Simplex {
  int i0, i1, i2, i3, i4;  // table row of p0000 .. p1111
  int w0, w1, w2, w3, w4;  // 16-f1, f1-f2, f2-f3, f3-f4, f4   (sum = 16)
};

__device__ __forceinline__ Simplex simplex_of(int va, int vb, int vc, int vd) {
  Simplex s;
  s.i0 = (((va >> 4) * kL + (vb >> 4)) * kL + (vc >> 4)) * kL + (vd >> 4);
  int ka = ((va & 15) << 13) | kStrideA;
  int kb = ((vb & 15) << 13) | kStrideB;
  int kc = ((vc & 15) << 13) | kStrideC;
  int kd = ((vd & 15) << 13) | 1;
  cswap_desc(ka, kb);
  cswap_desc(kc, kd);
  cswap_desc(ka, kc);
  cswap_desc(kb, kd);
  cswap_desc(kb, kc);
  const int f1 = ka >> 13, f2 = kb >> 13, f3 = kc >> 13, f4 = kd >> 13;
  s.i1 = s.i0 + (ka & 8191);
  s.i2 = s.i1 + (kb & 8191);
  s.i3 = s.i2 + (kc & 8191);
  s.i4 = s.i0 + kStrideAll;
  s.w0 = 16 - f1;
  s.w1 = f1 - f2;
  s.w2 = f2 - f3;
  s.w3 = f3 - f4;
  s.w4 = f4;
  return s;
}

__device__ __forceinline__ int blend1(const int8_t* __restrict__ t, const Simplex& s) {
  return s.w0 * (int)__ldg(t + s.i0) + s.w1 * (int)__ldg(t + s.i1) + s.w2 * (int)__ldg(t + s.i2) +
         s.w3 * (int)__ldg(t + s.i3) + s.w4 * (int)__ldg(t + s.i4);
}

// Roofline experiments only (wrong results by construction): EXP 1 = every lane reads row 0 (loads issued, no
// address divergence); EXP 2 = no table load at all (the index stands in for the value).
template <int EXP>
__device__ __forceinline__ int blend1x(const int8_t* __restrict__ t, const Simplex& s) {
  if (EXP == 1)
    return s.w0 * (int)__ldg(t + (s.i0 & 1)) + s.w1 * (int)__ldg(t + (s.i1 & 1)) + s.w2 * (int)__ldg(t + (s.i2 & 1)) +
           s.w3 * (int)__ldg(t + (s.i3 & 1)) + s.w4 * (int)__ldg(t + (s.i4 & 1));
  return s.w0 * s.i0 + s.w1 * s.i1 + s.w2 * s.i2 + s.w3 * s.i3 + s.w4 * s.i4;
}

// oC = 3 tables are repacked to one uint32 per row, bytes (c0, c1, c2, 0): one 32-bit load per
// vertex and three dp4a per vertex with the weight placed in the byte lane of the wanted channel.
__device__ __forceinline__ void blend3(const uint32_t* __restrict__ t, const Simplex& s, int& n0,
                                       int& n1, int& n2) {
  const int e0 = (int)__ldg(t + s.i0), e1 = (int)__ldg(t + s.i1), e2 = (int)__ldg(t + s.i2),
            e3 = (int)__ldg(t + s.i3), e4 = (int)__ldg(t + s.i4);
#define LERF_ACC3(e, w)              \
  n0 = __dp4a(e, (w), n0);           \
  n1 = __dp4a(e, (w) << 8, n1);      \
  n2 = __dp4a(e, (w) << 16, n2);
  LERF_ACC3(e0, s.w0)
  LERF_ACC3(e1, s.w1)
  LERF_ACC3(e2, s.w2)
  LERF_ACC3(e3, s.w3)
  LERF_ACC3(e4, s.w4)
#undef LERF_ACC3
}

// ---------------------------------------------------------------------------------------------
// tap geometry: mode pattern (eval_lut_sr.py:30-81) composed with the rotation
// ---------------------------------------------------------------------------------------------
// MODE 0 = 's', 1 = 'c', 2 = 't'.  (di, dj) is the tap offset in the rotated frame; rotating the
// image by r quarter turns, edge-padding bottom/right and un-rotating the result is the same as
// reading the un-rotated image at the offsets below, clamped to the image.
template <int MODE, int R, int K>
struct Tap {
  static constexpr int di = MODE == 0 ? (K >> 1) : (MODE == 1 ? 0 : K);
  static constexpr int dj = MODE == 0 ? (K & 1) : K;
  static constexpr int dy = R == 0 ? di : (R == 1 ? dj : (R == 2 ? -di : -dj));
  static constexpr int dx = R == 0 ? dj : (R == 1 ? -di : (R == 2 ? -dj : di));
};

constexpr int kHalo = 3;  // reach of modes c and t
constexpr int kTX = 32, kTY = 8;
constexpr int kPitch = kTX + 2 * kHalo + 2;  // 40: rows of the tile, bytes

template <int MODE, int R>
__device__ __forceinline__ Simplex simplex_at(const uint8_t* c) {
  const int va = c[Tap<MODE, R, 0>::dy * kPitch + Tap<MODE, R, 0>::dx];
  const int vb = c[Tap<MODE, R, 1>::dy * kPitch + Tap<MODE, R, 1>::dx];
  const int vc = c[Tap<MODE, R, 2>::dy * kPitch + Tap<MODE, R, 2>::dx];
  const int vd = c[Tap<MODE, R, 3>::dy * kPitch + Tap<MODE, R, 3>::dx];
  return simplex_of(va, vb, vc, vd);
}

struct StageTables {
  const void* t[6];
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// round_half_even(num / den) for num > 0, den even
__device__ __forceinline__ int rhe_div(int num, int den) {
  const int t = num + den / 2;
  int q = t / den;
  if (t - q * den == 0 && (q & 1)) --q;  // exact .5 -> even
  return q;
}

// One rotation-ensembled stage.  STAGE 1: 3 tables (s,c,t) used for all four rotations, output
// feat = clip(rhe(sum/48)).  STAGE 2: 6 tables ([mode][r&1]), oC outputs, code = clip(rhe(sum/192 + 127)).
template <int STAGE, int OC, int MINB>
__global__ void __launch_bounds__(kTX* kTY, (MINB >= 11 ? 3 : MINB))
    lut_stage_kernel(StageTables tabs, const uint8_t* __restrict__ in, InAddr ia, int H, int W, int y0,
                     int y1, uint8_t* __restrict__ out) {
  __shared__ uint8_t tile[(kTY + 2 * kHalo) * kPitch];
  const int p = blockIdx.z;
  const int bx = blockIdx.x * kTX, by = y0 + blockIdx.y * kTY;
  const uint8_t* src = in + (long long)(p / ia.channels) * ia.batch_stride +
                       (long long)(p % ia.channels) * ia.chan_stride;
  const int tid = threadIdx.y * kTX + threadIdx.x;
  for (int i = tid; i < (kTY + 2 * kHalo) * (kTX + 2 * kHalo); i += kTX * kTY) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = clampi(by + r - kHalo, 0, H - 1), gx = clampi(bx + c - kHalo, 0, W - 1);
    tile[r * kPitch + c] = __ldg(src + (long long)gy * ia.row_stride + (long long)gx * ia.pix_stride);
  }
  __syncthreads();
  // A warp covers an 8x4 pixel patch, not a 32x1 row: neighbours in 2-D have closer values than the ends of a
  // 32-pixel row, so the 32 table addresses of a gather fall into fewer cache lines (measured).
  const int lane = tid & 31, wrp = tid >> 5;
  const int tx = (wrp & 3) * 8 + (lane & 7), ty = (wrp >> 2) * 4 + (lane >> 3);
  const int x = bx + tx, y = by + ty;
  if (x >= W || y >= y1) return;
  const uint8_t* c = tile + (ty + kHalo) * kPitch + tx + kHalo;

  if (STAGE == 1) {
    int n = 0;
#define LERF_S1(M, R)                                                                  \
  n += (MINB >= 11 ? blend1x<MINB - 10>((const int8_t*)tabs.t[M], simplex_at<M, R>(c)) \
                   : blend1((const int8_t*)tabs.t[M], simplex_at<M, R>(c)));
    LERF_S1(0, 0) LERF_S1(0, 1) LERF_S1(0, 2) LERF_S1(0, 3)
    LERF_S1(1, 0) LERF_S1(1, 1) LERF_S1(1, 2) LERF_S1(1, 3)
    LERF_S1(2, 0) LERF_S1(2, 1) LERF_S1(2, 2) LERF_S1(2, 3)
#undef LERF_S1
    const int v = n <= 0 ? 0 : min(rhe_div(n, 48), 255);
    out[((long long)p * H + y) * W + x] = (uint8_t)v;
  } else if (OC == 1) {
    int n = 0;
#define LERF_S2(M, R) n += blend1((const int8_t*)tabs.t[2 * M + (R & 1)], simplex_at<M, R>(c));
    LERF_S2(0, 0) LERF_S2(0, 1) LERF_S2(0, 2) LERF_S2(0, 3)
    LERF_S2(1, 0) LERF_S2(1, 1) LERF_S2(1, 2) LERF_S2(1, 3)
    LERF_S2(2, 0) LERF_S2(2, 1) LERF_S2(2, 2) LERF_S2(2, 3)
#undef LERF_S2
    const int t = n + 127 * 192;
    out[((long long)p * H + y) * W + x] = (uint8_t)(t <= 0 ? 0 : min(rhe_div(t, 192), 255));
  } else {
    int n0 = 0, n1 = 0, n2 = 0;
#define LERF_S2(M, R) blend3((const uint32_t*)tabs.t[2 * M + (R & 1)], simplex_at<M, R>(c), n0, n1, n2);
    LERF_S2(0, 0) LERF_S2(0, 1) LERF_S2(0, 2) LERF_S2(0, 3)
    LERF_S2(1, 0) LERF_S2(1, 1) LERF_S2(1, 2) LERF_S2(1, 3)
    LERF_S2(2, 0) LERF_S2(2, 1) LERF_S2(2, 2) LERF_S2(2, 3)
#undef LERF_S2
    const long long o = ((long long)p * 3 * H + y) * W + x, ps = (long long)H * W;
    const int t0 = n0 + 127 * 192, t1 = n1 + 127 * 192, t2 = n2 + 127 * 192;
    out[o] = (uint8_t)(t0 <= 0 ? 0 : min(rhe_div(t0, 192), 255));
    out[o + ps] = (uint8_t)(t1 <= 0 ? 0 : min(rhe_div(t1, 192), 255));
    out[o + 2 * ps] = (uint8_t)(t2 <= 0 ? 0 : min(rhe_div(t2, 192), 255));
  }
}

// Generic single pass (any of the five modes, any oC): the drop-in for one call of
// FourSimplexInterpFaster.  Slow path by design -- the product path uses the stage kernels.
struct PassTaps {
  int di[4], dj[4];
};

__global__ void lut_pass_kernel(const int8_t* __restrict__ tab, const uint8_t* __restrict__ img, int C,
                                int h, int w, int hp, int wp, PassTaps tp, int oC, int32_t* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int c = blockIdx.z;
  if (i >= h || j >= w) return;
  const uint8_t* pl = img + (long long)c * hp * wp;
  const int va = pl[(i + tp.di[0]) * wp + j + tp.dj[0]];
  const int vb = pl[(i + tp.di[1]) * wp + j + tp.dj[1]];
  const int vc = pl[(i + tp.di[2]) * wp + j + tp.dj[2]];
  const int vd = pl[(i + tp.di[3]) * wp + j + tp.dj[3]];
  const Simplex s = simplex_of(va, vb, vc, vd);
  for (int o = 0; o < oC; ++o) {
    const int n = s.w0 * (int)tab[(long long)s.i0 * oC + o] + s.w1 * (int)tab[(long long)s.i1 * oC + o] +
                  s.w2 * (int)tab[(long long)s.i2 * oC + o] + s.w3 * (int)tab[(long long)s.i3 * oC + o] +
                  s.w4 * (int)tab[(long long)s.i4 * oC + o];
    out[((long long)(c * oC + o) * h + i) * w + j] = n;
  }
}

static bool mode_taps(char mode, PassTaps& t, int& pad) {
  static const int S[2][4] = {{0, 0, 1, 1}, {0, 1, 0, 1}};
  static const int D[2][4] = {{0, 0, 2, 2}, {0, 2, 0, 2}};
  static const int Y[2][4] = {{0, 1, 1, 2}, {0, 1, 2, 1}};
  static const int Cm[2][4] = {{0, 0, 0, 0}, {0, 1, 2, 3}};
  static const int T[2][4] = {{0, 1, 2, 3}, {0, 1, 2, 3}};
  const int(*m)[4];
  switch (mode) {  // eval_lut_sr.py:12-18, :30-81
    case 's': m = S; pad = 1; break;
    case 'd': m = D; pad = 2; break;
    case 'y': m = Y; pad = 2; break;
    case 'c': m = Cm; pad = 3; break;
    case 't': m = T; pad = 3; break;
    default: return false;
  }
  for (int k = 0; k < 4; ++k) { t.di[k] = m[0][k]; t.dj[k] = m[1][k]; }
  return true;
}

}  // namespace lerf

using namespace lerf;

static int g_lut_variant[2] = {0, 0};  // per stage: 0 = production, 1..19 legacy row-major kernel, 20+ cell kernel

extern "C" {

int lerf_abi_version(void) { return LERF_ABI_VERSION; }
const char* lerf_last_error_string(void) { return g_last_error.c_str(); }
long long lerf_launch_count(void) { return g_launches; }
void lerf_launch_count_reset(void) { g_launches = 0; }

int lerf_luts_create(const int8_t* const host_tables[9], int oC2, int device, lerf_luts_t** out) {
  if (!host_tables || !out) return fail(LERF_EINVAL, "lerf_luts_create: null argument");
  if (oC2 != 1 && oC2 != 3) return fail(LERF_EINVAL, "lerf_luts_create: oC must be 1 or 3, got %d", oC2);
  for (int i = 0; i < 9; ++i)
    if (!host_tables[i]) return fail(LERF_EINVAL, "lerf_luts_create: table %d is null", i);
  LERF_CUDA(cudaSetDevice(device));
  const size_t s1_bytes = (kEntries + 255) / 256 * 256;  // 256-B aligned slots
  const size_t s2_entry = oC2 == 3 ? 4 : 1;
  const size_t s2_bytes = (kEntries * s2_entry + 255) / 256 * 256;
  const size_t total = 3 * s1_bytes + 6 * s2_bytes;
  std::vector<uint8_t> host(total, 0);
  for (int i = 0; i < 3; ++i) memcpy(host.data() + i * s1_bytes, host_tables[i], kEntries);
  for (int i = 0; i < 6; ++i) {
    uint8_t* dst = host.data() + 3 * s1_bytes + i * s2_bytes;
    const int8_t* src = host_tables[3 + i];
    if (oC2 == 3) {
      for (int e = 0; e < kEntries; ++e) {
        dst[4 * e + 0] = (uint8_t)src[3 * e + 0];
        dst[4 * e + 1] = (uint8_t)src[3 * e + 1];
        dst[4 * e + 2] = (uint8_t)src[3 * e + 2];
        dst[4 * e + 3] = 0;
      }
    } else {
      memcpy(dst, src, kEntries);
    }
  }
  lerf_luts_impl* L = new lerf_luts_impl();
  memset(L, 0, sizeof(*L));
  L->device = device;
  L->oC2 = oC2;
  L->block_bytes = total;
  L->num_sms = 148;
  cudaDeviceGetAttribute(&L->num_sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = cudaMalloc(&L->block, total);
  if (e != cudaSuccess) {
    delete L;
    return fail(LERF_ENOMEM, "cudaMalloc(%zu) for the LUT block failed: %s", total, cudaGetErrorString(e));
  }
  e = cudaMemcpy(L->block, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(L->block);
    delete L;
    return fail(LERF_ECUDA, "LUT upload failed: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < 3; ++i) L->s1[i] = (const int8_t*)((uint8_t*)L->block + i * s1_bytes);
  for (int i = 0; i < 6; ++i) L->s2[i] = (uint8_t*)L->block + 3 * s1_bytes + i * s2_bytes;
  int rc = build_cell_tables(L, host_tables);
  if (rc) {
    cudaFree(L->cell_block);
    cudaFree(L->block);
    delete L;
    return rc;
  }
  // Reserve persisting L2 for the tables (best effort; the window itself is per stream).
  size_t want = L->cell_block_bytes;
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  cudaGetLastError();
  *out = reinterpret_cast<lerf_luts_t*>(L);
  return LERF_OK;
}

void lerf_luts_destroy(lerf_luts_t* luts) {
  if (!luts) return;
  lerf_luts_impl* L = reinterpret_cast<lerf_luts_impl*>(luts);
  cudaSetDevice(L->device);
  cudaFree(L->block);
  cudaFree(L->cell_block);
  delete L;
}

int lerf_luts_oc(const lerf_luts_t* luts) {
  return luts ? reinterpret_cast<const lerf_luts_impl*>(luts)->oC2 : 0;
}

int lerf_luts_pin_l2(const lerf_luts_t* luts, lerf_stream_t stream) {
  if (!luts) return fail(LERF_EINVAL, "lerf_luts_pin_l2: null handle");
  const lerf_luts_impl* L = reinterpret_cast<const lerf_luts_impl*>(luts);
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  attr.accessPolicyWindow.base_ptr = L->cell_block;  // the tables the production kernels read
  attr.accessPolicyWindow.num_bytes = L->cell_block_bytes;
  attr.accessPolicyWindow.hitRatio = 1.0f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  LERF_CUDA(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  return LERF_OK;
}

int lerf_lut_pass(const int8_t* table, const uint8_t* img, int C, int h, int w, char mode, int oC,
                  int32_t* out, lerf_stream_t stream) {
  PassTaps tp;
  int pad;
  if (!mode_taps(mode, tp, pad)) return fail(LERF_EINVAL, "Mode %c not implemented.", mode);
  if (!table || !img || !out) return fail(LERF_EINVAL, "lerf_lut_pass: null pointer");
  if (C < 0 || h < 0 || w < 0 || oC < 1) return fail(LERF_EINVAL, "lerf_lut_pass: bad sizes");
  if (C == 0 || h == 0 || w == 0) return LERF_OK;
  if (C > 65535) return fail(LERF_EINVAL, "lerf_lut_pass: more than 65535 planes");
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, C);
  lut_pass_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(table, img, C, h, w, h + pad, w + pad, tp, oC, out);
  LERF_LAUNCHED();
  return LERF_OK;
}

static int stage_args_ok(const char* who, const void* luts, const void* a, const void* b, int planes, int H,
                         int W, int y0, int y1) {
  if (!luts || !a || !b) return fail(LERF_EINVAL, "%s: null pointer", who);
  if (planes < 0 || H < 1 || W < 1) return fail(LERF_EINVAL, "%s: bad sizes planes=%d H=%d W=%d", who, planes, H, W);
  if (y0 < 0 || y1 > H || y0 > y1) return fail(LERF_EINVAL, "%s: bad row band [%d,%d) of %d", who, y0, y1, H);
  if (planes > 65535) return fail(LERF_EINVAL, "%s: more than 65535 planes per call", who);
  return LERF_OK;
}

int lerf_lut_stage1(const lerf_luts_t* luts, const uint8_t* in, int planes, int H, int W, int in_channels,
                    long long in_batch_stride, long long in_chan_stride, long long in_row_stride,
                    long long in_pix_stride, int y0, int y1, uint8_t* feat, lerf_stream_t stream) {
  int rc = stage_args_ok("lerf_lut_stage1", luts, in, feat, planes, H, W, y0, y1);
  if (rc) return rc;
  if (in_channels < 1) return fail(LERF_EINVAL, "lerf_lut_stage1: in_channels must be >= 1");
  if (planes == 0 || y0 == y1) return LERF_OK;
  const lerf_luts_impl* L = reinterpret_cast<const lerf_luts_impl*>(luts);
  StageTables t;
  for (int i = 0; i < 3; ++i) t.t[i] = L->s1[i];
  for (int i = 3; i < 6; ++i) t.t[i] = nullptr;
  InAddr ia{in_channels, in_batch_stride, in_chan_stride, in_row_stride, in_pix_stride};
  if (g_lut_variant[0] == 0 || g_lut_variant[0] >= 20)  // production: cell-packed tables (lut_cell.cu)
    return launch_stage_cell(L, 1, in, ia, planes, H, W, y0, y1, feat, g_lut_variant[0] >= 20 ? g_lut_variant[0] - 20 : 0,
                             (cudaStream_t)stream);
  dim3 block(kTX, kTY), grid((W + kTX - 1) / kTX, (y1 - y0 + kTY - 1) / kTY, planes);
  switch (g_lut_variant[0]) {  // legacy row-major-table kernel (kept for A/B timing and as a second implementation)
    case 4: lut_stage_kernel<1, 1, 4><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat); break;
    case 11: lut_stage_kernel<1, 1, 11><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat); break;
    case 12: lut_stage_kernel<1, 1, 12><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat); break;
    default: lut_stage_kernel<1, 1, 3><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat);
  }
  LERF_LAUNCHED();
  return LERF_OK;
}

/* Testing / tuning hook (see lerf_b200.h). */
void lerf_debug_lut_variant(int stage, int variant) {
  if (stage == 1 || stage == 2) g_lut_variant[stage - 1] = variant;
}

int lerf_lut_stage2(const lerf_luts_t* luts, const uint8_t* feat, int planes, int H, int W, int y0, int y1,
                    uint8_t* codes, lerf_stream_t stream) {
  int rc = stage_args_ok("lerf_lut_stage2", luts, feat, codes, planes, H, W, y0, y1);
  if (rc) return rc;
  if (planes == 0 || y0 == y1) return LERF_OK;
  const lerf_luts_impl* L = reinterpret_cast<const lerf_luts_impl*>(luts);
  StageTables t;
  for (int i = 0; i < 6; ++i) t.t[i] = L->s2[i];
  InAddr ia{1, (long long)H * W, 0, W, 1};
  if (g_lut_variant[1] == 0 || g_lut_variant[1] >= 20)
    return launch_stage_cell(L, 2, feat, ia, planes, H, W, y0, y1, codes, g_lut_variant[1] >= 20 ? g_lut_variant[1] - 20 : 0,
                             (cudaStream_t)stream);
  dim3 block(kTX, kTY), grid((W + kTX - 1) / kTX, (y1 - y0 + kTY - 1) / kTY, planes);
  if (L->oC2 == 3) {
    switch (g_lut_variant[1]) {
      case 3: lut_stage_kernel<2, 3, 3><<<grid, block, 0, (cudaStream_t)stream>>>(t, feat, ia, H, W, y0, y1, codes); break;
      default: lut_stage_kernel<2, 3, 4><<<grid, block, 0, (cudaStream_t)stream>>>(t, feat, ia, H, W, y0, y1, codes);
    }
  } else {
    lut_stage_kernel<2, 1, 4><<<grid, block, 0, (cudaStream_t)stream>>>(t, feat, ia, H, W, y0, y1, codes);
  }
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // extern "C"
