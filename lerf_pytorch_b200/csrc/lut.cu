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
#include "lut_rm.cuh"

namespace lerf {

thread_local std::string g_last_error;
thread_local long long g_launches = 0;
thread_local DebugState g_dbg;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#ifdef LERF_EXPERIMENTS  // first implementation of the stages on the shipped row-major tables, kept for A/B runs
template <int STAGE, int OC, int MINB>
__global__ void __launch_bounds__(rm::kTX* rm::kTY, (MINB >= 11 ? 3 : MINB))
    lut_stage_kernel(rm::StageTables tabs, const uint8_t* __restrict__ in, InAddr ia, int H, int W, int y0, int y1,
                     uint8_t* __restrict__ out) {
  __shared__ uint8_t tile[rm::kTileBytes];
  rm::lut_stage_body<STAGE, OC, (MINB >= 11 ? MINB - 10 : 0)>(tabs, in, ia, H, W, y0, y1, out, blockIdx.x, blockIdx.y,
                                                             blockIdx.z, tile);
}
#endif

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
  const rm::Simplex s = rm::simplex_of(va, vb, vc, vd);
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
  for (int i = 0; i < 3; ++i) L->cell_hash[i] = g_dbg.cell_hash[i];
  int rc = build_cell_tables(L, host_tables);
  if (!rc) rc = build_pw_tables(L);
  if (rc) {
    cudaFree(L->cp_block);
    cudaFree(L->pw_block);
    cudaFree(L->mt_block);
    cudaFree(L->cell_block);
    cudaFree(L->block);
    delete L;
    return rc;
  }
  // Reserve persisting L2 for the tables (best effort; the window itself is per stream).
  size_t want = L->cell_block_bytes, have = 0;
  if (cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize) != cudaSuccess) have = 0;
  if (want > have) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);  // only ever raise the device-wide limit
  cudaGetLastError();
  *out = reinterpret_cast<lerf_luts_t*>(L);
  return LERF_OK;
}

void lerf_luts_destroy(lerf_luts_t* luts) {
  if (!luts) return;
  lerf_luts_impl* L = reinterpret_cast<lerf_luts_impl*>(luts);
  cudaSetDevice(L->device);
  cudaCtxResetPersistingL2Cache();  // lines this set's window marked persisting go back to normal
  cudaGetLastError();
  cudaFree(L->block);
  cudaFree(L->cell_block);
  cudaFree(L->mt_block);
  cudaFree(L->pw_block);
  cudaFree(L->cp_block);
  delete L;
}

int lerf_luts_oc(const lerf_luts_t* luts) {
  return luts ? reinterpret_cast<const lerf_luts_impl*>(luts)->oC2 : 0;
}

// Best effort by contract: a device that offers no access-policy window (MIG slices, vGPU) or a smaller one is not an
// error, the kernels do not need it.  What it buys on B200 is little in any case: the production tables are far smaller
// than the 126 MB L2 and stay resident without help (ncu r2: L2 sector hit rate 96.9 % in stage 1, 93.9 % in a cold
// single-frame stage-2 launch).
int lerf_luts_pin_l2(const lerf_luts_t* luts, lerf_stream_t stream) {
  if (!luts) return fail(LERF_EINVAL, "lerf_luts_pin_l2: null handle");
  const lerf_luts_impl* L = reinterpret_cast<const lerf_luts_impl*>(luts);
  int max_window = 0;
  if (cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, L->device) != cudaSuccess || max_window <= 0) {
    cudaGetLastError();
    return LERF_OK;
  }
  const bool pw = g_dbg.l2_window == 1 && L->pw_block;
  void* base = pw ? L->pw_block : L->cell_block;
  size_t bytes = pw ? L->pw_block_bytes : L->cell_block_bytes;
  if (bytes > (size_t)max_window) bytes = (size_t)max_window;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  attr.accessPolicyWindow.base_ptr = base;
  attr.accessPolicyWindow.num_bytes = bytes;
  attr.accessPolicyWindow.hitRatio = pw ? 0.25f : 1.0f;  // a quarter of the order planes of a table are ever touched
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
  return LERF_OK;
}

void lerf_debug_l2_window(int which) { g_dbg.l2_window = which; }
void lerf_debug_carveout(int percent) { g_dbg.carveout = percent; }

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
  InAddr ia{in_channels, in_batch_stride, in_chan_stride, in_row_stride, in_pix_stride};
  const int v1 = g_dbg.lut_variant[0];
#ifdef LERF_EXPERIMENTS
  if (v1 >= 80)  // paired-window / cell-pair tables (lut_pw.cu)
    return launch_stage_pw(L, 1, in, ia, planes, H, W, y0, y1, feat, v1 - 80, (cudaStream_t)stream);
  if (v1 >= 1 && v1 < 20) {  // row-major-table kernel
    rm::StageTables t;
    for (int i = 0; i < 3; ++i) t.t[i] = L->s1[i];
    for (int i = 3; i < 6; ++i) t.t[i] = nullptr;
    dim3 block(rm::kTX * rm::kTY), grid((W + rm::kTX - 1) / rm::kTX, (y1 - y0 + rm::kTY - 1) / rm::kTY, planes);
    switch (v1) {
      case 4: lut_stage_kernel<1, 1, 4><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat); break;
      case 11: lut_stage_kernel<1, 1, 11><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat); break;
      case 12: lut_stage_kernel<1, 1, 12><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat); break;
      default: lut_stage_kernel<1, 1, 3><<<grid, block, 0, (cudaStream_t)stream>>>(t, in, ia, H, W, y0, y1, feat);
    }
    LERF_LAUNCHED();
    return LERF_OK;
  }
#endif
  // production: cell-packed tables (lut_cell.cu)
  return launch_stage_cell(L, 1, in, ia, planes, H, W, y0, y1, feat, v1 >= 20 && v1 < 40 ? v1 - 20 : 0, (cudaStream_t)stream);
}

/* Testing / tuning hook: block-swizzle weights used by the NEXT lerf_luts_create (see lerf_b200.h). */
void lerf_debug_cell_hash(int ha, int hb, int hc) {
  g_dbg.cell_hash[0] = ha; g_dbg.cell_hash[1] = hb; g_dbg.cell_hash[2] = hc;
}

/* Testing / tuning hook (see lerf_b200.h). */
void lerf_debug_lut_variant(int stage, int variant) {
  if (stage == 1 || stage == 2) g_dbg.lut_variant[stage - 1] = variant;
}

int lerf_lut_stage2(const lerf_luts_t* luts, const uint8_t* feat, int planes, int H, int W, int y0, int y1,
                    uint8_t* codes, lerf_stream_t stream) {
  int rc = stage_args_ok("lerf_lut_stage2", luts, feat, codes, planes, H, W, y0, y1);
  if (rc) return rc;
  if (planes == 0 || y0 == y1) return LERF_OK;
  const lerf_luts_impl* L = reinterpret_cast<const lerf_luts_impl*>(luts);
  InAddr ia{1, (long long)H * W, 0, W, 1};
  const int v2 = g_dbg.lut_variant[1];
  // production: paired-window tables (lut_pw.cu) for LeRF-G, the cell kernel for LeRF-L (oC = 1 cells are 16 bytes and hit L1).
  // 20..39 selects the cell kernel and 80.. the window kernel for either model: each is the other's second implementation.
  if ((v2 == 0 && L->oC2 == 3) || v2 >= 80)
    return launch_stage_pw(L, 2, feat, ia, planes, H, W, y0, y1, codes, v2 >= 80 ? v2 - 80 : 0, (cudaStream_t)stream);
#ifdef LERF_EXPERIMENTS
  if (v2 >= 60 && L->oC2 == 3)  // max-tap block tables (lut_mt.cuh), production of r1
    return launch_stage2_mt(L, feat, planes, H, W, y0, y1, codes, v2 - 60, (cudaStream_t)stream);
  if (v2 >= 40 && L->oC2 == 3)  // table-format mix (lut_mix.cuh)
    return launch_stage2_mix(L, feat, planes, H, W, y0, y1, codes, v2 - 40, (cudaStream_t)stream);
  if (v2 >= 1 && v2 < 20) {  // row-major-table kernel
    rm::StageTables t;
    for (int i = 0; i < 6; ++i) t.t[i] = L->s2[i];
    dim3 block(rm::kTX * rm::kTY), grid((W + rm::kTX - 1) / rm::kTX, (y1 - y0 + rm::kTY - 1) / rm::kTY, planes);
    if (L->oC2 == 3) {
      if (v2 == 3) lut_stage_kernel<2, 3, 3><<<grid, block, 0, (cudaStream_t)stream>>>(t, feat, ia, H, W, y0, y1, codes);
      else lut_stage_kernel<2, 3, 4><<<grid, block, 0, (cudaStream_t)stream>>>(t, feat, ia, H, W, y0, y1, codes);
    } else {
      lut_stage_kernel<2, 1, 4><<<grid, block, 0, (cudaStream_t)stream>>>(t, feat, ia, H, W, y0, y1, codes);
    }
    LERF_LAUNCHED();
    return LERF_OK;
  }
#endif
  return launch_stage_cell(L, 2, feat, ia, planes, H, W, y0, y1, codes, v2 >= 20 && v2 < 40 ? v2 - 20 : 0, (cudaStream_t)stream);
}

int lerf_build_has_experiments(void) {
#ifdef LERF_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"
