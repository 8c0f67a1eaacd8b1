// Role-interleaved ("horizontally fused") SR pipeline kernel for sm_100a.
//
// The three steps of the path -- stage 1 (reference: resample/eval_lut_sr.py:541-577), stage 2 (:579-628) and the
// steerable resampling + epilogue (resize_right/resize_right2d_numpy.py:162-223, eval_lut_sr.py:663-665) -- are
// each bound by a DIFFERENT unit of the SM (ncu, profiles/): stage 1 by the L1 data stage (128-bit cell gathers),
// stage 2 by L2 bandwidth (one 32-byte max-tap block per lookup, L1 bypassed), the resampler by issue slots (FP64/XU/FP32 arithmetic).
// Run back to back, each leaves the other two units idle.  This kernel runs all three AT ONCE on different plane
// groups of a batch: launch j does stage 1 of group j, stage 2 of group j-1 and the resampling of group j-2
// (a software pipeline over the batch; stream order provides the dependencies).  Inside a launch the blocks of
// the three roles are interleaved in proportion to their counts, so every SM hosts a mix of the three roles for
// the whole launch and all roles finish together.
//
// Every block is exactly one role and runs the same device body as the stand-alone kernels (lut_cell_body.cuh,
// lut_mt.cuh, resample_int.cuh): the bytes produced are identical to the three-launch path.
#ifdef LERF_EXPERIMENTS
#include "lut_cell_body.cuh"
#include "lut_mt.cuh"
#include "resample_int.cuh"

namespace lerf {

namespace {

struct RoleGrid {
  int gx, gy, planes, p0;  // blocks = gx * gy * planes; p0 = first (global) plane of the group
  __host__ __device__ long long blocks() const { return (long long)gx * gy * planes; }
};

struct Role1 {  // stage 1 on cell-packed tables
  cellk::CellTables tabs;
  const uint8_t* in;
  InAddr ia;
  int y0, y1;
  uint8_t* feat;
  RoleGrid g;
};

struct Role2 {  // stage 2 (oC = 3) on max-tap block tables
  mt::MtTables tabs;
  const uint8_t* feat;
  int y0, y1;
  uint8_t* codes;
  RoleGrid g;
};

template <int S>
struct Role3 {  // integer-scale Gaussian resampling
  const uint8_t* feat;
  const uint8_t* codes;
  rsi::IntGeom<S> geom;
  const rsi::CoefTabs* ct;
  int channels, ly0, oy0, oy1;
  void* out;
  RoleGrid g;
};

template <int S>
struct PipeArgs {
  int H, W, oH, oW;
  unsigned n1, n2, n3;  // blocks per role (0 = role absent from this launch)
  Role1 r1;
  Role2 r2;
  Role3<S> r3;
};

__device__ __forceinline__ void split_block(unsigned l, const RoleGrid& g, int& bx, int& by, int& p) {
  bx = (int)(l % (unsigned)g.gx);
  const unsigned t = l / (unsigned)g.gx;
  by = (int)(t % (unsigned)g.gy);
  p = g.p0 + (int)(t / (unsigned)g.gy);
}

template <int S, int FMT, int MINB>
__global__ void __launch_bounds__(256, MINB) sr_pipeline_kernel(const __grid_constant__ PipeArgs<S> a) {
  __shared__ __align__(16) unsigned char smem_raw[sizeof(rsi::Smem)];
  static_assert(sizeof(rsi::Smem) >= cellk::kTileWords * 4 && sizeof(rsi::Smem) >= (8 + 2 * mt::kHalo) * mt::kPitch * 8, "smem union");
  // Proportional interleave: among blocks [0, b) there are floor(b * n3 / N) role-3 blocks; the others are split
  // between roles 1 and 2 the same way.  Blocks are dispatched in index order, so all roles progress at the same
  // fraction and drain together.
  const unsigned b = blockIdx.x;
  const unsigned long long N = (unsigned long long)a.n1 + a.n2 + a.n3;
  const unsigned c3 = (unsigned)(((unsigned long long)b * a.n3) / N);
  const bool is3 = (unsigned)((((unsigned long long)b + 1) * a.n3) / N) > c3;
  int bx, by, p;
  if (is3) {
    split_block(c3, a.r3.g, bx, by, p);
    rsi::resize_int_body<S, FMT, 6>(a.r3.feat, a.r3.codes, a.H, a.W, a.oH, a.oW, a.r3.geom, a.r3.ct, a.r3.channels,
                                 a.r3.ly0, a.r3.oy0, a.r3.oy1, a.r3.out, bx, by, p, *reinterpret_cast<rsi::Smem*>(smem_raw));
    return;
  }
  const unsigned k = b - c3;
  const unsigned long long n12 = (unsigned long long)a.n1 + a.n2;
  const unsigned c1 = (unsigned)(((unsigned long long)k * a.n1) / n12);
  const bool is1 = (unsigned)((((unsigned long long)k + 1) * a.n1) / n12) > c1;
  if (is1) {
    split_block(c1, a.r1.g, bx, by, p);
    cellk::lut_stage_cell_body<1, 1>(a.r1.tabs, a.r1.in, a.r1.ia, a.H, a.W, a.r1.y0, a.r1.y1, a.r1.feat, bx, by, p,
                                     reinterpret_cast<uint32_t*>(smem_raw));
  } else {
    split_block(k - c1, a.r2.g, bx, by, p);
    mt::lut_stage2_mt_body<1, 1>(a.r2.tabs, a.r2.feat, a.H, a.W, a.r2.y0, a.r2.y1, a.r2.codes, bx, by, p,
                                 reinterpret_cast<uint2*>(smem_raw));
  }
}


template <int S>
int run_pipeline(const lerf_luts_impl* L, const lerf_sr_plan_impl* P, const uint8_t* in, int planes, const InAddr& ia,
                 float max_sigma, int oy0, int oy1, uint8_t* feat, uint8_t* codes, void* out, int fmt, cudaStream_t st) {
  const int H = P->H, W = P->W;
  auto clampr = [&](int r) { return r < 0 ? 0 : (r > H - 1 ? H - 1 : r); };
  // input rows the band depends on: taps -> +-3 rows of stage 2 -> +-3 rows of stage 1
  const int c0 = clampr(P->h_left_y[oy0]), c1 = clampr(P->h_left_y[oy1 - 1] + 1);
  const int f0 = clampr(c0 - 3), f1 = clampr(c1 + 3);
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  int gsz = g_dbg.pipe_group > 0 ? g_dbg.pipe_group : (planes >= 12 ? ia.channels : 1);
  if (gsz > planes) gsz = planes;
  const int G = (planes + gsz - 1) / gsz;

  if (!rsi::ref_tap_ok<S>(P, max_sigma)) return -1;  // the resampler role takes production's nearest-tap weights (resample_int.cuh MODE 6)
  PipeArgs<S> a;
  memset(&a, 0, sizeof(a));
  a.H = H; a.W = W; a.oH = P->oH; a.oW = P->oW;
  for (int i = 0; i < 6; ++i) a.r1.tabs.t[i] = i < 3 ? L->c1[i] : nullptr;
  a.r1.tabs.h = cell::Hash{(uint32_t)L->cell_hash[0], (uint32_t)L->cell_hash[1], (uint32_t)L->cell_hash[2]};
  a.r1.in = in; a.r1.ia = ia; a.r1.y0 = f0; a.r1.y1 = f1 + 1; a.r1.feat = feat;
  for (int i = 0; i < 6; ++i) a.r2.tabs.t[i] = L->mt2[i];
  a.r2.feat = feat; a.r2.y0 = c0; a.r2.y1 = c1 + 1; a.r2.codes = codes;
  a.r3.feat = feat; a.r3.codes = codes; a.r3.geom = rsi::make_geom<S>(P, max_sigma, true, /*signed_diff=*/true); a.r3.ct = rsi::plan_coef_tabs(P, max_sigma, st);
  if (!a.r3.ct) return fail(LERF_ECUDA, "uploading the hyper decode tables failed");
  a.r3.channels = ia.channels; a.r3.ly0 = ly0; a.r3.oy0 = oy0; a.r3.oy1 = oy1; a.r3.out = out;
  const int gx12 = (W + cellk::kTX - 1) / cellk::kTX;
  static_assert(cellk::kTX == mt::kTX && cellk::kTY == 8, "stage tiles");
  for (int j = 0; j < G + 2; ++j) {
    auto group = [&](int g, RoleGrid& rg, int gx, int gy) {
      if (g < 0 || g >= G) { rg = RoleGrid{1, 1, 0, 0}; return 0u; }
      const int p0 = g * gsz, n = (p0 + gsz <= planes ? gsz : planes - p0);
      rg = RoleGrid{gx, gy, n, p0};
      return (unsigned)rg.blocks();
    };
    a.n1 = group(j, a.r1.g, gx12, (f1 + 1 - f0 + cellk::kTY - 1) / cellk::kTY);
    a.n2 = group(j - 1, a.r2.g, gx12, (c1 + 1 - c0 + 7) / 8);
    a.n3 = group(j - 2, a.r3.g, (W + 1 + rsi::kCX - 1) / rsi::kCX, (ly1 - ly0 + 1 + rsi::kCY - 1) / rsi::kCY);
    const unsigned long long N = (unsigned long long)a.n1 + a.n2 + a.n3;
    if (N == 0) continue;
    if (N > 0x7fffffffull) return fail(LERF_EINVAL, "lerf_sr_fused: too many blocks in one pipeline launch");
#define LERF_GO(F, B) sr_pipeline_kernel<S, F, B><<<(unsigned)N, 256, 0, st>>>(a)
#define LERF_GOB(F)                 \
  if (g_dbg.pipe_minb == 4) LERF_GO(F, 4); \
  else if (g_dbg.pipe_minb == 2) LERF_GO(F, 2); \
  else LERF_GO(F, 3)
    switch (fmt) {
      case LERF_OUT_F32: LERF_GOB(LERF_OUT_F32); break;
      case LERF_OUT_U8: LERF_GOB(LERF_OUT_U8); break;
      case LERF_OUT_U8_HWC: LERF_GOB(LERF_OUT_U8_HWC); break;
      default: return fail(LERF_EINVAL, "unknown out_format %d", fmt);
    }
#undef LERF_GOB
#undef LERF_GO
    LERF_LAUNCHED();
  }
  return LERF_OK;
}

}  // namespace

// Called by lerf_sr_fused.  Returns -1 when the pipeline does not apply (the caller then issues the three plain launches).
int sr_pipeline(const lerf_luts_impl* L, int kind, const lerf_sr_plan_impl* P, const uint8_t* in, int planes, const InAddr& ia,
                float max_sigma, int oy0, int oy1, uint8_t* feat, uint8_t* codes, void* out, int fmt, cudaStream_t st) {
  if (!g_dbg.pipe_enabled || kind != LERF_KIND_GAUSS || L->oC2 != 3 || !L->mt2[0] || !P->int_scale || planes < 2) return -1;
  if (!(max_sigma >= 0.0f) || max_sigma > 64.0f) return -1;
  if (fmt == LERF_OUT_F32 && ((uintptr_t)out & 15)) return -1;
  switch (P->int_scale) {
    case 2: return run_pipeline<2>(L, P, in, planes, ia, max_sigma, oy0, oy1, feat, codes, out, fmt, st);
    case 3: return run_pipeline<3>(L, P, in, planes, ia, max_sigma, oy0, oy1, feat, codes, out, fmt, st);
    case 4: return run_pipeline<4>(L, P, in, planes, ia, max_sigma, oy0, oy1, feat, codes, out, fmt, st);
    case 8: return run_pipeline<8>(L, P, in, planes, ia, max_sigma, oy0, oy1, feat, codes, out, fmt, st);
    default: return -1;
  }
}


}  // namespace lerf

#else  // product build: the pipeline kernel is an experiment; lerf_sr_fused issues its three launches
#include "common.cuh"
namespace lerf {
int sr_pipeline(const lerf_luts_impl*, int, const lerf_sr_plan_impl*, const uint8_t*, int, const InAddr&, float, int, int, uint8_t*,
                uint8_t*, void*, int, cudaStream_t) {
  return -1;
}
}  // namespace lerf
#endif
