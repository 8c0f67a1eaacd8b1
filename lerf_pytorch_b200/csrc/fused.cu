// Whole SR path in one C-ABI call: stage 1 -> stage 2 -> resampling (+ uint8 epilogue) for a batch
// of images, restricted to an output row band so several GPUs can share one frame.
// Replaces the body of eltr._worker, resample/eval_lut_sr.py:541-665 (reference), without file I/O.
#include "common.cuh"

using namespace lerf;

extern "C" {

size_t lerf_sr_scratch_bytes(int planes, int oC, int H, int W) {
  if (planes < 0 || oC < 1 || H < 1 || W < 1) return 0;
  const size_t plane = (size_t)H * W;
  const size_t feat = ((size_t)planes * plane + 255) / 256 * 256;
  return feat + (size_t)planes * oC * plane;
}

int lerf_sr_fused(const lerf_luts_t* luts, int kind, const lerf_sr_plan_t* plan, const uint8_t* in, int planes,
                  int in_channels, long long in_batch_stride, long long in_chan_stride, long long in_row_stride,
                  long long in_pix_stride, float max_sigma, int oy0, int oy1, void* scratch, void* out,
                  int out_format, lerf_stream_t stream) {
  if (!luts || !plan || !in || !scratch || !out) return fail(LERF_EINVAL, "lerf_sr_fused: null pointer");
  const lerf_luts_impl* L = reinterpret_cast<const lerf_luts_impl*>(luts);
  const lerf_sr_plan_impl* P = reinterpret_cast<const lerf_sr_plan_impl*>(plan);
  const int oC = kind == LERF_KIND_GAUSS ? 3 : 1;
  if (kind != LERF_KIND_GAUSS && kind != LERF_KIND_LINEAR) return fail(LERF_EINVAL, "lerf_sr_fused: unknown kind %d", kind);
  if (L->oC2 != oC) return fail(LERF_EINVAL, "lerf_sr_fused: LUT set has oC=%d but kind %d needs %d", L->oC2, kind, oC);
  if (oy0 < 0 || oy1 > P->oH || oy0 > oy1) return fail(LERF_EINVAL, "lerf_sr_fused: bad output band [%d,%d) of %d", oy0, oy1, P->oH);
  if (planes == 0 || oy0 == oy1) return LERF_OK;
  const int H = P->H, W = P->W;
  // input rows the band depends on: taps -> +-3 rows of stage 2 -> +-3 rows of stage 1
  auto clampr = [&](int r) { return r < 0 ? 0 : (r > H - 1 ? H - 1 : r); };
  const int supp = P->general ? P->support : 2;
  const int c0 = clampr(P->h_left_y[oy0]), c1 = clampr(P->h_left_y[oy1 - 1] + supp - 1);
  const int f0 = clampr(c0 - 3), f1 = clampr(c1 + 3);
  uint8_t* feat = (uint8_t*)scratch;
  uint8_t* codes = feat + ((size_t)planes * H * W + 255) / 256 * 256;
  if (in_channels < 1 || (planes % in_channels)) return fail(LERF_EINVAL, "lerf_sr_fused: planes=%d not a multiple of in_channels=%d", planes, in_channels);
  // Batches: the role-interleaved pipeline kernel (pipeline.cu) overlaps the three steps on different plane groups.
  const InAddr ia{in_channels, in_batch_stride, in_chan_stride, in_row_stride, in_pix_stride};
  int rc = sr_pipeline(L, kind, P, in, planes, ia, max_sigma, oy0, oy1, feat, codes, out, out_format, (cudaStream_t)stream);
  if (rc != -1) return rc;
  rc = lerf_lut_stage1(luts, in, planes, H, W, in_channels, in_batch_stride, in_chan_stride, in_row_stride,
                           in_pix_stride, f0, f1 + 1, feat, stream);
  if (rc) return rc;
  rc = lerf_lut_stage2(luts, feat, planes, H, W, c0, c1 + 1, codes, stream);
  if (rc) return rc;
  return lerf_resize_sr(kind, plan, feat, codes, planes, in_channels, max_sigma, oy0, oy1, out, out_format, stream);
}

/* Testing / tuning hook (see lerf_b200.h). */
void lerf_debug_pipeline(int enabled, int min_blocks, int group_planes) {
  g_dbg.pipe_enabled = enabled != 0;
  g_dbg.pipe_minb = min_blocks;
  g_dbg.pipe_group = group_planes;
}

}  // extern "C"
