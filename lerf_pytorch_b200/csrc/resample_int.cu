// Integer-scale LeRF-G SR resampler: plain launches.  Kernel body and design notes: resample_int.cuh.
#include <math.h>

#include "resample_int.cuh"

namespace lerf {

using namespace rsi;

int g_variant = 0;  // testing hook: 0 = plain form, 5 blocks/SM; 1 = hoisted form, 3 blocks/SM; 2 = plain form, 4 blocks/SM

template <int S, int FMT, bool HOIST, int MINB>
__global__ void __launch_bounds__(kCX* kCY, MINB)
    resize_sr_int_gauss_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH,
                               int oW, const __grid_constant__ IntGeom<S> g, float max_sigma, int channels, int ly0,
                               int oy0, int oy1, void* __restrict__ out) {
  __shared__ Smem sm;
  resize_int_body<S, FMT, HOIST>(feat, codes, H, W, oH, oW, g, max_sigma, channels, ly0, oy0, oy1, out, blockIdx.x,
                                 blockIdx.y, blockIdx.z, sm);
}

template <int S>
static int launch_int(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                      float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  const IntGeom<S> g = make_geom<S>(P, max_sigma);
  // cell rows touched by the output band
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  dim3 block(kCX * kCY), grid((P->W + 1 + kCX - 1) / kCX, (ly1 - ly0 + 1 + kCY - 1) / kCY, planes);
#define LERF_GK(F, HO, B)                                                                                              \
  resize_sr_int_gauss_kernel<S, F, HO, B><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, max_sigma, \
                                                                  channels, ly0, oy0, oy1, out)
#define LERF_GO(F)                                   \
  if (g_variant == 1) LERF_GK(F, true, 3);           \
  else if (g_variant == 2) LERF_GK(F, false, 4);     \
  else LERF_GK(F, false, 5)
  switch (fmt) {
    case LERF_OUT_F32: LERF_GO(LERF_OUT_F32); break;
    case LERF_OUT_U8: LERF_GO(LERF_OUT_U8); break;
    case LERF_OUT_U8_HWC: LERF_GO(LERF_OUT_U8_HWC); break;
    default: return fail(LERF_EINVAL, "unknown out_format %d", fmt);
  }
#undef LERF_GO
#undef LERF_GK
  LERF_LAUNCHED();
  return LERF_OK;
}

// Called by lerf_resize_sr when the plan is periodic.  Returns -1 when this path does not apply.
int resize_sr_int_gauss(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                        float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  if (!(max_sigma >= 0.0f) || max_sigma > 64.0f) return -1;
  if (fmt == LERF_OUT_F32 && ((uintptr_t)out & 15)) return -1;
  switch (P->int_scale) {
    case 2: return launch_int<2>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 3: return launch_int<3>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 4: return launch_int<4>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 8: return launch_int<8>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    default: return -1;
  }
}

void resize_int_config(int variant) { g_variant = variant; }

}  // namespace lerf
