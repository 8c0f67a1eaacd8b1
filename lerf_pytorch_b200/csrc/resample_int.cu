// Integer-scale LeRF-G SR resampler: plain launches.  Kernel body and design notes: resample_int.cuh.
#include <math.h>

#include "resample_int.cuh"

namespace lerf {

using namespace rsi;

// Device copy of the per-code tables for `max_sigma`, cached in the plan (a plan, like the reference's resizer objects,
// is not thread-safe).  The upload is stream-ordered before the kernels that read it.
const CoefTabs* rsi::plan_coef_tabs(const lerf_sr_plan_impl* P, float max_sigma, cudaStream_t st) {
  lerf_sr_plan_impl* M = const_cast<lerf_sr_plan_impl*>(P);
  if (!M->coef_dev) {
    if (cudaMalloc(&M->coef_dev, sizeof(CoefTabs)) != cudaSuccess) return nullptr;
    M->coef_host = malloc(sizeof(CoefTabs));
    M->coef_sigma = -1.0f;
  }
  if (M->coef_sigma != max_sigma) {
    make_coef_tabs(max_sigma, *(CoefTabs*)M->coef_host);
    if (cudaMemcpyAsync(M->coef_dev, M->coef_host, sizeof(CoefTabs), cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
    M->coef_sigma = max_sigma;
  }
  return (const CoefTabs*)M->coef_dev;
}

// testing hook: 0 = production = ROWQ form (unsigned fixed point, row term hoisted; resample_int.cuh), 4 blocks/SM;
// 4 = ROWQ, 5 blocks/SM; 5 = plain form (4 FP64 per exponent), 5 blocks/SM (production until r1d); 2 = plain, 4 blocks/SM;
// 1 = fully hoisted signed form, 3 blocks/SM
int g_variant = 0;

template <int S, int FMT, int MODE, int MINB>
__global__ void __launch_bounds__(kCX* kCY, MINB)
    resize_sr_int_gauss_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH,
                               int oW, const __grid_constant__ IntGeom<S> g, const CoefTabs* __restrict__ ct,
                               int channels, int ly0, int oy0, int oy1, void* __restrict__ out) {
  __shared__ Smem sm;
  resize_int_body<S, FMT, MODE>(feat, codes, H, W, oH, oW, g, ct, channels, ly0, oy0, oy1, out, blockIdx.x, blockIdx.y,
                                 blockIdx.z, sm);
}

template <int S>
static int launch_int(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                      float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  const IntGeom<S> g = make_geom<S>(P, max_sigma, /*unsigned_form=*/g_variant == 0 || g_variant == 4);
  const CoefTabs* ct = plan_coef_tabs(P, max_sigma, st);
  if (!ct) return fail(LERF_ECUDA, "uploading the hyper decode tables failed");
  // cell rows touched by the output band
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  dim3 block(kCX * kCY), grid((P->W + 1 + kCX - 1) / kCX, (ly1 - ly0 + 1 + kCY - 1) / kCY, planes);
#define LERF_GK(F, HO, B)                                                                                              \
  resize_sr_int_gauss_kernel<S, F, HO, B><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, ct, \
                                                                  channels, ly0, oy0, oy1, out)
#define LERF_GO(F)                                   \
  if (g_variant == 1) LERF_GK(F, 1, 3);              \
  else if (g_variant == 2) LERF_GK(F, 0, 4);         \
  else if (g_variant == 5) LERF_GK(F, 0, 5);         \
  else if (g_variant == 4) LERF_GK(F, 2, 5);         \
  else LERF_GK(F, 2, 4)
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
