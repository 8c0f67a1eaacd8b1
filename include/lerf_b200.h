/*
 * lerf_b200.h -- C ABI of the B200-native LeRF LUT-inference hot path.
 *
 * Shared library: lerf_pytorch_b200/liblerf_b200.so (built by __graft_entry__.build()).
 * Plain C: raw device pointers, sizes and a CUDA stream handle; no torch / C++ types.
 * Every entry point returns 0 on success, a LERF_E* code otherwise; the message of the
 * last failure on the calling thread is returned by lerf_last_error_string().  No call
 * synchronises the stream unless it says so.  All image memory is caller-owned DEVICE
 * memory; only the LUT handle and the SR geometry plan are callee-owned.
 *
 * The reference (ddlee-cn/LeRF-PyTorch) has no FFI or plugin registry: its boundary for
 * this path is three Python call signatures.  Each function below names the reference
 * interface it replaces (file:line in the reference tree); INTEGRATION.md shows the
 * ctypes binding a maintainer adds behind those Python signatures.
 * Test and tuning switches are NOT part of this interface: they live in lerf_b200_testing.h.
 *
 * Layouts
 *   image planes : uint8 or float32, planar [P][H][W]; P = batch * channels, every plane is
 *                  processed independently (the reference loops colour channels the same way).
 *   codes        : uint8 planar [P*oC][H][W], plane p*oC+k; k = rho, sigma_x, sigma_y for
 *                  LeRF-G (oC = 3), alpha for LeRF-L (oC = 1).  hyper = codes/255
 *                  (eval_lut_sr.py:623-628, channel order :651-661).
 *   outputs      : planar [P][oH][oW] float32 or uint8; LERF_OUT_U8_HWC interleaves groups of
 *                  `channels` planes into [B][oH][oW][channels] (eval_lut_sr.py:663-665).
 */
#ifndef LERF_B200_H_
#define LERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LERF_OK 0
#define LERF_EINVAL 1   /* bad argument (mirrors the reference's ValueError) */
#define LERF_ECUDA 2    /* a CUDA runtime call failed */
#define LERF_ENOMEM 3
#define LERF_EUNSUPPORTED 4

#define LERF_ABI_VERSION 1

/* resampling function family */
#define LERF_KIND_GAUSS 0   /* LeRF-G: steerable anisotropic Gaussian (resize_right2d_numpy.py:142-223) */
#define LERF_KIND_LINEAR 1  /* LeRF-L: amplified linear            (resize_right2d_numpy.py:225-282) */

/* output formats */
#define LERF_OUT_F32 0     /* float32 planar, what resize()/warp() return in the reference (float64 there) */
#define LERF_OUT_U8 1      /* uint8 planar, clip(round_half_even(x),0,255) (eval_lut_sr.py:663) */
#define LERF_OUT_U8_HWC 2  /* uint8 interleaved, the array the eval scripts save */

/* fixed interpolation kernels of lerf_warp_fixed (resize_right/interp_methods.py:32-100; support in brackets) */
#define LERF_WARP_NEAREST 0   /* box2d     (1)  NearestWarp2dNumpy   resize_right2d_numpy.py:460-467 */
#define LERF_WARP_BILINEAR 1  /* linear2d  (2)  BilinearWarp2dNumpy  :469-476 */
#define LERF_WARP_BICUBIC 2   /* cubic2d   (4)  BicubicWarp2dNumpy   :451-458 */
#define LERF_WARP_LANCZOS2 3  /* lanczos2d (4)  Lanczos2Warp2dNumpy  :478-485 */
#define LERF_WARP_LANCZOS3 4  /* lanczos3d (6)  Lanczos3Warp2dNumpy  :487-494 */

typedef struct lerf_luts lerf_luts_t;       /* device-resident LUT set */
typedef struct lerf_sr_plan lerf_sr_plan_t; /* device-resident SR geometry (set_shape) */
typedef void* lerf_stream_t;                /* cudaStream_t */

int lerf_abi_version(void);
const char* lerf_last_error_string(void);

/* ---- LUT set -------------------------------------------------------------------------------
 * Replaces the LUT loader eval_lut_sr.py:750-775 (= eval_lut_warp.py:308-333).
 * host_tables: 9 int8 tables of 17^4 rows, modes "sct":
 *   [0..2] = s1_s r0, s1_c r0, s1_t r0   (oC = 1)
 *   [3..8] = s2_s r0, s2_s r1, s2_c r0, s2_c r1, s2_t r0, s2_t r1   (oC = oC2 = 3 or 1)
 * Uploads and repacks them once on `device`, and (best effort) marks the block L2-persisting. */
int lerf_luts_create(const int8_t* const host_tables[9], int oC2, int device, lerf_luts_t** out);
void lerf_luts_destroy(lerf_luts_t* luts);
int lerf_luts_oc(const lerf_luts_t* luts);
/* Apply the L2 access-policy window for the LUT block to `stream` (cudaStreamSetAttribute). */
int lerf_luts_pin_l2(const lerf_luts_t* luts, lerf_stream_t stream);

/* ---- one LUT pass ---------------------------------------------------------------------------
 * Replaces FourSimplexInterpFaster(weight, img_in, h, w, interval=4, rot, mode, oC),
 * eval_lut_sr.py:24-470, up to the final np.rot90 (a view operation the host side does).
 *   table  : DEVICE int8 [17^4][oC] (the `weight` argument, row-major as np.load gives it)
 *   img    : DEVICE uint8 [C][h+pad][w+pad], already rotated and edge-padded by the caller
 *   mode   : 's','d','y','c','t' (eval_lut_sr.py:30-81); anything else -> LERF_EINVAL (:82-84)
 *   out    : DEVICE int32 [C*oC][h][w] = N, the exact integer numerator; the reference's return
 *            value is N/16 as float64. */
int lerf_lut_pass(const int8_t* table, const uint8_t* img, int C, int h, int w, char mode, int oC,
                  int32_t* out, lerf_stream_t stream);

/* ---- rotation-ensembled stages ---------------------------------------------------------------
 * lerf_lut_stage1 replaces the loop eval_lut_sr.py:541-577 (= eval_lut_warp.py:104-140):
 *   feat = clip(round_half_even(sum_N / 48), 0, 255), 12 passes (modes s,c,t x 4 rotations).
 * lerf_lut_stage2 replaces eval_lut_sr.py:579-628 (= eval_lut_warp.py:142-191):
 *   code = clip(round_half_even(sum_N / 192 + 127), 0, 255), r in {0,2} -> table r0, {1,3} -> r1.
 * in  : uint8, element (p, y, x) at in[(p / in_channels) * in_batch_stride + (p % in_channels) *
 *       in_chan_stride + y * in_row_stride + x * in_pix_stride]   (planar: H*W*C, H*W, W, 1;
 *       HWC as PIL decodes it: H*W*C, 1, W*C, C)
 * Rows [y0, y1) of the H x W image are produced (row-band sharding: neighbours outside the band
 * are still read from the full image; only true image edges clamp -- SURVEY.md A.3).
 * feat / codes are full-size planar buffers [P][H][W] / [P*oC][H][W]; only the band is written. */
int lerf_lut_stage1(const lerf_luts_t* luts, const uint8_t* in, int planes, int H, int W,
                    int in_channels, long long in_batch_stride, long long in_chan_stride,
                    long long in_row_stride, long long in_pix_stride, int y0, int y1,
                    uint8_t* feat, lerf_stream_t stream);
int lerf_lut_stage2(const lerf_luts_t* luts, const uint8_t* feat, int planes, int H, int W, int y0,
                    int y1, uint8_t* codes, lerf_stream_t stream);


/* ---- SR geometry plan --------------------------------------------------------------------------
 * Replaces Resize2dNumpy.set_shape / get_distance (resize_right2d_numpy.py:18-140) for support 2.
 * The per-axis tables are computed by the caller in float64 in the reference's operation order
 * (the geometry is separable): for output index o,
 *   left[o]  = UNPADDED index of the first tap (ceil(p - supp/2 - eps32), :85-90)
 *   dist[2*o + k] = (p + pad0) - (left + pad0 + k), k = 0,1   (:131-134)
 * HOST pointers; copied to the device by the call. */
int lerf_sr_plan_create(int H, int W, int oH, int oW, const int32_t* left_y, const double* dist_y,
                        const int32_t* left_x, const double* dist_x, int device,
                        lerf_sr_plan_t** out);
/* The same with the reference's NON-DEFAULT operator parameters: `support` taps per axis (the resizers' support_sz, any
 * value >= 1; dist tables hold `support` entries per output: dist[support*o + k]; left may be any first-tap index), the
 * np.pad mode of the IMAGE (pad_mode=..., :208; the hypers always replicate), and aa_scale = min_scale_factor when a height
 * factor below 1 has turned antialiasing on (:51-55: the caller grows `support` to ceil(support / s); the Gaussian kind
 * then scales its distances by aa_scale, :186-193), 1.0 otherwise.  Such a plan is served by the float64
 * operation-order kernel only. */
#define LERF_PAD_CONSTANT 0
#define LERF_PAD_EDGE 1
#define LERF_PAD_REFLECT 2
#define LERF_PAD_SYMMETRIC 3
#define LERF_PAD_WRAP 4
int lerf_sr_plan_create_ex(int H, int W, int oH, int oW, int support, const int32_t* left_y, const double* dist_y,
                           const int32_t* left_x, const double* dist_x, int pad_mode, double aa_scale, int device,
                           lerf_sr_plan_t** out);
void lerf_sr_plan_destroy(lerf_sr_plan_t* plan);

/* ---- SR resampling ------------------------------------------------------------------------------
 * Replaces SteeringGaussianResize2dNumpy.resize (:162-223) / AmplifiedLinearResize2dNumpy.resize
 * (:243-282) plus, for the uint8 formats, the epilogue eval_lut_sr.py:663-665.
 *   feat  : uint8 planar [P][H][W] (integer-valued image, what stage 1 produces)
 *   codes : uint8 planar [P*oC][H][W]; oC = 3 for LERF_KIND_GAUSS, 1 for LERF_KIND_LINEAR
 *   out   : rows [oy0, oy1) of each output plane are written, in `out_format`.
 *   channels : planes per image (only used by LERF_OUT_U8_HWC).
 * The image is zero outside its borders ('constant' pad, :208), the hypers replicate ('edge', :172). */
int lerf_resize_sr(int kind, const lerf_sr_plan_t* plan, const uint8_t* feat, const uint8_t* codes,
                   int planes, int channels, float max_sigma, int oy0, int oy1, void* out,
                   int out_format, lerf_stream_t stream);


/* Same operator on float32 image / float32 hyper planes in [0,1], for callers that bring their own
 * (non-LUT) hyper-parameters exactly like the reference's resize(input, rho, sigma_x, sigma_y).
 * hyper planes: h0,h1,h2 each [P][H][W] (h1,h2 ignored for LERF_KIND_LINEAR). */
int lerf_resize_sr_f32(int kind, const lerf_sr_plan_t* plan, const float* img, const float* h0,
                       const float* h1, const float* h2, int planes, float max_sigma, float* out,
                       lerf_stream_t stream);

/* ---- homographic warp ----------------------------------------------------------------------------
 * Replaces Warp2dNumpy.set_shape/get_distance (:292-407) and SteeringGaussianWarp2dNumpy.warp
 * (:516-577) / AmplifiedLinearWarp2dNumpy.warp (:597-635), plus NearestWarp2dNumpy (:460-467) and
 * the validity-mask logic of eval_lut_warp.py:197-204,229.
 *   minv   : HOST double[9], inverse of the input->output homography (np.linalg.inv(matrix), :327)
 *   pad0_y/pad0_x : leading pads for support 2, from output pixel (0,0) only (:363-369)
 *   mask   : optional DEVICE uint8 [oH][oW] (NULL to skip): 1 where the nearest-neighbour warp
 *            (support 1, pads mask_pad0_y/x) of a white image with a `mask_border`-pixel black frame
 *            equals 255.
 * Outside the input the float output is NaN where the reference produces 0/0. */
int lerf_warp(int kind, const uint8_t* feat, const uint8_t* codes, int planes, int channels, int H,
              int W, int oH, int oW, const double minv[9], int pad0_y, int pad0_x, float max_sigma,
              void* out, int out_format, uint8_t* mask, int mask_pad0_y, int mask_pad0_x,
              int mask_border, lerf_stream_t stream);

/* float32 flavour of lerf_warp (caller-supplied hyper planes in [0,1], no mask). */
int lerf_warp_f32(int kind, const float* img, const float* h0, const float* h1, const float* h2, int planes,
                  int H, int W, int oH, int oW, const double minv[9], int pad0_y, int pad0_x,
                  float max_sigma, float* out, lerf_stream_t stream);

/* lerf_warp / lerf_warp_f32 with the reference's NON-DEFAULT operator parameters: `support` taps per axis (support_sz of
 * the warp classes, :497 / :580) and the np.pad mode LERF_PAD_* of the image (:559; taps are clipped to in-1 in padded
 * coordinates, :397-398, so only the LEADING pad is ever read; the hypers replicate).  pad0_y / pad0_x are the leading pads
 * for THIS support.  Give feat + codes (uint8; any out_format) or img + h0[, h1, h2] (float32; LERF_OUT_F32), the other
 * group NULL.  No mask (its support is 1 whatever this one is: use lerf_warp with out = NULL).  Float64 operation-order
 * kernel only. */
int lerf_warp_ex(int kind, const uint8_t* feat, const uint8_t* codes, const float* img, const float* h0, const float* h1,
                 const float* h2, int planes, int channels, int H, int W, int oH, int oW, const double minv[9], int support,
                 int pad_mode, int pad0_y, int pad0_x, float max_sigma, void* out, int out_format, lerf_stream_t stream);

/* Fixed-kernel warps, the baselines the reference compares LeRF with: Warp2dNumpy.warp
 * (resize_right2d_numpy.py:409-449) with a separable kernel LERF_WARP_*.  img: DEVICE planar [P][H][W], float32 or
 * (img_is_u8 != 0) uint8; out: DEVICE float32 planar [P][oH][oW]; pad0_y/pad0_x: leading pads for THIS kernel's support
 * (lerf_warp_fixed_support(kernel)), from output pixel (0,0) only (:363-369).  NaN where the reference produces 0/0. */
int lerf_warp_fixed_support(int kernel);
int lerf_warp_fixed(int kernel, const void* img, int img_is_u8, int planes, int H, int W, int oH, int oW,
                    const double minv[9], int pad0_y, int pad0_x, float* out, lerf_stream_t stream);

/* ---- fused SR path ---------------------------------------------------------------------------------
 * Stage 1 + stage 2 + resampling + epilogue for a batch of images in one call: the body of
 * eltr._worker (eval_lut_sr.py:541-665) without file I/O.  `in` is addressed like lerf_lut_stage1.
 * `scratch` is caller-owned device memory of at least lerf_sr_scratch_bytes(planes, oC, H, W) bytes.
 * Output rows [oy0, oy1) are produced (row-band sharding across GPUs). */
size_t lerf_sr_scratch_bytes(int planes, int oC, int H, int W);
int lerf_sr_fused(const lerf_luts_t* luts, int kind, const lerf_sr_plan_t* plan, const uint8_t* in,
                  int planes, int in_channels, long long in_batch_stride, long long in_chan_stride,
                  long long in_row_stride, long long in_pix_stride, float max_sigma, int oy0, int oy1,
                  void* scratch, void* out, int out_format, lerf_stream_t stream);


/* ---- result images as PNG files, on the device (output side) ---------------------------------
 * Replaces the host-side `Image.fromarray(...).save(...)` of the result images (resample/eval_lut_sr.py:667-708,
 * eval_lut_warp.py:224-262).  img: DEVICE uint8 [H][W][channels] (channels 1, 2, 3, 4 = grey, grey + alpha, RGB, RGBA),
 * contiguous; png: DEVICE, 16-byte aligned, at least lerf_png_stored_bytes(...) rounded up to 4 bytes; scratch32: DEVICE, 32
 * bytes, 8-byte aligned.  The file is a complete PNG with filter type 0 and STORED deflate blocks (lossless, no
 * compression: H * (1 + W * channels) bytes of image data + 5 per 65535 + 63); Adler-32 and CRC-32 are computed on the
 * device.  Copy lerf_png_stored_bytes(...) bytes to the host and write them to disk.  lerf_png_stored_bytes returns -1 for a
 * bad shape or an image whose data does not fit one IDAT chunk (2^31 - 1 bytes). */
long long lerf_png_stored_bytes(int H, int W, int channels);
int lerf_png_encode_stored(const uint8_t* img, int H, int W, int channels, uint8_t* png, long long png_capacity, void* scratch32,
                           lerf_stream_t stream);


/* ---- LUT fine-tuning operators ---------------------------------------------------------------
 * Forward and backward of the two differentiable pieces the reference trains through when it fine-tunes its LUTs
 * (resample/train_model.py --lutft); float32 like its torch path.
 *
 * lerf_lut_ft_forward / _backward replace SWF2LUT.InterpTorchBatch (resample/model.py:172-385):
 *   weight : DEVICE float32 [17^4][oC], the table in int8 units (the caller applies *127, the BPDA round and the clamp of
 *            :177-179); img : DEVICE float32 [P][h+pad][w+pad], integer-valued 0..255, edge-padded by the mode's pad
 *            (1 for s, 2 for d / y, 3 for c / t); out / grad_out : [P][oC][h][w] (already divided by q = 16).
 *   grad_weight : [17^4][oC], ACCUMULATED into (zero it first); the pixels carry no gradient (floor_divide / %).
 *   lsb_like_reference != 0 reproduces model.py:229-232,240-243, which reads the LSBs of modes c and t from the pixels
 *   of mode y; 0 uses the mode's own taps like the inference path (eval_lut_sr.py:30-81).
 * lerf_resize_sr_f32_backward is the gradient of lerf_resize_sr_f32 for LERF_KIND_GAUSS, i.e. of
 * SteeringGaussianResize2dTorch.resize (resize_right/resize_right2d_torch.py:154-197): grad_img and grad_h0..2 are
 * [P][H][W], ACCUMULATED into, and may be NULL where a gradient is not needed. */
int lerf_lut_ft_forward(const float* weight, int oC, const float* img, int planes, int h, int w, char mode,
                        int lsb_like_reference, float* out, lerf_stream_t stream);
int lerf_lut_ft_backward(const float* grad_out, int oC, const float* img, int planes, int h, int w, char mode,
                         int lsb_like_reference, float* grad_weight, lerf_stream_t stream);
int lerf_resize_sr_f32_backward(int kind, const lerf_sr_plan_t* plan, const float* img, const float* h0, const float* h1,
                                const float* h2, int planes, float max_sigma, const float* grad_out, float* grad_img,
                                float* grad_h0, float* grad_h1, float* grad_h2, lerf_stream_t stream);

/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
long long lerf_launch_count(void);
void lerf_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* LERF_B200_H_ */
