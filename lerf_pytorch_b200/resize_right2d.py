"""Steerable resampling operators with the reference's class names and call signatures.

Mirrors resize_right/resize_right2d_numpy.py (reference): ``set_shape`` then ``resize`` / ``warp``.
The geometry follows the float64 numpy flavour (the one the LUT eval scripts use); the torch
flavour's names are provided as aliases with the same float64 geometry (SURVEY.md 8c explains why
the reference's fp32 torch grid is not reproduced).  Work is done by csrc/resample.cu.

Inputs may be numpy arrays ([C,H,W], returns numpy float32) or torch tensors ([C,H,W] or [B,C,H,W],
returns a CUDA float32 tensor of the same rank).  ``*_codes`` methods take the uint8 tensors the LUT
stages produce and are the product fast path.
"""
import ctypes
from math import ceil

import numpy as np
import torch

from . import _lib
from ._lib import (LERF_KIND_GAUSS, LERF_KIND_LINEAR, LERF_OUT_F32, LERF_OUT_U8, LERF_OUT_U8_HWC, LERF_WARP_BICUBIC,
                   LERF_WARP_BILINEAR, LERF_WARP_LANCZOS2, LERF_WARP_LANCZOS3, LERF_WARP_NEAREST)

_EPS = np.finfo(np.float32).eps  # resize_right2d_numpy.py:12
_FMT = {"f32": LERF_OUT_F32, "u8": LERF_OUT_U8, "u8_hwc": LERF_OUT_U8_HWC}


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _cuda_device(t=None):
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


def sr_axis_tables(in_sz, out_sz, scale, support_sz=2):
    """One axis of Resize2dNumpy.get_distance (resize_right2d_numpy.py:57-140) in the same operation order.

    Returns (left int32 [out] un-padded first tap, dist float64 [out, supp], (pad0, pad1)).
    """
    o = np.arange(out_sz)
    p = o / float(scale) + (in_sz - 1) / 2 - (out_sz - 1) / (2 * float(scale))       # :70-79
    left = np.int_(np.ceil(p - support_sz / 2 - _EPS))                               # :85-90
    pad0 = -int(left[0])                                                             # :101
    pad1 = int(left[-1]) + (support_sz - 1) - in_sz + 1
    if pad0 < 0 or pad1 < 0:
        raise ValueError("index can't contain negative values")  # what np.pad raises in the reference
    fov = left + pad0                                                                # :102
    pp = p + pad0                                                                    # :103
    dist = np.stack([pp - (fov + k) for k in range(support_sz)], axis=1)             # :131-134
    return left.astype(np.int32), np.ascontiguousarray(dist, dtype=np.float64), (pad0, pad1)


_PAD_MODES = {"constant": 0, "edge": 1, "reflect": 2, "symmetric": 3, "wrap": 4}  # np.pad modes of the IMAGE (LERF_PAD_*)


class _Plan(object):
    def __init__(self, H, W, oH, oW, ly, dy, lx, dx, device, support=2, pad_mode="constant", aa_scale=1.0):
        h = ctypes.c_void_p()
        if support == 2 and pad_mode == "constant" and aa_scale == 1.0:  # what the eval scripts use: every fast path
            _lib.check(_lib.lib().lerf_sr_plan_create(H, W, oH, oW, ly.ctypes.data, dy.ctypes.data, lx.ctypes.data,
                                                      dx.ctypes.data, device.index or 0, ctypes.byref(h)))
        else:  # non-default operator parameters: the float64 support kernel
            _lib.check(_lib.lib().lerf_sr_plan_create_ex(H, W, oH, oW, int(support), ly.ctypes.data, dy.ctypes.data,
                                                         lx.ctypes.data, dx.ctypes.data, _PAD_MODES[pad_mode],
                                                         float(aa_scale), device.index or 0, ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h is not None:
                _lib.lib().lerf_sr_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _prep_f32(x, device):
    """numpy / torch, 3-D or 4-D -> (contiguous float32 CUDA tensor [P,H,W], was_numpy, lead_shape)."""
    was_numpy = not isinstance(x, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(x)) if was_numpy else x
    t = t.to(device=device, dtype=torch.float32)
    lead = tuple(t.shape[:-2])
    return t.reshape((-1,) + tuple(t.shape[-2:])).contiguous(), was_numpy, lead


class Resize2d(object):
    """Resize2dNumpy (resize_right2d_numpy.py:10-140): geometry only."""

    kind = None

    def __init__(self, support_sz=4, device="GPU", pad_mode="constant"):
        if pad_mode not in _PAD_MODES:
            raise NotImplementedError("pad_mode %r: np.pad modes %s are implemented" % (pad_mode, sorted(_PAD_MODES)))
        if int(support_sz) < 1 or int(support_sz) > 64:
            raise ValueError("support_sz must be in [1, 64]")
        self.eps = _EPS
        self.device = device
        self.support_sz = support_sz
        self.pad_mode = pad_mode
        self.antialias = False
        self._plan = None

    def set_shape(self, in_shape, scale_factors=None, out_shape=None):
        in_shape = list(in_shape)
        if len(in_shape) == 4:  # torch flavour passes [B,C,H,W]
            in_shape = in_shape[1:]
            if out_shape is not None and len(out_shape) == 4:
                out_shape = list(out_shape)[1:]
            if isinstance(scale_factors, (list, tuple)) and len(scale_factors) == 4:
                scale_factors = list(scale_factors)[1:]
        self.in_shape = in_shape
        # set_scale_and_out_sz, :25-49
        if out_shape is not None:
            out_shape = list(out_shape) + list(in_shape[len(out_shape):])
            if scale_factors is None:
                scale_factors = [o / i for o, i in zip(out_shape, in_shape)]
        if scale_factors is not None:
            scale_factors = scale_factors if isinstance(scale_factors, (list, tuple)) else [scale_factors, scale_factors]
            scale_factors = [1] * (len(in_shape) - len(scale_factors)) + list(scale_factors)
            if out_shape is None:
                out_shape = [ceil(s * i) for s, i in zip(scale_factors, in_shape)]
        if scale_factors is None:
            raise ValueError("either scale_factors or out_shape is required")
        self.scale_factors = [float(s) for s in scale_factors]
        self.out_shape = out_shape
        self.in_sz = [in_shape[1], in_shape[2]]
        self.out_sz = [out_shape[1], out_shape[2]]
        self._aa_scale = 1.0
        if self.scale_factors[0] < 1.0 or self.scale_factors[1] < 1.0:
            # :51-55.  The reference tests scale_factors[0] (the channel factor, always 1) and [1] (H): a HEIGHT factor below
            # 1 turns antialiasing on FOR GOOD and grows support_sz by 1 / s_h (it compounds over set_shape calls, like in
            # the reference); a width factor below 1 alone does not.
            self.antialias = True
            self.min_scale_factor = min(self.scale_factors[1], self.scale_factors[0])
            self.support_sz = ceil(self.support_sz / self.min_scale_factor)
        if self.antialias and self.kind == LERF_KIND_GAUSS:
            self._aa_scale = float(self.min_scale_factor)  # :186-193 (the amplified-linear resize() has no such branch)
        ly, dy, pad_y = sr_axis_tables(self.in_sz[0], self.out_sz[0], self.scale_factors[1], self.support_sz)
        lx, dx, pad_x = sr_axis_tables(self.in_sz[1], self.out_sz[1], self.scale_factors[2], self.support_sz)
        self.pad_vec = ((0, 0), pad_y, pad_x)  # :129
        self._tables = (ly, dy, lx, dx)
        self._plan = None
        self._plan_dev = None

    def _get_plan(self, device):
        if self._plan is None or self._plan_dev != device:
            ly, dy, lx, dx = self._tables
            with torch.cuda.device(device):
                self._plan = _Plan(self.in_sz[0], self.in_sz[1], self.out_sz[0], self.out_sz[1], ly, dy, lx, dx, device,
                                   int(self.support_sz), self.pad_mode, self._aa_scale)
            self._plan_dev = device
        return self._plan.h

    # ---- fast path: uint8 feat + uint8 hyper codes -----------------------------------------
    def resize_codes(self, feat, codes, channels=3, out_format="f32", rows=None, out=None):
        """feat uint8 [P,H,W], codes uint8 [P*oC,H,W] -> [P,oH,oW] float32/uint8 or [B,oH,oW,channels] uint8."""
        P, H, W = feat.shape
        if [H, W] != self.in_sz:
            raise ValueError("input is %dx%d but set_shape() was given %dx%d" % (H, W, self.in_sz[0], self.in_sz[1]))
        oC = 3 if self.kind == LERF_KIND_GAUSS else 1
        if tuple(codes.shape) != (P * oC, H, W) or codes.dtype != torch.uint8 or feat.dtype != torch.uint8:
            raise ValueError("codes must be uint8 [P*%d,H,W]" % oC)
        oH, oW = self.out_sz
        dev = feat.device
        if out is None:
            if out_format == "f32":
                out = torch.empty((P, oH, oW), dtype=torch.float32, device=dev)
            elif out_format == "u8":
                out = torch.empty((P, oH, oW), dtype=torch.uint8, device=dev)
            else:
                out = torch.empty((P // channels, oH, oW, channels), dtype=torch.uint8, device=dev)
        oy0, oy1 = (0, oH) if rows is None else rows
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().lerf_resize_sr(self.kind, self._get_plan(dev), feat.contiguous().data_ptr(),
                                                 codes.contiguous().data_ptr(), P, channels, float(self.max_sigma),
                                                 oy0, oy1, out.data_ptr(), _FMT[out_format], _stream_ptr(dev)))
        return out

    def _resize_f32(self, input, h0, h1, h2):
        dev = _cuda_device(input)
        img, was_numpy, lead = _prep_f32(input, dev)
        hs = [_prep_f32(h, dev)[0] if h is not None else None for h in (h0, h1, h2)]
        P, H, W = img.shape
        if [H, W] != self.in_sz:
            raise ValueError("input is %dx%d but set_shape() was given %dx%d" % (H, W, self.in_sz[0], self.in_sz[1]))
        for h in hs:
            if h is not None and tuple(h.shape) != (P, H, W):
                raise ValueError("hyper-parameter planes must have the input's shape")
        out = torch.empty((P,) + tuple(self.out_sz), dtype=torch.float32, device=dev)
        p = [h.data_ptr() if h is not None else None for h in hs]
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().lerf_resize_sr_f32(self.kind, self._get_plan(dev), img.data_ptr(), p[0], p[1], p[2], P,
                                                     float(self.max_sigma), out.data_ptr(), _stream_ptr(dev)))
        out = out.reshape(lead + tuple(self.out_sz))
        return out.cpu().numpy() if was_numpy else out


class SteeringGaussianResize2d(Resize2d):
    """SteeringGaussianResize2dNumpy (resize_right2d_numpy.py:142-223)."""

    kind = LERF_KIND_GAUSS

    def __init__(self, support_sz=4, device="GPU", pad_mode="constant", max_sigma=10):
        super().__init__(support_sz, device, pad_mode)
        self.h = 5.0
        self.max_sigma = max_sigma

    def resize(self, input, rho, sigma_x, sigma_y):
        if torch.is_tensor(input) and input.dim() == 4 and int(self.support_sz) == 2 and self.pad_mode == "constant" and \
                not self.antialias and torch.is_grad_enabled() and any(
                torch.is_tensor(t) and t.requires_grad for t in (input, rho, sigma_x, sigma_y)):
            from .lut_finetune import steering_gaussian_resize  # the differentiable torch flavour (fine-tuning, 8f item 4)
            return steering_gaussian_resize(self, input, rho, sigma_x, sigma_y)
        return self._resize_f32(input, rho, sigma_x, sigma_y)


class AmplifiedLinearResize2d(Resize2d):
    """AmplifiedLinearResize2dNumpy (resize_right2d_numpy.py:225-282)."""

    kind = LERF_KIND_LINEAR

    def __init__(self, support_sz=2, device="GPU", pad_mode="constant", max_sigma=1):
        super().__init__(support_sz, device, pad_mode)
        self.h = 5.0
        self.max_sigma = max_sigma

    def resize(self, input, alpha):
        return self._resize_f32(input, alpha, None, None)


def _warp_pad0(minv, in_sz, support_sz):
    """Leading pads of Warp2dNumpy.calc_pad_sz (:363-369): from output pixel (0,0) only."""
    g = np.dot(minv, np.array([[0.0], [0.0], [1.0]]))[:, 0]      # :327 for gridy = (0, 0)
    xi, yi = g[0] / g[2], g[1] / g[2]                           # :330-331
    pr = np.clip(yi, 0, in_sz[0])                               # rows <- y (:335-338)
    pc = np.clip(xi, 0, in_sz[1])
    lr = int(np.ceil(pr - support_sz / 2 - _EPS))
    lc = int(np.ceil(pc - support_sz / 2 - _EPS))
    return max(-lr, 0), max(-lc, 0)


class Warp2d(object):
    """Warp2dNumpy (resize_right2d_numpy.py:284-449): geometry only."""

    kind = None

    def __init__(self, support_sz=4, device="GPU", pad_mode="constant"):
        if pad_mode not in _PAD_MODES:
            raise NotImplementedError("pad_mode=%r: one of %s" % (pad_mode, sorted(_PAD_MODES)))
        if not 1 <= int(support_sz) <= 64:
            raise ValueError("support_sz must be in 1..64")
        self.eps = _EPS
        self.device = device
        self.support_sz = support_sz
        self.pad_mode = pad_mode
        self.antialias = False

    def set_shape(self, in_shape, matrix, out_shape):
        in_shape = list(in_shape)
        out_shape = list(out_shape)
        if len(in_shape) == 4:
            in_shape = in_shape[1:]
        if len(out_shape) == 4:
            out_shape = out_shape[1:]
        self.in_shape = in_shape
        m = matrix.detach().cpu().numpy() if isinstance(matrix, torch.Tensor) else np.asarray(matrix)
        if m.ndim == 3:
            m = m[0]
        self.matrix = m
        out_shape = list(out_shape) + list(in_shape[len(out_shape):])   # :301
        self.out_shape = out_shape
        self.in_sz = [in_shape[1], in_shape[2]]
        self.out_sz = [out_shape[1], out_shape[2]]
        self.minv = np.ascontiguousarray(np.linalg.inv(m), dtype=np.float64)  # :327
        self.pad0 = _warp_pad0(self.minv, self.in_sz, self.support_sz)

    def _default_params(self):
        """True for the parameters eval_lut_warp.py runs with (--suppSize 2, constant padding): the tuned kernels; anything
        else goes through lerf_warp_ex (operation-order float64, any support, np.pad modes)."""
        return self.support_sz == 2 and self.pad_mode == "constant"

    def _check(self, H, W):
        if [H, W] != self.in_sz:
            raise ValueError("input is %dx%d but set_shape() was given %dx%d" % (H, W, self.in_sz[0], self.in_sz[1]))

    def warp_codes(self, feat, codes, channels=3, out_format="f32", with_mask=False, mask_border=4, out=None, mask=None):
        """Fast path: feat uint8 [P,H,W] + codes uint8 [P*oC,H,W]; optionally also the validity mask of
        eval_lut_warp.py:197-204,229 (uint8 [oH,oW], 1 = valid)."""
        P, H, W = feat.shape
        self._check(H, W)
        oH, oW = self.out_sz
        dev = feat.device
        if out is None:
            if out_format == "f32":
                out = torch.empty((P, oH, oW), dtype=torch.float32, device=dev)
            elif out_format == "u8":
                out = torch.empty((P, oH, oW), dtype=torch.uint8, device=dev)
            else:
                out = torch.empty((P // channels, oH, oW, channels), dtype=torch.uint8, device=dev)
        if with_mask and mask is None:
            mask = torch.empty((oH, oW), dtype=torch.uint8, device=dev)
        mp = _warp_pad0(self.minv, self.in_sz, 1) if with_mask else (0, 0)
        L = _lib.lib()
        with torch.cuda.device(dev):
            if self._default_params():
                _lib.check(L.lerf_warp(self.kind, feat.contiguous().data_ptr(), codes.contiguous().data_ptr(), P,
                                       channels, H, W, oH, oW, self.minv.ctypes.data, self.pad0[0], self.pad0[1],
                                       float(self.max_sigma), out.data_ptr(), _FMT[out_format],
                                       mask.data_ptr() if with_mask else None, mp[0], mp[1], mask_border,
                                       _stream_ptr(dev)))
            else:
                _lib.check(L.lerf_warp_ex(self.kind, feat.contiguous().data_ptr(), codes.contiguous().data_ptr(), None,
                                          None, None, None, P, channels, H, W, oH, oW, self.minv.ctypes.data,
                                          int(self.support_sz), _PAD_MODES[self.pad_mode], self.pad0[0], self.pad0[1],
                                          float(self.max_sigma), out.data_ptr(), _FMT[out_format], _stream_ptr(dev)))
                if with_mask:   # the mask is a support-1 nearest warp whatever this operator's support is
                    _lib.check(L.lerf_warp(LERF_KIND_GAUSS, None, None, 0, 1, H, W, oH, oW, self.minv.ctypes.data, 0, 0,
                                           1.0, None, LERF_OUT_F32, mask.data_ptr(), mp[0], mp[1], mask_border,
                                           _stream_ptr(dev)))
        return (out, mask) if with_mask else out

    def _warp_f32(self, input, h0, h1, h2):
        dev = _cuda_device(input)
        img, was_numpy, lead = _prep_f32(input, dev)
        hs = [_prep_f32(h, dev)[0] if h is not None else None for h in (h0, h1, h2)]
        P, H, W = img.shape
        self._check(H, W)
        oH, oW = self.out_sz
        out = torch.empty((P, oH, oW), dtype=torch.float32, device=dev)
        p = [h.data_ptr() if h is not None else None for h in hs]
        with torch.cuda.device(dev):
            if self._default_params():
                _lib.check(_lib.lib().lerf_warp_f32(self.kind, img.data_ptr(), p[0], p[1], p[2], P, H, W, oH, oW,
                                                    self.minv.ctypes.data, self.pad0[0], self.pad0[1],
                                                    float(self.max_sigma), out.data_ptr(), _stream_ptr(dev)))
            else:
                _lib.check(_lib.lib().lerf_warp_ex(self.kind, None, None, img.data_ptr(), p[0], p[1], p[2], P, 1, H, W,
                                                   oH, oW, self.minv.ctypes.data, int(self.support_sz),
                                                   _PAD_MODES[self.pad_mode], self.pad0[0], self.pad0[1],
                                                   float(self.max_sigma), out.data_ptr(), LERF_OUT_F32,
                                                   _stream_ptr(dev)))
        out = out.reshape(lead + (oH, oW))
        return out.cpu().numpy() if was_numpy else out


class SteeringGaussianWarp2d(Warp2d):
    """SteeringGaussianWarp2dNumpy (resize_right2d_numpy.py:496-577)."""

    kind = LERF_KIND_GAUSS

    def __init__(self, support_sz=4, device="GPU", pad_mode="constant", max_sigma=10):
        super().__init__(support_sz, device, pad_mode)
        self.h = 5.0
        self.max_sigma = max_sigma

    def warp(self, input, rho, sigma_x, sigma_y):
        return self._warp_f32(input, rho, sigma_x, sigma_y)


class AmplifiedLinearWarp2d(Warp2d):
    """AmplifiedLinearWarp2dNumpy (resize_right2d_numpy.py:579-635)."""

    kind = LERF_KIND_LINEAR

    def __init__(self, support_sz=2, device="GPU", pad_mode="constant", max_sigma=1):
        super().__init__(support_sz, device, pad_mode)
        self.h = 5.0
        self.max_sigma = max_sigma

    def warp(self, input, alpha):
        return self._warp_f32(input, alpha, None, None)


class NearestWarp2d(Warp2d):
    """NearestWarp2dNumpy (resize_right2d_numpy.py:460-467), as used for the validity mask
    (eval_lut_warp.py:197-204): ``mask(border)`` returns uint8 [oH,oW], 1 where the warped white frame is 255."""

    kind = LERF_KIND_GAUSS

    def __init__(self, support_sz=1, device="GPU", pad_mode="constant"):
        super().__init__(support_sz, device, pad_mode)
        self.max_sigma = 1

    def mask(self, border=4, device=None):
        dev = _cuda_device() if device is None else torch.device(device)
        oH, oW = self.out_sz
        H, W = self.in_sz
        m = torch.empty((oH, oW), dtype=torch.uint8, device=dev)
        mp = _warp_pad0(self.minv, self.in_sz, 1)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().lerf_warp(LERF_KIND_GAUSS, None, None, 0, 1, H, W, oH, oW, self.minv.ctypes.data, 0, 0,
                                            1.0, None, LERF_OUT_F32, m.data_ptr(), mp[0], mp[1], border,
                                            _stream_ptr(dev)))
        return m

    def warp(self, input):
        """Nearest-neighbour warp of an image (Warp2dNumpy.warp with box2d, :409-449, :460-467)."""
        return _fixed_warp(self, LERF_WARP_NEAREST, input)


def _fixed_warp(self, kernel, input):
    """Warp2dNumpy.warp (:409-449) with a fixed separable kernel through lerf_warp_fixed: numpy [C,H,W] in -> numpy float32
    out, torch CUDA [C,H,W] / [B,C,H,W] in -> torch float32 out; uint8 or float32 values."""
    supp = _lib.lib().lerf_warp_fixed_support(kernel)
    if self.support_sz != supp:
        raise NotImplementedError("support_sz=%r: this kernel's support is %d" % (self.support_sz, supp))
    dev = _cuda_device(input if isinstance(input, torch.Tensor) else None)
    was_numpy = not isinstance(input, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(input)) if was_numpy else input
    is_u8 = t.dtype == torch.uint8
    t = t.to(dev).contiguous() if is_u8 else t.to(device=dev, dtype=torch.float32).contiguous()
    lead = tuple(t.shape[:-2])
    H, W = int(t.shape[-2]), int(t.shape[-1])
    if [H, W] != self.in_sz:
        raise ValueError("input is %dx%d but set_shape() was given %dx%d" % (H, W, self.in_sz[0], self.in_sz[1]))
    P = int(np.prod(lead)) if lead else 1
    oH, oW = self.out_sz
    out = torch.empty((P, oH, oW), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().lerf_warp_fixed(kernel, t.data_ptr(), 1 if is_u8 else 0, P, H, W, oH, oW, self.minv.ctypes.data,
                                              self.pad0[0], self.pad0[1], out.data_ptr(), _stream_ptr(dev)))
    out = out.reshape(lead + (oH, oW))
    return out.cpu().numpy() if was_numpy else out


class _FixedWarp2d(Warp2d):
    """Bicubic / Bilinear / Lanczos2 / Lanczos3Warp2dNumpy (resize_right2d_numpy.py:451-494): the baselines of the paper."""

    kernel = None
    default_support = None

    def __init__(self, support_sz=None, device="GPU", pad_mode="constant"):
        super().__init__(self.default_support if support_sz is None else support_sz, device, pad_mode)

    def warp(self, input):
        return _fixed_warp(self, self.kernel, input)


class BilinearWarp2d(_FixedWarp2d):
    kernel, default_support = LERF_WARP_BILINEAR, 2


class BicubicWarp2d(_FixedWarp2d):
    kernel, default_support = LERF_WARP_BICUBIC, 4


class Lanczos2Warp2d(_FixedWarp2d):
    kernel, default_support = LERF_WARP_LANCZOS2, 4


class Lanczos3Warp2d(_FixedWarp2d):
    kernel, default_support = LERF_WARP_LANCZOS3, 6


# The reference's names (numpy flavour is what the LUT eval scripts import; torch flavour = same API on [B,C,H,W])
SteeringGaussianResize2dNumpy = SteeringGaussianResize2dTorch = SteeringGaussianResize2d
AmplifiedLinearResize2dNumpy = AmplifiedLinearResize2dTorch = AmplifiedLinearResize2d
SteeringGaussianWarp2dNumpy = SteeringGaussianWarp2dTorch = SteeringGaussianWarp2d
AmplifiedLinearWarp2dNumpy = AmplifiedLinearWarp2dTorch = AmplifiedLinearWarp2d
NearestWarp2dNumpy = NearestWarp2dTorch = NearestWarp2d
BilinearWarp2dNumpy = BilinearWarp2d
BicubicWarp2dNumpy = BicubicWarp2d
Lanczos2Warp2dNumpy = Lanczos2Warp2d
Lanczos3Warp2dNumpy = Lanczos3Warp2d
