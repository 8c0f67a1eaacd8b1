"""LUT fine-tuning operators as CUDA autograd functions.

* ``interp_torch_batch`` / ``LutFineTune``  <- SWF2LUT.InterpTorchBatch / .forward / .predict (resample/model.py:172-431):
  4-simplex interpolation of a TRAINABLE float table; the gradient flows to the table (lerf_lut_ft_backward).
* ``steering_gaussian_resize``           <- SteeringGaussianResize2dTorch.resize (resize_right/resize_right2d_torch.py:154-197):
  gradients flow to the image and to the three hyper-parameter maps (lerf_resize_sr_f32_backward).

Both run the C ABI of include/lerf_b200.h; there is no torch fallback.  Geometry is this package's float64-exact one
(``sr_axis_tables``): at integer scales it coincides with the reference's float32 torch grid, at other scales the torch path
picks different taps on columns whose projected coordinate is an exact integer (hard part 3).
"""
import ctypes

import torch

from . import _lib
from ._lib import LERF_KIND_GAUSS
from .lut_interp import mode_pad_dict


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _InterpFn(torch.autograd.Function):
    """out = simplex_interp(weight_q, img) / q.  weight_q: [17^4, oC] float32 in int8 units; img: [B, C, h+pad, w+pad]."""

    @staticmethod
    def forward(ctx, weight_q, img, mode, lsb_like_reference):
        if not weight_q.is_cuda or not img.is_cuda:
            raise ValueError("interp_torch_batch needs CUDA tensors (there is no CPU fallback)")
        w = weight_q.detach().contiguous().float()
        x = img.detach().contiguous().float()
        B, C, hp, wp = x.shape
        pad = mode_pad_dict[mode]
        h, wd = hp - pad, wp - pad
        oC = w.shape[1]
        out = torch.empty((B, C * oC, h, wd), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().lerf_lut_ft_forward(w.data_ptr(), oC, x.data_ptr(), B * C, h, wd, mode.encode(),
                                                      1 if lsb_like_reference else 0, out.data_ptr(), _stream_ptr(x.device)))
        ctx.save_for_backward(x)
        ctx.meta = (mode, bool(lsb_like_reference), tuple(w.shape), h, wd)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        mode, bug, wshape, h, wd = ctx.meta
        B, C = x.shape[0], x.shape[1]
        g = grad_out.contiguous().float()
        gw = torch.zeros(wshape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().lerf_lut_ft_backward(g.data_ptr(), wshape[1], x.data_ptr(), B * C, h, wd, mode.encode(),
                                                       1 if bug else 0, gw.data_ptr(), _stream_ptr(x.device)))
        return gw, None, None, None


def round_func(x):
    """Backward-pass differentiable approximation of round (model.py:164-170): identity in the backward."""
    return x + (torch.round(x) - x).detach()


def interp_torch_batch(weight, outC, mode, img_in, bd=None, interval=4, lsb_like_reference=True):
    """SWF2LUT.InterpTorchBatch(weight, outC, mode, img_in, bd) (model.py:172-385).  ``weight``: float [17^4, outC] in units
    of 1/127 (a trainable parameter); ``img_in``: [B, C, h+bd, w+bd] integer-valued.  Returns [B, C*outC, h, w]."""
    if interval != 4:
        raise ValueError("only interval=4 (17**4 tables) is supported")
    if mode not in mode_pad_dict:
        raise ValueError("Mode {} not implemented.".format(mode))
    if bd is not None and bd != mode_pad_dict[mode]:
        raise ValueError("bd=%r does not match the pad of mode %r" % (bd, mode))
    if weight.shape[1] != outC:
        raise ValueError("weight has %d output channels, outC=%d" % (weight.shape[1], outC))
    wq = torch.clamp(round_func(weight * 127), -127, 127)  # :175-179
    return _InterpFn.apply(wq, img_in, mode, lsb_like_reference)


class LutFineTune(torch.nn.Module):
    """The trainable-LUT model of the reference (SWF2LUT, model.py:132-431) on the CUDA operators above: same parameter
    names (``weight_s<stage>_<mode>r<r>``), same ``forward(x, stage, mode, r)`` and ``predict(x, stage)``.
    ``lut_dict``: the int8 tables as load_lut_dict returns them."""

    def __init__(self, lut_dict, modes="sct", modes2="sct", stages=2, norm=255, outC=3, lsb_like_reference=True):
        super().__init__()
        self.modes, self.modes2, self.stages, self.norm, self.outC = modes, modes2, stages, norm, outC
        self.interval = 4
        self.lsb_like_reference = lsb_like_reference
        for key, tab in lut_dict.items():
            arr = torch.as_tensor(tab).float().reshape(-1, 1 if key.startswith("s1") else outC) / 127.0  # :150,160
            self.register_parameter("weight_" + key, torch.nn.Parameter(arr))

    def forward(self, x, stage, mode, r):
        weight = getattr(self, "weight_s{}_{}r{}".format(stage, mode, r))
        return interp_torch_batch(weight, 1 if stage == 1 else self.outC, mode, x, mode_pad_dict[mode], self.interval,
                                  self.lsb_like_reference)

    def _ensemble(self, x, stage, mode, rots, table_r):
        pred = 0
        pad = mode_pad_dict[mode]
        for r in rots:
            xin = torch.nn.functional.pad(torch.rot90(x, r, [2, 3]), (0, pad, 0, pad), mode="replicate")
            pred = pred + round_func(torch.rot90(self.forward(xin, stage, mode, table_r), (4 - r) % 4, [2, 3]))
        return pred

    def predict(self, x, stage=None):
        """model.py:399-431: x in [0, 1]; stage 2 returns the hyper maps, any other value the stage-1 feature image."""
        x = round_func(x * 255.0)
        if stage == 2:
            pred = 0
            for mode in self.modes2:
                pred = pred + self._ensemble(x, self.stages, mode, [0, 2], 0) + self._ensemble(x, self.stages, mode, [1, 3], 1)
            avg, bias, norm = len(self.modes2) * 4, self.norm // 2, float(self.norm)
            return torch.clamp(round_func(pred / avg + bias), 0, self.norm) / norm
        for s in range(self.stages - 1):
            pred = 0
            for mode in self.modes:
                pred = pred + self._ensemble(x, s + 1, mode, [0, 1, 2, 3], 0)
            if s + 1 == self.stages - 1:
                avg, bias, norm = len(self.modes), 0, 1
            else:
                avg, bias, norm = len(self.modes) * 4, self.norm // 2, float(self.norm)
            x = torch.clamp(round_func(pred / avg) + bias, 0, self.norm) / norm
        return x


class _GaussResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, resizer, img, rho, sigma_x, sigma_y):
        x, h0, h1, h2 = (t.detach().contiguous().float() for t in (img, rho, sigma_x, sigma_y))
        B, C, H, W = x.shape
        if [H, W] != resizer.in_sz:
            raise ValueError("input is %dx%d but set_shape() was given %dx%d" % (H, W, resizer.in_sz[0], resizer.in_sz[1]))
        dev = x.device
        out = torch.empty((B, C) + tuple(resizer.out_sz), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().lerf_resize_sr_f32(LERF_KIND_GAUSS, resizer._get_plan(dev), x.data_ptr(), h0.data_ptr(),
                                                     h1.data_ptr(), h2.data_ptr(), B * C, float(resizer.max_sigma),
                                                     out.data_ptr(), _stream_ptr(dev)))
        ctx.save_for_backward(x, h0, h1, h2)
        ctx.resizer = resizer
        ctx.need = [img.requires_grad, rho.requires_grad, sigma_x.requires_grad, sigma_y.requires_grad]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, h0, h1, h2 = ctx.saved_tensors
        r = ctx.resizer
        dev = x.device
        g = grad_out.contiguous().float()
        grads = [torch.zeros_like(x) if need else None for need in ctx.need]
        ptr = [t.data_ptr() if t is not None else None for t in grads]
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().lerf_resize_sr_f32_backward(LERF_KIND_GAUSS, r._get_plan(dev), x.data_ptr(), h0.data_ptr(),
                                                              h1.data_ptr(), h2.data_ptr(), x.shape[0] * x.shape[1],
                                                              float(r.max_sigma), g.data_ptr(), ptr[0], ptr[1], ptr[2], ptr[3],
                                                              _stream_ptr(dev)))
        return (None,) + tuple(grads)


def steering_gaussian_resize(resizer, input, rho, sigma_x, sigma_y):
    """Differentiable ``resizer.resize(input, rho, sigma_x, sigma_y)`` for [B, C, H, W] CUDA tensors
    (SteeringGaussianResize2dTorch.resize, resize_right2d_torch.py:154-197).  ``resizer``: a SteeringGaussianResize2d after
    ``set_shape``."""
    return _GaussResizeFn.apply(resizer, input, rho, sigma_x, sigma_y)
