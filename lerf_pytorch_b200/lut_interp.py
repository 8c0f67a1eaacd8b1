"""LUT evaluation: drop-in ``FourSimplexInterpFaster`` and the two rotation-ensembled stages.

Mirrors resample/eval_lut_sr.py:12-18, :24-470 and :541-628 of the reference, backed by the CUDA
kernels in csrc/lut.cu through the C ABI.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from .luts import ENTRIES, LutSet

mode_pad_dict = {"s": 1, "d": 2, "y": 2, "c": 3, "t": 3}  # eval_lut_sr.py:12-18

_table_cache = {}


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _device_table(weight, oC, device):
    """int8 device copy of a LUT, cached per source object (the eval loop passes the same 9 arrays 24 times)."""
    key = (id(weight), oC, str(device))
    hit = _table_cache.get(key)
    if hit is not None and hit[0]() is weight:
        return hit[1]
    w = weight.detach().cpu().numpy() if isinstance(weight, torch.Tensor) else np.asarray(weight)
    r = np.rint(w)
    if not np.all(r == w) or r.min() < -128 or r.max() > 127:
        raise ValueError("LUT tables must hold int8 values")
    t = np.ascontiguousarray(r.astype(np.int8).reshape(-1, oC))
    if t.shape[0] != ENTRIES:
        raise ValueError("LUT must have 17**4 rows (interval=4), got %d" % t.shape[0])
    dev = torch.from_numpy(t).to(device)
    try:
        _table_cache[key] = (weakref.ref(weight), dev)
        if len(_table_cache) > 64:
            _table_cache.pop(next(iter(_table_cache)))
    except TypeError:
        pass
    return dev


def _as_u8_cuda(img, device):
    """Integer-valued samples in 0..255 (float32 in the reference) -> uint8 CUDA tensor."""
    if isinstance(img, torch.Tensor):
        t = img.to(device)
    else:
        t = torch.from_numpy(np.ascontiguousarray(img)).to(device)
    if t.dtype != torch.uint8:
        r = t.round()
        if not bool(((r == t) & (t >= 0) & (t <= 255)).all()):
            raise ValueError("LUT stages need integer-valued samples in 0..255")
        t = r.to(torch.uint8)
    return t.contiguous()


def FourSimplexInterpFaster(weight, img_in, h, w, interval, rot, upscale=4, mode="s", oC=1):
    """Drop-in for eval_lut_sr.py:24-470: one LUT pass over a rotated, edge-padded [C,h+pad,w+pad] image.

    Returns float64 ``[C*oC, h', w']`` (after the final rot90), as a numpy array if ``img_in`` is numpy,
    else a CUDA tensor.  ``upscale`` is ignored, as in the reference.
    """
    if interval != 4:
        raise ValueError("only interval=4 (17**4 tables) is supported")
    if mode not in mode_pad_dict:
        raise ValueError("Mode {} not implemented.".format(mode))  # eval_lut_sr.py:84
    want_numpy = not isinstance(img_in, torch.Tensor)
    device = img_in.device if (not want_numpy and img_in.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    pad = mode_pad_dict[mode]
    img = _as_u8_cuda(img_in, device)
    C, hp, wp = img.shape
    if hp < h + pad or wp < w + pad:
        raise ValueError("img_in is smaller than (h+pad, w+pad)")
    if (hp, wp) != (h + pad, w + pad):
        img = img[:, : h + pad, : w + pad].contiguous()
    tab = _device_table(weight, oC, device)
    out = torch.empty((C * oC, h, w), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().lerf_lut_pass(tab.data_ptr(), img.data_ptr(), C, h, w, mode.encode(), oC,
                                            out.data_ptr(), _stream_ptr(device)))
    res = torch.rot90(out.to(torch.float64) / 16.0, rot, [1, 2])  # eval_lut_sr.py:468-469
    return res.cpu().numpy() if want_numpy else res


def _in_addressing(img, layout):
    """(planes, H, W, channels, batch_stride, chan_stride, row_stride, pix_stride) of a uint8 CUDA tensor."""
    if layout == "HWC":
        if img.dim() == 3:
            img = img.unsqueeze(0)
        B, H, W, C = img.shape
        return img, B * C, H, W, C, H * W * C, 1, W * C, C
    if layout == "CHW":
        if img.dim() == 3:
            img = img.unsqueeze(0)
        B, C, H, W = img.shape
        return img, B * C, H, W, C, C * H * W, H * W, W, 1
    raise ValueError("layout must be 'HWC' or 'CHW'")


def lut_stage1(luts, img_u8, layout="HWC", rows=None, out=None):
    """Stage 1 (eval_lut_sr.py:541-577): uint8 image -> uint8 ``feat`` planar [B*C, H, W]."""
    assert isinstance(luts, LutSet)
    img = img_u8.contiguous()
    if img.dtype != torch.uint8 or not img.is_cuda:
        raise ValueError("lut_stage1 needs a uint8 CUDA tensor")
    img, P, H, W, C, bs, cs, rs, ps = _in_addressing(img, layout)
    y0, y1 = (0, H) if rows is None else rows
    feat = out if out is not None else torch.empty((P, H, W), dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().lerf_lut_stage1(luts.handle, img.data_ptr(), P, H, W, C, bs, cs, rs, ps, y0, y1,
                                              feat.data_ptr(), _stream_ptr(img.device)))
    return feat


def lut_stage2(luts, feat, rows=None, out=None):
    """Stage 2 (eval_lut_sr.py:579-628): ``feat`` [P,H,W] -> uint8 hyper codes [P*oC, H, W] (hyper = codes/255)."""
    assert isinstance(luts, LutSet)
    if feat.dtype != torch.uint8 or not feat.is_cuda or feat.dim() != 3:
        raise ValueError("lut_stage2 needs a uint8 CUDA tensor [P,H,W]")
    feat = feat.contiguous()
    P, H, W = feat.shape
    y0, y1 = (0, H) if rows is None else rows
    codes = out if out is not None else torch.empty((P * luts.oC, H, W), dtype=torch.uint8, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(_lib.lib().lerf_lut_stage2(luts.handle, feat.data_ptr(), P, H, W, y0, y1, codes.data_ptr(),
                                              _stream_ptr(feat.device)))
    return codes


def lut_stages(luts, img_u8, layout="HWC"):
    """Both stages: returns ``(feat [P,H,W] uint8, codes [P*oC,H,W] uint8)``."""
    feat = lut_stage1(luts, img_u8, layout)
    return feat, lut_stage2(luts, feat)
