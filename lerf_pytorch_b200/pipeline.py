"""Whole-path engines: the body of ``eltr._worker`` of the reference's eval scripts without file I/O.

* ``LerfSR``   <- resample/eval_lut_sr.py:541-665   (LUT stages -> set_shape -> resize -> uint8 epilogue)
* ``LerfWarp`` <- resample/eval_lut_warp.py:100-222 (LUT stages -> set_shape -> mask -> warp -> uint8 epilogue)

Both take uint8 images as PIL decodes them ([H,W,C] or a batch [B,H,W,C]) that are already CUDA tensors,
and return CUDA tensors; ``LerfSR.run_host`` is the host-to-host entry (pinned uint8 in, pinned uint8 out)
used for end-to-end timing.
"""
import ctypes

import torch

from . import _lib
from ._lib import LERF_KIND_GAUSS, LERF_KIND_LINEAR
from .lut_interp import lut_stage1, lut_stage2
from .luts import LutSet
from .resize_right2d import (AmplifiedLinearResize2d, AmplifiedLinearWarp2d, SteeringGaussianResize2d,
                             SteeringGaussianWarp2d, _FMT)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class LerfSR(object):
    """Arbitrary-scale SR through the LUT path for images of one size."""

    def __init__(self, luts, scale_h, scale_w=None, max_sigma=10, support_sz=2):
        assert isinstance(luts, LutSet)
        self.luts = luts
        self.scale = (float(scale_h), float(scale_h if scale_w is None else scale_w))
        if luts.linear:
            self.resizer = AmplifiedLinearResize2d()  # eval_lut_sr.py:483: constructed without arguments
        else:
            self.resizer = SteeringGaussianResize2d(support_sz=support_sz, max_sigma=max_sigma)
        self.kind = LERF_KIND_LINEAR if luts.linear else LERF_KIND_GAUSS
        self._shape = None
        self._scratch = {}
        self._slots = None

    def set_shape(self, H, W, C=3):
        if self._shape != (H, W, C):
            self.resizer.set_shape([C, H, W], scale_factors=list(self.scale))
            self._shape = (H, W, C)
        return self.resizer.out_sz

    @property
    def out_sz(self):
        return self.resizer.out_sz

    def _get_scratch(self, planes, H, W, device, slot=0):
        need = _lib.lib().lerf_sr_scratch_bytes(planes, self.luts.oC, H, W)
        cur = self._scratch.get(slot)
        if cur is None or cur.numel() < need or cur.device != device:
            cur = self._scratch[slot] = torch.empty(need, dtype=torch.uint8, device=device)
        return cur

    def alloc_out(self, B, C, out_format, device):
        oH, oW = self.out_sz
        if out_format == "f32":
            return torch.empty((B, C, oH, oW), dtype=torch.float32, device=device)
        if out_format == "u8":
            return torch.empty((B, C, oH, oW), dtype=torch.uint8, device=device)
        return torch.empty((B, oH, oW, C), dtype=torch.uint8, device=device)

    def __call__(self, img, out_format="f32", rows=None, out=None, layout="HWC", record=None, slot=0):
        """img: uint8 CUDA [H,W,C] / [B,H,W,C] (layout 'HWC') or [C,H,W] / [B,C,H,W] ('CHW').

        Returns float32 [B,C,oH,oW] ('f32'), uint8 [B,C,oH,oW] ('u8') or uint8 [B,oH,oW,C] ('u8_hwc');
        the batch dimension is dropped when the input had none.  ``rows=(oy0, oy1)`` restricts the work to an
        output row band (the rest of ``out`` is left untouched) -- row-band sharding across GPUs.
        ``record(name)``, if given, is called after each kernel launch (bench.py puts CUDA events there); the
        launches are the same three lerf_sr_fused issues.
        """
        if img.dtype != torch.uint8 or not img.is_cuda:
            raise ValueError("LerfSR needs a uint8 CUDA tensor")
        squeeze = img.dim() == 3
        if squeeze:
            img = img.unsqueeze(0)
        img = img.contiguous()
        if layout == "HWC":
            B, H, W, C = img.shape
            addr = (C, H * W * C, 1, W * C, C)
        else:
            B, C, H, W = img.shape
            addr = (C, C * H * W, H * W, W, 1)
        self.set_shape(H, W, C)
        dev = img.device
        oH, oW = self.out_sz
        if out is None:
            out = self.alloc_out(B, C, out_format, dev)
        oy0, oy1 = (0, oH) if rows is None else rows
        scratch = self._get_scratch(B * C, H, W, dev, slot)
        with torch.cuda.device(dev):
            self.luts.pin_l2()
            if record is not None and rows is None:
                L, st, P = _lib.lib(), _stream_ptr(dev), B * C
                feat = scratch[:P * H * W]
                codes = scratch[(P * H * W + 255) // 256 * 256:]
                _lib.check(L.lerf_lut_stage1(self.luts.handle, img.data_ptr(), P, H, W, addr[0], addr[1], addr[2], addr[3],
                                             addr[4], 0, H, feat.data_ptr(), st))
                record("lut_stage1")
                _lib.check(L.lerf_lut_stage2(self.luts.handle, feat.data_ptr(), P, H, W, 0, H, codes.data_ptr(), st))
                record("lut_stage2")
                _lib.check(L.lerf_resize_sr(self.kind, self.resizer._get_plan(dev), feat.data_ptr(), codes.data_ptr(), P, C,
                                            float(self.resizer.max_sigma), 0, oH, out.data_ptr(), _FMT[out_format], st))
                record("resize_sr")
                return out[0] if squeeze else out
            _lib.check(_lib.lib().lerf_sr_fused(self.luts.handle, self.kind, self.resizer._get_plan(dev), img.data_ptr(),
                                                B * C, addr[0], addr[1], addr[2], addr[3], addr[4],
                                                float(self.resizer.max_sigma), oy0, oy1, scratch.data_ptr(),
                                                out.data_ptr(), _FMT[out_format], _stream_ptr(dev)))
        return out[0] if squeeze else out

    def graphed(self, shape, out_format="f32", layout="HWC"):
        """Small images (the Set5-sized inputs the reference's scripts actually run, cfg-1) are bound by launch latency:
        three launches of a few microseconds of work each.  ``graphed`` captures them ONCE for an input ``shape`` into a
        CUDA graph with static buffers and returns a callable ``g(img) -> out`` that copies ``img`` into the static input
        and replays the graph (``g.input`` / ``g.output`` are the static tensors: fill ``g.input`` yourself and call
        ``g.replay()`` to skip the copy).  Same kernels, same results."""
        return GraphedSR(self, tuple(shape), out_format, layout)

    def run_host(self, host_in, host_out, depth=3, bands=4):
        """Host-to-host entry: ``host_in`` pinned uint8 [B,H,W,C], ``host_out`` pinned uint8 [B,oH,oW,C].

        Frames are pipelined over ``depth`` streams (H2D copy, the three kernels, D2H copy) so PCIe transfers overlap
        compute; each frame is produced in ``bands`` output row bands (``rows=`` of ``__call__``: the stages run on
        the band's input rows + halo), so the first device-to-host copy starts after a quarter of a frame's compute
        instead of a whole frame's -- the path is bound by the D2H link and that start-up is all that is not hidden.
        Returns after everything has landed in ``host_out``.
        """
        B, H, W, C = host_in.shape
        self.set_shape(H, W, C)
        oH, oW = self.out_sz
        dev = self.luts.device
        if self._slots is None or self._slots[0] != (H, W, C, depth):
            slots = []
            for _ in range(depth):
                slots.append((torch.cuda.Stream(dev), torch.empty((H, W, C), dtype=torch.uint8, device=dev),
                              torch.empty((oH, oW, C), dtype=torch.uint8, device=dev)))
            self._slots = ((H, W, C, depth), slots)
        slots = self._slots[1]
        nb = max(1, min(int(bands), oH // 256))
        edges = [oH * k // nb for k in range(nb + 1)]
        cur = torch.cuda.current_stream(dev)
        for st, _, _ in slots:
            st.wait_stream(cur)
        for i in range(B):
            st, d_in, d_out = slots[i % depth]
            with torch.cuda.stream(st):
                d_in.copy_(host_in[i], non_blocking=True)
                for k in range(nb):
                    r0, r1 = edges[k], edges[k + 1]
                    self(d_in, out_format="u8_hwc", out=d_out.unsqueeze(0), slot=1 + i % depth, rows=(r0, r1))
                    host_out[i, r0:r1].copy_(d_out[r0:r1], non_blocking=True)
        for st, _, _ in slots:
            cur.wait_stream(st)
        cur.synchronize()
        return host_out

    def stages(self, img, layout="HWC"):
        """The intermediate products (feat uint8 [P,H,W], codes uint8 [P*oC,H,W]) for parity checks."""
        feat = lut_stage1(self.luts, img, layout)
        return feat, lut_stage2(self.luts, feat)


class GraphedSR(object):
    """A captured ``LerfSR.__call__`` for one input shape (see ``LerfSR.graphed``)."""

    def __init__(self, sr, shape, out_format, layout):
        dev = sr.luts.device
        self.input = torch.empty(shape, dtype=torch.uint8, device=dev)
        batched = len(shape) == 4
        x = self.input if batched else self.input.unsqueeze(0)
        B, C = x.shape[0], (x.shape[3] if layout == "HWC" else x.shape[1])
        H, W = (x.shape[1], x.shape[2]) if layout == "HWC" else (x.shape[2], x.shape[3])
        sr.set_shape(H, W, C)
        out = sr.alloc_out(B, C, out_format, dev)
        self._slot = "graph%d" % id(self)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture: plans, coefficient tables, function attributes
            for _ in range(2):
                sr(x, out_format=out_format, out=out, layout=layout, slot=self._slot)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            sr(x, out_format=out_format, out=out, layout=layout, slot=self._slot)
        self.output = out if batched else out[0]
        self._sr = sr  # keeps the scratch slot and the plan alive

    def replay(self):
        self.graph.replay()
        return self.output

    def __call__(self, img):
        self.input.copy_(img, non_blocking=True)
        return self.replay()


class LerfWarp(object):
    """Homographic warping through the LUT path (one image per call, like the reference)."""

    def __init__(self, luts, max_sigma=10, support_sz=2, border=4, pad_mode="constant"):
        assert isinstance(luts, LutSet)
        self.luts = luts
        self.border = border  # eval_lut_warp.py:36
        if luts.linear:
            self.warper = AmplifiedLinearWarp2d()  # eval_lut_warp.py:38
        else:
            self.warper = SteeringGaussianWarp2d(support_sz=support_sz, max_sigma=max_sigma, pad_mode=pad_mode)

    def batch(self, imgs, matrices, out_hw, out_format="f32", with_mask=True, out=None, masks=None):
        """Several (image, homography) pairs of one input size onto one canvas size: ``imgs`` uint8 CUDA [B,H,W,C],
        ``matrices`` B 3x3 arrays.  The LUT stages run ONCE over the whole batch (one launch each instead of B small
        ones), the warps follow image by image into ``out`` ([B,C,oH,oW] float32 / uint8, or [B,oH,oW,C] for 'u8_hwc')
        and ``masks`` ([B,oH,oW] uint8), allocated once here if not given.  Returns (out, masks)."""
        if imgs.dtype != torch.uint8 or not imgs.is_cuda or imgs.dim() != 4:
            raise ValueError("LerfWarp.batch needs a uint8 CUDA tensor [B,H,W,C]")
        B, H, W, C = imgs.shape
        oH, oW = int(out_hw[0]), int(out_hw[1])
        dev = imgs.device
        self.luts.pin_l2()
        feat = lut_stage1(self.luts, imgs, "HWC")
        codes = lut_stage2(self.luts, feat)
        oC = self.luts.oC
        if out is None:
            if out_format == "u8_hwc":
                out = torch.empty((B, oH, oW, C), dtype=torch.uint8, device=dev)
            else:
                out = torch.empty((B, C, oH, oW), dtype=torch.float32 if out_format == "f32" else torch.uint8, device=dev)
        if with_mask and masks is None:
            masks = torch.empty((B, oH, oW), dtype=torch.uint8, device=dev)
        for b in range(B):
            self.warper.set_shape([C, H, W], matrices[b], [C, oH, oW])
            o = out[b:b + 1] if out_format == "u8_hwc" else out[b]
            self.warper.warp_codes(feat[b * C:(b + 1) * C], codes[b * C * oC:(b + 1) * C * oC], channels=C, out_format=out_format,
                                   with_mask=with_mask, mask_border=self.border, out=o, mask=masks[b] if with_mask else None)
        return out, masks

    def __call__(self, img, matrix, out_hw, out_format="f32", with_mask=True):
        """img uint8 CUDA [H,W,C]; matrix 3x3 input->output; out_hw = (oH, oW) of the target canvas.

        Returns (out, mask): out float32/uint8 [C,oH,oW] (or [oH,oW,C] for 'u8_hwc'), mask uint8 [oH,oW].
        """
        if img.dtype != torch.uint8 or not img.is_cuda or img.dim() != 3:
            raise ValueError("LerfWarp needs a uint8 CUDA tensor [H,W,C]")
        H, W, C = img.shape
        self.luts.pin_l2()
        feat = lut_stage1(self.luts, img, "HWC")
        codes = lut_stage2(self.luts, feat)
        self.warper.set_shape([C, H, W], matrix, [C, int(out_hw[0]), int(out_hw[1])])
        res = self.warper.warp_codes(feat, codes, channels=C, out_format=out_format, with_mask=with_mask,
                                     mask_border=self.border)
        out, mask = res if with_mask else (res, None)
        if out_format == "u8_hwc":
            out = out[0]
        return out, mask
