"""Device-resident LUT set.  Mirrors the loader at resample/eval_lut_sr.py:750-775 (reference)."""
import ctypes
import os

import numpy as np
import torch

from . import _lib

MODES = "sct"  # the shipped/default --modes / --modes2 (common/option.py:21-22); the fused stages specialise on it
ENTRIES = 17 ** 4


def lut_keys():
    keys = ["s1_%sr0" % m for m in MODES]
    for m in MODES:
        keys += ["s2_%sr0" % m, "s2_%sr1" % m]
    return keys


def load_lut_dict(exp_dir, lut_name="LUTft", linear=False, modes=MODES, modes2=MODES, stages=2):
    """Same keys and files as the reference's ``lutDict`` (eval_lut_sr.py:751-775); int8 is kept."""
    if stages != 2:
        raise ValueError("only --stages 2 (the shipped models) is supported")
    luts = {}
    for stage, cur_modes, rots, oC in ((1, modes, "0", 1), (2, modes2, "01", 1 if linear else 3)):
        for m in cur_modes:
            for r in rots:
                path = os.path.join(exp_dir, "{}_s{}_{}r{}.npy".format(lut_name, stage, m, r))
                luts["s{}_{}r{}".format(stage, m, r)] = np.array(np.load(path)).astype(np.int8).reshape(-1, oC)
    return luts


def _as_int8_table(w, oC):
    w = np.asarray(w.detach().cpu().numpy() if isinstance(w, torch.Tensor) else w)
    if w.dtype != np.int8:
        r = np.rint(w)
        if not np.all(r == w) or r.min() < -128 or r.max() > 127:
            raise ValueError("LUT tables must hold int8 values")
        w = r.astype(np.int8)
    w = np.ascontiguousarray(w.reshape(-1, oC))
    if w.shape[0] != ENTRIES:
        raise ValueError("LUT must have 17**4 rows (interval=4), got %d" % w.shape[0])
    return w


class LutSet(object):
    """The nine tables of one model on one GPU (callee-owned handle of the C ABI)."""

    def __init__(self, lut_dict, linear=False, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.linear = bool(linear)
        self.oC = 1 if linear else 3
        tabs = []
        for i, key in enumerate(lut_keys()):
            if key not in lut_dict:
                raise KeyError("LUT dict lacks %r (modes other than 'sct' go through FourSimplexInterpFaster)" % key)
            tabs.append(_as_int8_table(lut_dict[key], 1 if i < 3 else self.oC))
        arr = (ctypes.c_void_p * 9)(*[t.ctypes.data for t in tabs])
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().lerf_luts_create(arr, self.oC, self.device.index or 0, ctypes.byref(handle)))
        self._h = handle
        self._pinned_streams = set()

    @classmethod
    def from_dir(cls, exp_dir, lut_name="LUTft", linear=False, device=None):
        return cls(load_lut_dict(exp_dir, lut_name, linear), linear=linear, device=device)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("LutSet was destroyed")
        return self._h

    def pin_l2(self, stream=None):
        """Put the L2 persisting access-policy window of the LUT block on ``stream`` (once per stream)."""
        s = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        if s not in self._pinned_streams:
            _lib.check(_lib.lib().lerf_luts_pin_l2(self.handle, ctypes.c_void_p(s)))
            self._pinned_streams.add(s)

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().lerf_luts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
