#!/usr/bin/env python
"""Golden vectors for the fixed-kernel warps: runs the REFERENCE's own classes
(resize_right/resize_right2d_numpy.py:451-494 Bicubic/Bilinear/Lanczos2/Lanczos3Warp2dNumpy on Warp2dNumpy.warp :409-449)
in this container and writes tests/golden/fixed_warp.npz.  The image and the homographies are the ones of warp.npz.

    python tests/golden/make_golden_fixed.py        (needs /root/reference; the GPU box only reads the .npz)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LERF_REFERENCE", "/root/reference")
os.chdir(REF)  # the reference modules do sys.path.insert(0, "./")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from resize_right.resize_right2d_numpy import (  # noqa: E402
    BicubicWarp2dNumpy, BilinearWarp2dNumpy, Lanczos2Warp2dNumpy, Lanczos3Warp2dNumpy)

KERNELS = {"bilinear": BilinearWarp2dNumpy, "bicubic": BicubicWarp2dNumpy, "lanczos2": Lanczos2Warp2dNumpy,
           "lanczos3": Lanczos3Warp2dNumpy}

if __name__ == "__main__":
    w = np.load(os.path.join(HERE, "warp.npz"))
    img = w["img"].astype(np.float32)
    oshape = (3, 46, 42)
    g = {"out_shape": np.array(oshape), "which": np.array([0, 2, 3])}
    for i in (0, 2, 3):  # a mild homography, one partly outside the canvas, one with the input larger than the canvas
        M = w["mats"][i]
        for name, cls in KERNELS.items():
            rs = cls()
            rs.set_shape(img.shape, M, oshape)
            g["%s_%d" % (name, i)] = rs.warp(img)
            g["pad_%s_%d" % (name, i)] = np.array(rs.pad_vec)
            print(name, i, "support", rs.support_sz, "pad", rs.pad_vec, "NaN:", int(np.isnan(g["%s_%d" % (name, i)]).sum()))
    np.savez_compressed(os.path.join(HERE, "fixed_warp.npz"), **g)
