#!/usr/bin/env python
"""CPU experiment: how many distinct 128-byte lines does one warp-wide LUT gather touch, per table layout?
A warp covers an 8x4 pixel patch (as the stage kernels do).  Input: the natural-like synthetic plane and a uniform-random one.  Layouts of the 16-byte cell block (stage 1) / 32-byte max-tap block (stage 2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import natural_image, uniform_image, lut_dir  # noqa: E402


def taps(mode, rot):
    out = []
    for k in range(4):
        di, dj = ((k >> 1), (k & 1)) if mode == 0 else ((0, k) if mode == 1 else (k, k))
        dy, dx = [(di, dj), (dj, -di), (-di, -dj), (-dj, di)][rot]
        out.append((dy, dx))
    return out


def patches(a):  # [H, W] -> [n_warps, 32] over 8x4 patches
    H, W = a.shape
    H4, W8 = H // 4 * 4, W // 8 * 8
    return a[:H4, :W8].reshape(H4 // 4, 4, W8 // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)


def distinct(lines):  # [n, 32] -> mean number of distinct values per row
    s = np.sort(lines, axis=1)
    return float((1 + (s[:, 1:] != s[:, :-1]).sum(axis=1)).mean())


def run(plane, name):
    H, W = plane.shape
    pad = np.pad(plane, 3, mode="edge").astype(np.int64)
    res = {}
    for mode in range(3):
        for rot in range(4):
            t = [pad[3 + dy:3 + dy + H, 3 + dx:3 + dx + W] for dy, dx in taps(mode, rot)]
            m = [x >> 4 for x in t]
            l = [x & 15 for x in t]
            a, b, c, d = m
            cell = ((a * 16 + b) * 16 + c) * 16 + d
            t1 = np.argmax(np.stack(l), axis=0)
            lay = {
                "cell16 plain      (line = a,b,c,d>>3)": cell >> 3,
                "cell16 swizzled   (production)": ((cell & ~15) | ((d + 9 * a + 5 * b + 3 * c) & 15)) >> 3,
                "cell16 diagonal   (line = b-a,c-a,d-a,a>>3)": ((((b - a) & 15) * 16 + ((c - a) & 15)) * 16 + ((d - a) & 15)) * 2 + (a >> 3),
                "cell16 2x2x2 cube (line = a>>1,b>>1,c>>1,d)": (((a >> 1) * 8 + (b >> 1)) * 8 + (c >> 1)) * 16 + d,
                "mt32 plain        (line = cell)": cell,
                "mt32 t1-major     (line = t1, cell>>2)": t1 * 65536 + (cell >> 2),
                "mt32 diag t1-major(line = t1,b-a,c-a,d-a,a>>2)": (((t1 * 16 + ((b - a) & 15)) * 16 + ((c - a) & 15)) * 16 + ((d - a) & 15)) * 4 + (a >> 2),
            }
            for k, v in lay.items():
                res.setdefault(k, []).append(distinct(patches(v)))
    print(name)
    for k, v in res.items():
        print("  %-50s mean distinct lines per warp load: %.2f  (s %.2f, c %.2f, t %.2f)" % (
            k, np.mean(v), np.mean(v[0:4]), np.mean(v[4:8]), np.mean(v[8:12])))


if __name__ == "__main__":
    img = natural_image(3000, 256, 1024)
    run(img[:, :, 0], "stage 1, natural-like input")
    run(uniform_image(1, 256, 1024)[:, :, 0], "uniform input")
