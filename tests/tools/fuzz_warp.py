#!/usr/bin/env python
"""GPU box: randomised parity sweep of the warp path against the oracle -- random image sizes and homographies (rotation,
anisotropic scale 1..10, shear, perspective, canvases that cut the image), both models.  Bars: validity mask identical,
fp32 <= 1e-4 and uint8 <= 1 LSB inside the mask.  Outside the mask (clipped taps) it only reports how far the two are apart.

    python tests/tools/fuzz_warp.py [cases] [seed]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lerf_pytorch_b200 as lp  # noqa: E402
from oracle import lerf_oracle as orc  # noqa: E402
from util import lut_dir  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4242)
luts = {}
for m, lin in (("g", False), ("l", True)):
    ld = lp.load_lut_dict(lut_dir("lerf-" + m), linear=lin)
    luts[m] = (ld, lp.LutSet(ld, linear=lin), lin)
worst_in, worst_out = 0.0, 0.0
for case in range(n_cases):
    m = "g" if rng.random() < 0.6 else "l"
    ld, ls, lin = luts[m]
    H, W = int(rng.integers(10, 65)), int(rng.integers(10, 65))
    oH, oW = int(rng.integers(20, 260)), int(rng.integers(20, 260))
    th = rng.uniform(-0.6, 0.6)
    sx, sy = rng.uniform(1.0, 10.0, 2)
    sh = rng.uniform(-0.3, 0.3)
    A = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) @ np.array([[1.0, sh], [0.0, 1.0]]) @ np.diag([sx, sy])
    M = np.eye(3)
    M[:2, :2] = A
    M[2, :2] = rng.uniform(-1.0, 1.0, 2) / max(H, W) * 0.5
    c = M @ np.array([W / 2.0, H / 2.0, 1.0])
    c = c[:2] / c[2]
    M = np.array([[1, 0, oW / 2 - c[0] + rng.uniform(-20, 20)], [0, 1, oH / 2 - c[1] + rng.uniform(-20, 20)], [0, 0, 1.0]]) @ M
    img = rng.integers(0, 256, size=(H, W, 3)).astype(np.uint8)
    ref, rmask, _, _ = orc.lerf_warp(img, ld, M, (3, oH, oW), linear=lin)
    wp = lp.LerfWarp(ls)
    out, mask = wp(torch.from_numpy(img).cuda(), M, (oH, oW), out_format="f32")
    u8, _ = wp(torch.from_numpy(img).cuda(), M, (oH, oW), out_format="u8_hwc")
    mask_ok = np.array_equal(mask.cpu().numpy().astype(bool), rmask[0])
    o = out.cpu().numpy().astype(np.float64)
    inside = np.broadcast_to(rmask[0], ref.shape)
    err_in = float(np.max(np.abs(o[inside] - ref[inside]))) if inside.any() else 0.0
    both = np.isfinite(o) & np.isfinite(ref) & ~inside
    err_out = float(np.max(np.abs(o[both] - ref[both]))) if both.any() else 0.0
    nan_diff = int((np.isfinite(o) != np.isfinite(ref)).sum())
    want8 = orc.to_uint8_hwc(np.where(inside, ref, 0.0))
    lsb = int(np.max(np.abs((u8.cpu().numpy() * rmask[0][:, :, None]).astype(int) - want8.astype(int))))
    worst_in, worst_out = max(worst_in, err_in), max(worst_out, err_out)
    print("%3d lerf-%s %2dx%-2d -> %3dx%-3d scale %.1f/%.1f rot %+.2f: mask %s (%d valid), inside err %.3g u8 %d LSB; outside err %.3g, NaN pattern differs at %d" % (
        case, m, H, W, oH, oW, sx, sy, th, "identical" if mask_ok else "DIFFERS", int(rmask[0].sum()), err_in, lsb, err_out, nan_diff), flush=True)
    if not (mask_ok and err_in <= 1e-4 and lsb <= 1):
        print("PARITY VIOLATION in case %d" % case)
        sys.exit(1)
print("all %d cases within the parity bars; worst error inside the mask %.3g, outside %.3g" % (n_cases, worst_in, worst_out))
