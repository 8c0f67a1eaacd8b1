#!/usr/bin/env python
"""GPU box: randomised parity sweep of the SR path against the oracle -- random image sizes, scales (integer, near-integer,
anisotropic, large), both models, all output formats, random row bands.  Prints one line per case and a summary; exit code 1
on the first violation of the parity bars (stages bit-exact, fp32 <= 1e-4, uint8 <= 1 LSB, bands bit-identical).

    python tests/tools/fuzz_parity.py [cases] [seed]
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

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261017)
luts = {}
for m, lin in (("g", False), ("l", True)):
    ld = lp.load_lut_dict(lut_dir("lerf-" + m), linear=lin)
    luts[m] = (ld, lp.LutSet(ld, linear=lin), lin)
worst = 0.0
for case in range(n_cases):
    m = "g" if rng.random() < 0.6 else "l"
    ld, ls, lin = luts[m]
    H, W = int(rng.integers(1, 97)), int(rng.integers(1, 130))
    kind = rng.integers(0, 5)
    if kind == 0:
        sh = sw = float(rng.choice([2, 3, 4, 8]))
    elif kind == 1:
        sh, sw = float(rng.choice([2, 3, 4])), float(rng.choice([2, 3, 4, 8]))          # integer but anisotropic
    elif kind == 2:
        sh = sw = float(rng.choice([2, 3, 4])) + float(rng.choice([-1e-4, 1e-4, 1e-9]))  # near-integer
    elif kind == 3:
        sh, sw = float(rng.uniform(1.0, 6.0)), float(rng.uniform(1.0, 6.0))
    else:
        sh, sw = float(rng.uniform(1.0, 1.2)), float(rng.uniform(6.0, 12.0))
    img = rng.integers(0, 256, size=(H, W, 3)).astype(np.uint8) if rng.random() < 0.7 else \
        np.full((H, W, 3), int(rng.integers(0, 256)), dtype=np.uint8)
    ref, rfeat, rcodes = orc.lerf_sr(img, ld, sh, sw, linear=lin)
    sr = lp.LerfSR(ls, sh, sw)
    d = torch.from_numpy(img).cuda()
    out = sr(d, out_format="f32")
    feat, codes = sr.stages(d)
    ok = np.array_equal(feat.cpu().numpy(), rfeat) and np.array_equal(codes.cpu().numpy(), rcodes)
    o = out.cpu().numpy().astype(np.float64)
    fin = np.isfinite(ref)
    ok = ok and o.shape == ref.shape and np.array_equal(np.isfinite(o), fin)
    err = float(np.max(np.abs(o[fin] - ref[fin]))) if fin.any() else 0.0
    u8 = sr(d, out_format="u8_hwc").cpu().numpy()
    lsb = int(np.max(np.abs(u8.astype(int) - orc.to_uint8_hwc(ref).astype(int))))
    oH = out.shape[-2]
    cuts = sorted(set([0, oH] + [int(v) for v in rng.integers(0, oH + 1, size=2)]))
    band = torch.zeros_like(out)
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        sr(d, out_format="f32", rows=(r0, r1), out=band.unsqueeze(0))
    bands_ok = bool(torch.equal(band, out))
    worst = max(worst, err)
    good = ok and err <= 1e-4 and lsb <= 1 and bands_ok
    print("%3d lerf-%s %3dx%-3d x%.6g x%.6g -> %s  fp32 err %.3g, u8 %d LSB, stages %s, bands %s" % (
        case, m, H, W, sh, sw, tuple(ref.shape[1:]), err, lsb, "exact" if ok else "DIFFER", "identical" if bands_ok else "DIFFER"), flush=True)
    if not good:
        print("PARITY VIOLATION in case %d" % case)
        sys.exit(1)
print("all %d cases within the parity bars; worst fp32 max-abs error %.3g" % (n_cases, worst))
