#!/usr/bin/env python
"""GPU box: the paired-window stage kernels against the production kernels (bit for bit) on odd sizes, row bands,
both models, natural / uniform / constant inputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("LERF_B200_EXPERIMENTS", "1")
import lerf_pytorch_b200 as lp  # noqa: E402
import util  # noqa: E402

dev = torch.device("cuda", 0)
L = lp.lib()
rng = np.random.default_rng(1)
bad = 0
for model, linear in (("lerf-g", False), ("lerf-l", True)):
    for lutkind in ("shipped", "random"):
        ld = lp.load_lut_dict(util.lut_dir(model), linear=linear) if lutkind == "shipped" else util.random_luts(77, oC2=1 if linear else 3)
        luts = lp.LutSet(ld, linear=linear, device=dev)
        for (h, w) in ((1, 1), (2, 5), (31, 33), (32, 32), (33, 31), (64, 96), (70, 129), (257, 200), (3, 300), (300, 3)):
            for kind in ("uniform", "natural", "const"):
                if kind == "uniform":
                    img = util.uniform_image(int(rng.integers(1 << 30)), h, w)
                elif kind == "natural":
                    img = util.natural_image(int(rng.integers(1 << 30)), max(h, 8), max(w, 8))[:h, :w]
                else:
                    img = np.full((h, w, 3), int(rng.integers(256)), np.uint8)
                d = torch.from_numpy(np.ascontiguousarray(img)).to(dev)
                L.lerf_debug_lut_variant(1, 0); L.lerf_debug_lut_variant(2, 70 if not linear else 0)
                feat = lp.lut_stage1(luts, d)
                codes = lp.lut_stage2(luts, feat)
                for v in (80, 81, 90, 91, 92):
                    L.lerf_debug_lut_variant(1, v); L.lerf_debug_lut_variant(2, v if (v < 90 or (linear and v < 92)) else 80)
                    f2 = lp.lut_stage1(luts, d)
                    c2 = lp.lut_stage2(luts, feat)
                    ok = torch.equal(f2, feat) and torch.equal(c2, codes)
                    if not ok:
                        bad += 1
                        print("MISMATCH", model, lutkind, h, w, kind, v, int((f2 != feat).sum()), int((c2 != codes).sum()))
        L.lerf_debug_lut_variant(1, 0); L.lerf_debug_lut_variant(2, 0)
        luts.close()
print("pw_check:", "FAILED %d" % bad if bad else "all equal")
