#!/usr/bin/env python
"""One x3.5 LeRF-G resampler launch through the any-scale cell kernel and one through the tile kernel (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lerf_pytorch_b200 as lp  # noqa: E402
from scripts.bench_configs import LUTS, dev, natural  # noqa: E402

luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-g")), device=dev)
imgs = natural(8, 1356, 2040, 3000)
feat, codes = lp.LerfSR(luts, 4).stages(imgs)
sc = float(os.environ.get("SCALE", "3.5"))
rs = lp.SteeringGaussianResize2d(support_sz=2, max_sigma=10)
rs.set_shape([3, 1356, 2040], scale_factors=[sc, sc])
for force in (0, 2, 0, 2):
    lp.lib().lerf_debug_force_generic(force)
    out = rs.resize_codes(feat, codes)
    torch.cuda.synchronize()
lp.lib().lerf_debug_force_generic(0)
