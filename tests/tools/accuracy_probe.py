#!/usr/bin/env python
"""GPU box: max-abs / percentile error of the fp32 output against the float64 oracle on large frames.
Usage: python tests/tools/accuracy_probe.py [rows]   (rows of a 2040-wide frame; default 1356 = full cfg-3 frame)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lerf_pytorch_b200 as lp  # noqa: E402
from oracle import lerf_oracle as orc  # noqa: E402
from util import lut_dir, natural_image, uniform_image  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1356
ld = lp.load_lut_dict(lut_dir("lerf-g"))
ls = lp.LutSet(ld)
lp.lib().lerf_debug_resize_variant(int(os.environ.get("ACC_VARIANT", "0")))  # int-scale kernel variant under test
for name, img in (("uniform", uniform_image(3000, rows, 2040)), ("natural", natural_image(3001, min(rows, 400), 2040))):
    for S in tuple(int(v) for v in os.environ.get("ACC_SCALES", "4,2").split(",")):
        t = time.time()
        ref, rfeat, rcodes = orc.lerf_sr(img, ld, S, S)
        t = time.time() - t
        sr = lp.LerfSR(ls, S)
        d = torch.from_numpy(img).cuda()
        for force in (0, 1):
            lp.lib().lerf_debug_force_generic(force)
            out = sr(d, out_format="f32").cpu().numpy()
            err = np.abs(out.astype(np.float64) - ref)
            u8 = sr(d, out_format="u8_hwc").cpu().numpy()
            flips = int((u8 != orc.to_uint8_hwc(ref)).sum())
            print("%s x%d %-9s: %9d samples, max %.3g, p99.99 %.3g, mean %.3g, uint8 flips %d (max %d LSB)  [oracle %.1fs]" % (
                name, S, "generic" if force else "int-scale", err.size, err.max(), np.quantile(err, 0.9999), err.mean(), flips,
                int(np.abs(u8.astype(int) - orc.to_uint8_hwc(ref).astype(int)).max()), t), flush=True)
        lp.lib().lerf_debug_force_generic(0)
        feat, codes = sr.stages(d)
        assert np.array_equal(feat.cpu().numpy(), rfeat) and np.array_equal(codes.cpu().numpy(), rcodes)
