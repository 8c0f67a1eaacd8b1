#!/usr/bin/env python
"""GPU box, under compute-sanitizer: the r2 kernels on small odd-sized inputs (window kernel with its shared-memory
exchange, shuffle / staged uint8 epilogues, fine-tuning forward / backward, graph-free)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lerf_pytorch_b200 as lp  # noqa: E402
import util  # noqa: E402

dev = torch.device("cuda", 0)
for model, linear in (("lerf-g", False), ("lerf-l", True)):
    luts = lp.LutSet(lp.load_lut_dict(util.lut_dir(model), linear=linear), linear=linear, device=dev)
    for (h, w) in ((1, 1), (33, 31), (40, 70)):
        img = torch.from_numpy(util.uniform_image(h * 100 + w, h, w)).to(dev)
        feat = lp.lut_stage1(luts, img)
        codes = lp.lut_stage2(luts, feat)
        lp.lib().lerf_debug_lut_variant(2, 80 if linear else 24)  # the other implementation
        assert torch.equal(lp.lut_stage2(luts, feat), codes)
        lp.lib().lerf_debug_lut_variant(2, 0)
        lp.lib().lerf_debug_lut_variant(1, 27)  # stage 1 with one sort per lookup against the paired 16x2 sort
        assert torch.equal(lp.lut_stage1(luts, img), feat)
        lp.lib().lerf_debug_lut_variant(1, 0)
        for s in (2, 3, 4, 8, 2.5, 3.5, (3.3, 3.9)):
            sr = lp.LerfSR(luts, *(s if isinstance(s, tuple) else (s,)))
            for fmt in ("f32", "u8", "u8_hwc"):
                sr(img, out_format=fmt)
            oH = sr.out_sz[0]
            if oH >= 8:
                sr(img, out_format="u8_hwc", rows=(3, oH - 2))
        lp.lib().lerf_debug_force_generic(3)  # the any-scale cell kernel on a small scale as well
        lp.LerfSR(luts, 1.5, 2.0)(img, out_format="u8_hwc")
        lp.lib().lerf_debug_force_generic(0)
        if h > 1:  # non-default operator parameters: the support kernels of resize and warp
            lp.LerfSR(luts, 2.5, support_sz=4)(img, out_format="f32")
            M = np.array([[1.3, 0.1, -2.0], [-0.05, 1.2, 1.0], [2e-4, -1e-4, 1.0]])
            lp.LerfWarp(luts, support_sz=3, pad_mode="reflect")(img, M, (h + 9, w + 5), out_format="u8_hwc")
            lp.LerfWarp(luts)(img, M, (h + 9, w + 5), out_format="f32")
    luts.close()
for shp in ((1, 1, 3), (37, 41, 3), (3, 21844, 3), (70, 33, 1)):  # device PNG writer: short, odd and block-boundary sizes
    lp.encode_png(torch.randint(0, 256, shp, dtype=torch.uint8, device=dev))
ld = lp.load_lut_dict(util.lut_dir("lerf-g"))
m = lp.LutFineTune(ld).cuda()
x = torch.rand((1, 1, 20, 18), device=dev)
hyper = m.predict(m.predict(x, stage=1) / 255.0, stage=2)
hyper.mean().backward()
rs = lp.SteeringGaussianResize2dTorch(support_sz=2, max_sigma=10)
t = [torch.rand((1, 3, 9, 11), device=dev, requires_grad=True) for _ in range(4)]
rs.set_shape(list(t[0].shape), scale_factors=[2, 2])
rs.resize(t[0] * 255, t[1], t[2], t[3]).sum().backward()
torch.cuda.synchronize()
print("sanitize_probe: done")
