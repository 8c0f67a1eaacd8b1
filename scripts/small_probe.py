#!/usr/bin/env python
"""GPU box: per-kernel times of the path on one 256x256 image (cfg-1) for the stage-2 alternatives."""
import os
import sys
os.environ.setdefault("LERF_B200_EXPERIMENTS", "1")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

dev = torch.device("cuda", 0)
luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
bench.H, bench.W = 256, 256
img = bench.natural_frames_gpu(1, 1234, dev)[0]
L = lp.lib()
sr = lp.LerfSR(luts, 2)
sr.set_shape(256, 256, 3)
out = sr.alloc_out(1, 3, "f32", dev)


def per_kernel(rep=200):
    tot = {}
    for it in range(rep + 10):
        marks = [torch.cuda.Event(enable_timing=True)]
        marks[0].record()
        names = []

        def rec(n):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)
            names.append(n)
        sr(img, out_format="f32", out=out, record=rec)
        torch.cuda.synchronize()
        if it >= 10:
            for k, n in enumerate(names):
                tot[n] = tot.get(n, 0.0) + marks[k].elapsed_time(marks[k + 1])
    return {n: round(v / rep * 1e3, 1) for n, v in tot.items()}


for v2, name in ((0, "pw (production)"), (80, "pw 32x32 tiles"), (83, "pw 32x8 tiles"), (24, "cell 48B"), (70, "max-tap v10")):
    L.lerf_debug_lut_variant(2, v2)
    print("%-18s" % name, per_kernel(), "us (events around each launch, includes launch gaps)")
L.lerf_debug_lut_variant(2, 0)
g = sr.graphed(img.shape)
g.input.copy_(img)
ref = sr(img).clone()
assert torch.equal(g.replay(), ref)
import numpy as np
ts = []
for _ in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("graph replay: %.1f us" % (np.median(ts) * 1e3))
