#!/usr/bin/env python
"""GPU box: role-interleaved pipeline kernel vs the three plain launches (bitwise equality + us per frame)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

B = int(os.environ.get("PB_FRAMES", "8"))
REP = int(os.environ.get("PB_REP", "10"))
dev = torch.device("cuda", 0)
luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
L = lp.lib()
sr = lp.LerfSR(luts, 4)
frames = bench.natural_frames_gpu(B, 3000, dev)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REP):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / REP / B * 1e3


for fmt in os.environ.get("PB_FMTS", "f32,u8_hwc").split(","):
    L.lerf_debug_pipeline(0, 3, 0)
    ref = sr(frames, out_format=fmt).clone()
    out = torch.empty_like(ref)
    print("%-7s three launches                 %8.1f us/frame" % (fmt, timeit(lambda: sr(frames, out_format=fmt, out=out))), flush=True)
    for minb in (3, 4, 2):
        for grp in (0, 1, 2, 3, 6):
            L.lerf_debug_pipeline(1, minb, grp)
            out.zero_()
            sr(frames, out_format=fmt, out=out)
            ok = torch.equal(out, ref)
            t = timeit(lambda: sr(frames, out_format=fmt, out=out))
            print("%-7s pipeline minb %d group %d  %s  %8.1f us/frame" % (fmt, minb, grp, "bitwise-equal" if ok else "MISMATCH", t), flush=True)
L.lerf_debug_pipeline(0, 4, 0)
