#!/usr/bin/env python
"""GPU box: LerfSR.run_host (pinned uint8 in -> pinned uint8 HWC out, 8 frames 2040x1356 x4) for different pipeline depths and
row-band counts, against a plain pinned D2H copy of the same bytes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

dev = torch.device("cuda", 0)
luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
frames = bench.natural_frames_gpu(8, 3000, dev).cpu().pin_memory()
sr = lp.LerfSR(luts, 4)
sr.set_shape(bench.H, bench.W, 3)
oH, oW = sr.out_sz
out = torch.empty((8, oH, oW, 3), dtype=torch.uint8).pin_memory()
dbuf = torch.empty((8, oH, oW, 3), dtype=torch.uint8, device=dev)


def timeit(fn, rep=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / rep


ms = timeit(lambda: out.copy_(dbuf, non_blocking=True))
print("plain pinned D2H copy of the 8 results: %.2f ms (%.1f GB/s)" % (ms, out.numel() / ms / 1e6))
mpix = 8 * oH * oW / 1e6
for depth, bands in ((3, 4), (3, 8), (4, 8), (2, 4), (4, 16), (3, 2), (3, 1), (6, 4)):
    ms = timeit(lambda: sr.run_host(frames, out, depth=depth, bands=bands))
    print("depth %d bands %2d: %.2f ms per 8 frames, %.0f MPix/s, D2H %.1f GB/s" % (depth, bands, ms, mpix / ms * 1e3, out.numel() / ms / 1e6), flush=True)
