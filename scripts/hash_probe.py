#!/usr/bin/env python
"""GPU box: stage-1 time per 2040x1356 frame for different block-swizzle weights of the cell tables (lerf_debug_cell_hash)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

dev = torch.device("cuda", 0)
L = lp.lib()
frames = bench.natural_frames_gpu(8, 3000, dev)
ref = None
for hw in ((9, 5, 3), (0, 0, 0), (1, 1, 1), (1, 3, 5), (3, 5, 7), (7, 11, 13), (5, 3, 9), (11, 7, 3), (3, 9, 5), (13, 7, 11), (5, 9, 13), (9, 5, 3)):
    L.lerf_debug_cell_hash(*hw)
    luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
    feat = lp.lut_stage1(luts, frames)
    if ref is None:
        ref = feat.clone()
    assert torch.equal(feat, ref)
    for _ in range(3):
        lp.lut_stage1(luts, frames, out=feat)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        lp.lut_stage1(luts, frames, out=feat)
    b.record()
    torch.cuda.synchronize()
    print("hash %-14s stage 1 %.1f us/frame" % (hw, a.elapsed_time(b) / 10 / 8 * 1e3), flush=True)
    luts.close()
L.lerf_debug_cell_hash(9, 5, 3)
