#!/usr/bin/env python
"""GPU box: does running the three kernels of different frames concurrently (K streams) beat the batched sequence?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

B = 8
dev = torch.device("cuda", 0)
luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
sr = lp.LerfSR(luts, 4)
frames = bench.natural_frames_gpu(B, 3000, dev)
sr.set_shape(bench.H, bench.W, 3)
out = sr.alloc_out(B, 3, "f32", dev)


def timeit(fn, rep=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / rep


print("batched (3 launches over 8 frames): %.3f ms" % timeit(lambda: sr(frames, out_format="f32", out=out)))
for K in (2, 3, 4, 8):
    for G in (1, 2):  # frames per call
        if B % (K * G) and K * G > B:
            continue
        streams = [torch.cuda.Stream(dev) for _ in range(K)]

        def run():
            cur = torch.cuda.current_stream(dev)
            for s in streams:
                s.wait_stream(cur)
            for n, i in enumerate(range(0, B, G)):
                k = n % K
                with torch.cuda.stream(streams[k]):
                    sr(frames[i:i + G], out_format="f32", out=out[i:i + G], slot=10 + k)
            for s in streams:
                cur.wait_stream(s)

        print("K=%d streams, %d frame(s) per call: %.3f ms" % (K, G, timeit(run)))
