#!/usr/bin/env python
"""GPU box: wall time of the eval_lut_sr adapter on the Set5 fixtures with the host I/O inline (--io-threads 0, the
reference's structure) and pipelined.  The tables must be identical."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lerf_pytorch_b200 import eval_lut_sr, eval_lut_warp  # noqa: E402

LUTS = os.path.join(ROOT, "tests", "golden", "luts", "lerf-g")
DATA = os.path.join(ROOT, "tests", "golden", "data")
ref = None
for name, mod, sub in (("eval_lut_sr", eval_lut_sr, "rrBenchmark"), ("eval_lut_warp", eval_lut_warp, "WarpBenchmark")):
    tables = {}
    for threads in (0, 4, 0, 4):
        with tempfile.TemporaryDirectory() as tmp:
            t = time.perf_counter()
            lines, _ = mod.main(["-e", LUTS, "--testDir", os.path.join(DATA, sub), "--resultRoot", tmp, "--io-threads", str(threads)])
            dt = time.perf_counter() - t
        tables.setdefault(threads, lines)
        assert lines == tables[0], "the pipelined run changed the table"
        print("%s --io-threads %d: %.2f s wall (Set5, results saved)" % (name, threads, dt), flush=True)
