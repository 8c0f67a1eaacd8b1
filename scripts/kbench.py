#!/usr/bin/env python
"""GPU box: per-kernel device time (us per 2040x1356 frame) of the hot path and its tuning variants."""
import os
import sys

os.environ.setdefault("LERF_B200_EXPERIMENTS", "1")  # every tuning variant: liblerf_b200_exp.so

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

B = int(os.environ.get("KB_FRAMES", "4"))
REP = int(os.environ.get("KB_REP", "10"))
ONLY = os.environ.get("KB_ONLY", "")
dev = torch.device("cuda", 0)
luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
luts.pin_l2()
L = lp.lib()


def timeit(fn):
    if os.environ.get("KB_NOTIME"):
        fn()
        return 0.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REP):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / REP / B * 1e3  # us per frame


for kind in os.environ.get("KB_KINDS", "natural,uniform").split(","):
    frames = (bench.natural_frames_gpu if kind == "natural" else bench.uniform_frames_gpu)(B, 3000, dev)
    ref_feat = None
    s1_variants = ((0, "cell (production)"), (27, "cell, one sort/lookup"), (28, "cell paired minb6"), (29, "cell paired minb4"), (1, "row-major minb3"), (25, "cell minb5"), (26, "cell minb6"), (80, "pw minb3"), (81, "pw L1 minb3"), (82, "pw minb4"), (83, "pw L1 minb4"), (90, "cell-pair minb4"), (91, "cell-pair minb3"), (92, "cell-pair minb5"), (93, "cell-pair noalloc"))
    if ONLY == "pw":
        s1_variants = tuple(x for x in s1_variants if x[0] == 0 or x[0] >= 80)
    if ONLY == "s1":
        s1_variants = s1_variants[:4]
    for v, name in (s1_variants[:1] if ONLY in ("prod", "s2") else s1_variants):
        if ONLY == "s1" and v == 0:
            pass
        L.lerf_debug_lut_variant(1, v)
        feat = lp.lut_stage1(luts, frames)
        if ref_feat is None:
            ref_feat = feat.clone()
        assert v in (11, 12) or torch.equal(feat, ref_feat), "stage-1 variants disagree"
        print("%-8s stage1 %-24s %8.1f us/frame" % (kind, name, timeit(lambda: lp.lut_stage1(luts, frames, out=feat))), flush=True)
    L.lerf_debug_lut_variant(1, 0)
    if ONLY == "s1":
        continue
    if ONLY == "s2":
        codes = lp.lut_stage2(luts, ref_feat)
        for v, name in ((0, "production (folded)"), (89, "pw unfolded lookups")):
            L.lerf_debug_lut_variant(2, v)
            c2 = lp.lut_stage2(luts, ref_feat)
            assert torch.equal(c2, codes), "stage-2 variants disagree"
            print("%-8s stage2 %-24s %8.1f us/frame" % (kind, name, timeit(lambda: lp.lut_stage2(luts, ref_feat, out=c2))), flush=True)
        L.lerf_debug_lut_variant(2, 0)
        continue
    if ONLY != "prod":  # shared-memory carve-out of the production stage-1 kernel (percent of the unified L1 / shared memory)
        for pct in (0, 7, 14, 28, 50):
            L.lerf_debug_carveout(pct)
            print("%-8s stage1 cell, carve-out %3d %%     %8.1f us/frame" % (kind, pct, timeit(lambda: lp.lut_stage1(luts, frames, out=feat))), flush=True)
        L.lerf_debug_carveout(-1)
    codes = lp.lut_stage2(luts, ref_feat)
    s2_variants = ((0, "production (pw minb3)"), (70, "max-tap v10"), (1, "row-major minb4"), (23, "cell-48B minb3")) + tuple((60 + k, "max-tap v%d" % k) for k in range(13)) + ((42, "mix rm+max-tap 8/12"), (80, "pw minb3"), (81, "pw L1 minb3"), (82, "pw minb2"), (84, "pw pipelined minb2"), (85, "pw pipelined minb3"), (87, "pw 6-byte exchange minb4"), (88, "pw 6-byte exchange minb3"), (89, "pw unfolded lookups"))
    if ONLY == "pw":
        s2_variants = tuple(x for x in s2_variants if x[0] in (0, 70) or x[0] >= 80)
    for v, name in (s2_variants[:1] if ONLY == "prod" else s2_variants):
        L.lerf_debug_lut_variant(2, v)
        c2 = lp.lut_stage2(luts, ref_feat)
        assert torch.equal(c2, codes), "stage-2 variants disagree"
        print("%-8s stage2 %-24s %8.1f us/frame" % (kind, name, timeit(lambda: lp.lut_stage2(luts, ref_feat, out=c2))), flush=True)
    L.lerf_debug_lut_variant(2, 0)
    if ONLY != "prod":  # access-policy window on the paired-window block instead of the cell block
        for which, name in ((1, "window on pw block"), (0, "window on cell block")):
            L.lerf_debug_l2_window(which)
            luts._pinned_streams.clear()
            luts.pin_l2()
            print("%-8s stage2 production, %-20s %8.1f us/frame" % (kind, name, timeit(lambda: lp.lut_stage2(luts, ref_feat, out=c2))), flush=True)
    if ONLY != "prod":  # block-swizzle weights of the cell tables (baked in at LutSet creation)
        for hw in ():
            L.lerf_debug_cell_hash(*hw)
            l2 = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
            for st, v in ((1, 26), (1, 25), (2, 24), (2, 23)):
                L.lerf_debug_lut_variant(st, v)
                if st == 1:
                    f2 = lp.lut_stage1(l2, frames)
                    assert torch.equal(f2, ref_feat)
                    t = timeit(lambda: lp.lut_stage1(l2, frames, out=f2))
                else:
                    c2 = lp.lut_stage2(l2, ref_feat)
                    assert torch.equal(c2, codes)
                    t = timeit(lambda: lp.lut_stage2(l2, ref_feat, out=c2))
                print("%-8s stage%d cell v%d hash %-12s %8.1f us/frame" % (kind, st, v, hw, t), flush=True)
                L.lerf_debug_lut_variant(st, 0)
            l2.close()
        L.lerf_debug_cell_hash(9, 5, 3)
    rs = lp.SteeringGaussianResize2d(support_sz=2, max_sigma=10)
    rs.set_shape([3, bench.H, bench.W], scale_factors=[4, 4])
    for fmt in ("f32", "u8", "u8_hwc"):
        out = rs.resize_codes(ref_feat, codes, out_format=fmt)
        print("%-8s resize %-20s %8.1f us/frame" % (kind, "int-scale " + fmt, timeit(lambda: rs.resize_codes(ref_feat, codes, out_format=fmt, out=out))), flush=True)
        if fmt == "f32":  # geometry factors from kernel parameters (r1) against immediates (r2)
            L.lerf_debug_resize_variant(11)
            o2 = rs.resize_codes(ref_feat, codes, out_format=fmt)
            assert torch.equal(o2, out), "compile-time geometry changes the result"
            print("%-8s resize %-20s %8.1f us/frame" % (kind, "int-scale f32 r1 geom", timeit(lambda: rs.resize_codes(ref_feat, codes, out_format=fmt, out=o2))), flush=True)
            L.lerf_debug_resize_variant(13)  # weights relative to the smallest exponent (production until r2g)
            o2 = rs.resize_codes(ref_feat, codes, out_format=fmt)
            print("%-8s resize %-20s %8.1f us/frame   max |diff to production| %.3g" % (kind, "int-scale f32 min-tap", timeit(lambda: rs.resize_codes(ref_feat, codes, out_format=fmt, out=o2)), float((o2 - out).abs().max())), flush=True)
            L.lerf_debug_resize_variant(0)
        if fmt != "f32":  # weights relative to the smallest exponent, uint8 flavours
            L.lerf_debug_resize_variant(13)
            o2 = rs.resize_codes(ref_feat, codes, out_format=fmt)
            d = (o2.int() - out.int()).abs()
            print("%-8s resize %-20s %8.1f us/frame   max LSB diff to production %d, differing %.2e" % (kind, "int-scale " + fmt + " min-tap", timeit(lambda: rs.resize_codes(ref_feat, codes, out_format=fmt, out=o2)), int(d.max()), float((d > 0).float().mean())), flush=True)
            L.lerf_debug_resize_variant(0)
        if fmt != "f32":  # the byte-store epilogue of r1 against the staged tile
            L.lerf_debug_resize_variant(10)
            o2 = rs.resize_codes(ref_feat, codes, out_format=fmt)
            d = (o2.int() - out.int()).abs()  # single rounding (r2) against float32-then-integer (r1): rare .5 ties
            assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-4, "uint8 epilogues disagree"
            print("%-8s resize %-20s %8.1f us/frame" % (kind, "int-scale " + fmt + " r1 bytes", timeit(lambda: rs.resize_codes(ref_feat, codes, out_format=fmt, out=o2))), flush=True)
            L.lerf_debug_resize_variant(0)
            if fmt == "u8_hwc":  # tile copied out by lanes (funnel-shifted 128-bit stores) instead of bulk stores
                L.lerf_debug_resize_variant(12)
                o3 = rs.resize_codes(ref_feat, codes, out_format=fmt)
                assert torch.equal(o3, out), "bulk-store copy-out differs"
                print("%-8s resize %-20s %8.1f us/frame" % (kind, "int-scale u8_hwc lanes", timeit(lambda: rs.resize_codes(ref_feat, codes, out_format=fmt, out=o3))), flush=True)
                L.lerf_debug_resize_variant(0)
    if ONLY not in ("prod", "pw"):
        out = rs.resize_codes(ref_feat, codes, out_format="f32")
        for v in (0, 1, 2, 4, 5):
            L.lerf_debug_resize_variant(v)
            print("%-8s resize f32 variant %d           %8.1f us/frame" % (kind, v, timeit(lambda: rs.resize_codes(ref_feat, codes, out_format="f32", out=out))), flush=True)
        L.lerf_debug_resize_variant(0)
    if ONLY not in ("prod", "pw"):
        L.lerf_debug_force_generic(1)
        out = rs.resize_codes(ref_feat, codes)
        print("%-8s resize %-20s %8.1f us/frame" % (kind, "generic f32", timeit(lambda: rs.resize_codes(ref_feat, codes, out=out))), flush=True)
        L.lerf_debug_force_generic(0)
