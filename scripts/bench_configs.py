#!/usr/bin/env python
"""GPU box: device time of the hot path on ALL FIVE BASELINE.json configs, one JSON line each.

bench.py stays the headline (cfg-3); this script is the per-config table of DESIGN.md section 7.  Inputs are resident
uint8 CUDA tensors, outputs float32 planar (the headline mode u8 -> f32), CUDA events, >= 3 warm-ups.
``bytes`` is SURVEY 8(d)'s algorithmic figure C*H*W*1 + C*oH*oW*4.

    python scripts/bench_configs.py [cfg1 cfg2 cfg3 cfg4 cfg5]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

dev = torch.device("cuda", 0)
PEAK, _ = bench.measured_peak_gbs()
LUTS = os.path.join(ROOT, "tests", "golden", "luts")


def natural(n, h, w, seed):
    saved = bench.H, bench.W
    bench.H, bench.W = h, w
    try:
        return bench.natural_frames_gpu(n, seed, dev)
    finally:
        bench.H, bench.W = saved


def timeit(fn, rep):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, what, ms, in_samples, out_samples, extra=None):
    byt = in_samples + 4 * out_samples
    line = {"config": name, "workload": what, "ms": round(ms, 4), "out_MPix_per_s": round(out_samples / 3 / ms / 1e3, 1),
            "algorithmic_bytes": byt, "hbm_GBps": round(byt / ms / 1e6, 1), "hbm_frac": round(byt / ms / 1e6 / PEAK, 4)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def cfg1():
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-g")), device=dev)
    img = natural(1, 256, 256, 1234)[0]
    sr = lp.LerfSR(luts, 2)
    out = sr(img)
    ms = timeit(lambda: sr(img, out=out.unsqueeze(0)), 50)
    report("cfg-1", "LeRF-G x2 SR, one 256x256 image (3 launches, latency-bound)", ms, img.numel(), out.numel())
    g = sr.graphed(img.shape)
    g.input.copy_(img)
    assert torch.equal(g.replay(), out)
    ms = timeit(g.replay, 50)
    report("cfg-1/graph", "the same three launches replayed from a CUDA graph (LerfSR.graphed)", ms, img.numel(), out.numel())
    ms = timeit(lambda: g(img), 50)
    report("cfg-1/graph+copy", "graph replay incl. the copy of the image into the graph's static input", ms, img.numel(), out.numel())


def cfg2():
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-l"), linear=True), linear=True, device=dev)
    imgs = torch.cat([natural(1, 512, 512, 2000 + i) for i in range(16)])
    sr = lp.LerfSR(luts, 3.5)
    out = sr(imgs)
    ms = timeit(lambda: sr(imgs, out=out), 20)
    report("cfg-2", "LeRF-L x3.5 SR, batch of 16 512x512 images", ms, imgs.numel(), out.numel())


def cfg3():
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-g")), device=dev)
    imgs = natural(8, 1356, 2040, 3000)
    sr = lp.LerfSR(luts, 4)
    out = sr(imgs)
    ms = timeit(lambda: sr(imgs, out=out), 20)
    report("cfg-3", "LeRF-G x4 SR, 8 frames 2040x1356 (the bench.py step)", ms, imgs.numel(), out.numel())


def cfg4():
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-g")), device=dev)
    img = natural(1, 1024, 1024, 4000)[0]
    wp = lp.LerfWarp(luts)
    mats = []
    for cls, lo, hi, canvas in (("isc", 2.0, 4.0, 3072), ("osc", 4.0, 9.5, 8192)):
        tot, outs, Ms = 0.0, 0, []
        for k in range(8):  # 8 matrices per class from seed 4000+k
            rng = np.random.default_rng(4000 + k)
            a, d = rng.uniform(lo, hi, 2)
            b, c = rng.uniform(-0.15, 0.15, 2) * max(a, d)
            gh = rng.uniform(-0.6, 0.6, 2) / 1024
            M = np.array([[a, b, 0.0], [c, d, 0.0], [gh[0], gh[1], 1.0]])
            corners = np.array([[0, 0, 1], [1024, 0, 1], [0, 1024, 1], [1024, 1024, 1]], dtype=np.float64).T
            w = M @ corners
            w = w[:2] / w[2]
            M = np.array([[1, 0, canvas / 2 - w[0].mean()], [0, 1, canvas / 2 - w[1].mean()], [0, 0, 1.0]]) @ M
            Ms.append(M)
            out, mask = wp(img, M, (canvas, canvas))
            tot += timeit(lambda: wp(img, M, (canvas, canvas)), 5)
            outs += out.numel()
        report("cfg-4/" + cls, "LeRF-G homographic warp 1024x1024 -> %dx%d canvas, mean of 8 random homographies "
               "(stages + mask + warp, output allocated per call)" % (canvas, canvas), tot / 8, img.numel(), outs // 8)
        mats.append((cls, canvas, Ms))
    for cls, canvas, Ms in mats:  # the 8 homographies as ONE batch: stages once over the batch, outputs allocated once
        imgs = img.unsqueeze(0).expand(8, -1, -1, -1).contiguous()
        out, masks = wp.batch(imgs, Ms, (canvas, canvas))
        one, m1 = wp(img, Ms[3], (canvas, canvas))
        assert torch.equal(torch.nan_to_num(out[3]), torch.nan_to_num(one)) and torch.equal(masks[3], m1)
        ms = timeit(lambda: wp.batch(imgs, Ms, (canvas, canvas), out=out, masks=masks), 5)
        report("cfg-4/" + cls + "/batch", "the same 8 warps as one LerfWarp.batch call (per image)", ms / 8, img.numel(), out.numel() // 8)


def tile():
    """The tile kernel on Gaussian (LeRF-G) non-integer scales, 8 frames 2040x1356, against the x4 cell-owner kernel."""
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-g")), device=dev)
    imgs = natural(8, 1356, 2040, 3000)
    feat, codes = lp.LerfSR(luts, 4).stages(imgs)
    for sc, force in ((4, 0), (4, 3), (4, 2), (3.5, 3), (3.5, 2), (3.2, 3), (3.2, 2), (2.5, 3), (2.5, 2), (1.5, 3), (1.5, 2), ((1.5, 2.0), 3), ((1.5, 2.0), 2)):
        rs = lp.SteeringGaussianResize2d(support_sz=2, max_sigma=10)
        rs.set_shape([3, 1356, 2040], scale_factors=list(sc) if isinstance(sc, tuple) else [sc, sc])
        lp.lib().lerf_debug_force_generic(force)
        out = rs.resize_codes(feat, codes)
        ms = timeit(lambda: rs.resize_codes(feat, codes, out=out), 10)
        lp.lib().lerf_debug_force_generic(0)
        name = "tile" if force == 2 else ("cell-owner (integer scale)" if force == 0 else "cell (any scale)")
        print(json.dumps({"config": "resampler only, LeRF-G x%s, %s kernel" % (sc, name),
                          "ms": round(ms, 4), "G_samples_per_s": round(out.numel() / ms / 1e6, 1)}), flush=True)


def lin():
    """LeRF-L at x4 on 8 frames 2040x1356, resampler only: cell-owner kernel vs tile kernel."""
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-l"), linear=True), linear=True, device=dev)
    imgs = natural(8, 1356, 2040, 3000)
    feat, codes = lp.LerfSR(luts, 4).stages(imgs)
    for sc, force in ((4, 0), (4, 3), (4, 2), (3, 0), (3, 3), (3, 2), (3.5, 3), (3.5, 2), (2.5, 3), (2.5, 2)):
        rs = lp.AmplifiedLinearResize2d()
        rs.set_shape([3, 1356, 2040], scale_factors=[sc, sc])
        lp.lib().lerf_debug_force_generic(force)
        out = rs.resize_codes(feat, codes)
        ms = timeit(lambda: rs.resize_codes(feat, codes, out=out), 10)
        lp.lib().lerf_debug_force_generic(0)
        name = "tile" if force == 2 else ("cell-owner (integer scale)" if force == 0 else "cell (any scale)")
        print(json.dumps({"config": "resampler only, LeRF-L x%s, %s kernel" % (sc, name),
                          "ms": round(ms, 4), "G_samples_per_s": round(out.numel() / ms / 1e6, 1)}), flush=True)


def fixed():
    """Fixed-kernel warps on the cfg-4 in-scale geometry: 1024x1024 uint8 -> 3072x3072 float32."""
    img = natural(1, 1024, 1024, 4000)[0].permute(2, 0, 1).contiguous()
    M = np.array([[3.0, 0.2, 10.0], [-0.15, 2.9, 20.0], [1e-4, -2e-4, 1.0]])
    for cls in (lp.NearestWarp2dNumpy, lp.BilinearWarp2dNumpy, lp.BicubicWarp2dNumpy, lp.Lanczos2Warp2dNumpy, lp.Lanczos3Warp2dNumpy):
        rs = cls()
        rs.set_shape([3, 1024, 1024], M, (3, 3072, 3072))
        out = rs.warp(img)
        ms = timeit(lambda: rs.warp(img), 10)
        report("fixed/" + cls.__name__, "fixed-kernel warp 1024x1024 -> 3072x3072, support %d, float64 weights" % rs.support_sz, ms,
               img.numel(), out.numel())


def png():
    """The device PNG writer (lerf_png_encode_stored) on result images of the benched sizes, against PIL on the host."""
    import io
    import time
    from PIL import Image
    for (h, w) in ((512, 512), (5424, 8160)):
        img = natural(1, h, w, 6000)[0]
        n = lp.png_bytes(h, w, 3)
        buf = torch.empty((n + 3) & ~3, dtype=torch.uint8, device=dev)
        lp.encode_png(img, out=buf)
        ms = timeit(lambda: lp.encode_png(img, out=buf), 10)
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        ms_d2h = timeit(lambda: host.copy_(buf[:n], non_blocking=True), 5)
        arr = img.cpu().numpy()
        t0 = time.perf_counter()
        Image.fromarray(arr).save(io.BytesIO(), format="PNG")
        pil_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"config": "png/%dx%d" % (w, h), "workload": "device PNG writer, stored deflate blocks, RGB", "ms": round(ms, 4),
                          "file_MB": round(n / 1e6, 2), "image_GBps": round(img.numel() / ms / 1e6, 1), "d2h_ms": round(ms_d2h, 3),
                          "pil_host_encode_ms": round(pil_ms, 1)}), flush=True)


def cfg5():
    luts = lp.LutSet(lp.load_lut_dict(os.path.join(LUTS, "lerf-g")), device=dev)
    img = natural(1, 2160, 3840, 5000)[0]
    sr = lp.LerfSR(luts, 8)
    out = sr(img)  # 3 x 17280 x 30720 float32 = 6.4 GB
    oH = out.shape[-2]
    ms = timeit(lambda: sr(img, out=out.unsqueeze(0)), 5)
    report("cfg-5", "LeRF-G x8 SR, one 3840x2160 frame -> 30720x17280, whole frame on ONE GPU", ms, img.numel(), out.numel())
    band = (3 * oH // 8, 4 * oH // 8)
    ms = timeit(lambda: sr(img, out=out.unsqueeze(0), rows=band), 10)
    report("cfg-5/band", "same, ONE of 8 output row bands (what each of 8 GPUs runs; input band + 7-row halo)", ms,
           img.numel() // 8, out.numel() // 8)


if __name__ == "__main__":
    todo = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    for name in todo:
        globals()[name]()
        torch.cuda.empty_cache()
