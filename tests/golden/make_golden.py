#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules (resample/eval_lut_sr.py,
resize_right/resize_right2d_numpy.py, common/utils.py) without running their
__main__ blocks, feeds them the shipped Set5 fixtures and seeded synthetic
inputs, and stores inputs + outputs as .npz.  Nothing of the reference's source
is copied; the shipped LUT tables (model data, MIT licence) and a few Set5
images are copied as data so the GPU box can run parity tests and the bench.

Every array written here is the output of reference code, never of this repo's
oracle or kernels.
"""
import os
import shutil
import sys
import warnings

import numpy as np

REF = os.environ.get("LERF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

os.chdir(REF)  # the reference modules do sys.path.insert(0, "./")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import torch  # noqa: E402
from PIL import Image  # noqa: E402
from common.utils import PSNR, _rgb2ycbcr, cal_ssim, mPSNR  # noqa: E402
from resample.eval_lut_sr import FourSimplexInterpFaster, mode_pad_dict  # noqa: E402
from resize_right.resize_right2d_numpy import (  # noqa: E402
    AmplifiedLinearResize2dNumpy, AmplifiedLinearWarp2dNumpy, NearestWarp2dNumpy,
    SteeringGaussianResize2dNumpy, SteeringGaussianWarp2dNumpy)

NAMES = ["baby", "bird", "butterfly", "head", "woman"]


def ref_load_luts(exp_dir, linear):
    """eval_lut_sr.py:750-775 verbatim in effect (float32 tables)."""
    lut = {}
    for stage, modes, rots, oC in ((1, "sct", "0", 1), (2, "sct", "01", 1 if linear else 3)):
        for m in modes:
            for r in rots:
                p = os.path.join(exp_dir, "LUTft_s%d_%sr%s.npy" % (stage, m, r))
                lut["s%d_%sr%s" % (stage, m, r)] = np.array(np.load(p)).astype(np.float32).reshape(-1, oC)
    return lut


def ref_lut_stages(img_hwc_f32, lut, oC, modes="sct", modes2="sct"):
    """The loops of eval_lut_sr.py:541-628 driven through the reference's function."""
    img_lr = img_hwc_f32
    pred = 0
    for mode in modes:
        weight = lut["s1_%sr0" % mode]
        pad = mode_pad_dict[mode]
        for r in [0, 1, 2, 3]:
            rot = np.rot90(img_lr, r)
            h, w, _ = rot.shape
            img_in = np.pad(rot, ((0, pad), (0, pad), (0, 0)), mode="edge").transpose((2, 0, 1))
            pred += FourSimplexInterpFaster(weight, img_in, h, w, 4, 4 - r, upscale=1, mode=mode, oC=1)
    img_lr = np.round(np.clip((pred / len(modes)) + 0, 0, 255)).astype(np.float32).transpose((1, 2, 0))
    pred = 0
    for mode in modes2:
        pad = mode_pad_dict[mode]
        for rs, key in (([0, 2], "s2_%sr0"), ([1, 3], "s2_%sr1")):
            weight = lut[key % mode]
            for r in rs:
                rot = np.rot90(img_lr, r)
                h, w, _ = rot.shape
                img_in = np.pad(rot, ((0, pad), (0, pad), (0, 0)), mode="edge").transpose((2, 0, 1))
                pred += FourSimplexInterpFaster(weight, img_in, h, w, 4, 4 - r, upscale=1, mode=mode, oC=oC)
    img_hyper = np.round(np.clip((pred / (len(modes2) * 4)) + 255 // 2, 0, 255)).astype(np.float32) / float(255)
    return img_lr.transpose((2, 0, 1)), img_hyper  # feat [C,H,W] float32, hyper [C*oC,H,W] float32


def ref_sr(img_hwc_u8, lut, linear, sh, sw):
    feat, hyper = ref_lut_stages(img_hwc_u8.astype(np.float32), lut, 1 if linear else 3)
    if linear:
        rs = AmplifiedLinearResize2dNumpy()
        rs.set_shape(feat.shape, scale_factors=[sh, sw])
        out = rs.resize(feat, hyper)
    else:
        rs = SteeringGaussianResize2dNumpy(support_sz=2, max_sigma=10)
        rs.set_shape(feat.shape, scale_factors=[sh, sw])
        C = hyper.shape[0]
        out = rs.resize(feat, hyper[list(range(0, C, 3))], hyper[list(range(1, C + 1, 3))],
                        hyper[list(range(2, C + 2, 3))])
    return out, feat, hyper


def ref_warp(img_hwc_u8, lut, linear, matrix, gt_shape):
    feat, hyper = ref_lut_stages(img_hwc_u8.astype(np.float32), lut, 1 if linear else 3)
    if linear:
        rs = AmplifiedLinearWarp2dNumpy()
        rs.set_shape(feat.shape, matrix, gt_shape)
        out = rs.warp(feat, hyper)
    else:
        rs = SteeringGaussianWarp2dNumpy(support_sz=2, max_sigma=10)
        rs.set_shape(feat.shape, matrix, gt_shape)
        C = hyper.shape[0]
        out = rs.warp(feat, hyper[list(range(0, C, 3))], hyper[list(range(1, C + 1, 3))],
                      hyper[list(range(2, C + 2, 3))])
    white = np.array(np.zeros_like(feat))
    h, w = white.shape[-2:]
    white[:, 4:h - 4, 4:w - 4] = 255
    nn = NearestWarp2dNumpy()
    nn.set_shape(feat.shape, matrix, gt_shape)
    mask = nn.warp(white) == 255
    return out, mask, feat, hyper


def codes_of(hyper):
    c = np.round(hyper * 255.0)
    assert np.abs(c - hyper * 255.0).max() < 1e-3
    return c.astype(np.uint8)


def main():
    out_luts = os.path.join(HERE, "luts")
    for model in ("lerf-g", "lerf-l"):
        os.makedirs(os.path.join(out_luts, model), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(REF, "models", model))):
            if f.startswith("LUTft_") and f.endswith(".npy"):
                shutil.copyfile(os.path.join(REF, "models", model, f), os.path.join(out_luts, model, f))
    lut_g = ref_load_luts(os.path.join(REF, "models/lerf-g"), False)
    lut_l = ref_load_luts(os.path.join(REF, "models/lerf-l"), True)

    # ---- 1. single LUT pass: all five modes, both oC, all rotations, random tables ----------
    rng = np.random.default_rng(101)
    g = {}
    tabs = {1: rng.integers(-128, 128, size=(17 ** 4, 1)).astype(np.int8),
            3: rng.integers(-128, 128, size=(17 ** 4, 3)).astype(np.int8)}
    h, w, C = 9, 11, 2
    imgs = {
        "uniform": rng.integers(0, 256, size=(C, h + 3, w + 3)).astype(np.uint8),
        # many LSB ties and the MSB=15 / value 255 corner
        "ties": (rng.integers(0, 16, size=(C, h + 3, w + 3)) * 16 + rng.integers(0, 3, size=(C, h + 3, w + 3)) * 15)
        .clip(0, 255).astype(np.uint8),
    }
    g["table_oc1"], g["table_oc3"] = tabs[1], tabs[3]
    for k, v in imgs.items():
        g["img_" + k] = v
    for iname, img in imgs.items():
        for mode in "sdyct":
            pad = mode_pad_dict[mode]
            for oC in (1, 3):
                for rot in (0, 1, 2, 3):
                    o = FourSimplexInterpFaster(tabs[oC].astype(np.float32), img[:, :h + pad, :w + pad].astype(np.float32),
                                                h, w, 4, rot, upscale=1, mode=mode, oC=oC)
                    g["out_%s_%s_oc%d_rot%d" % (iname, mode, oC, rot)] = np.ascontiguousarray(o)
    np.savez_compressed(os.path.join(HERE, "lut_pass.npz"), **g)

    # ---- 2. LUT stages on the Set5 x4 LR fixtures + seeded synthetic images ------------------
    g = {}
    set5 = {}
    for n in NAMES:
        set5[n] = np.array(Image.open(os.path.join(REF, "data/rrBenchmark/Set5/LR_bicubic/rrLR_X4.00_4.00", n + ".png")))
        g["in_" + n] = set5[n]
    rng = np.random.default_rng(202)
    g["in_rand37x29"] = rng.integers(0, 256, size=(37, 29, 3)).astype(np.uint8)
    g["in_rand5x4"] = rng.integers(0, 256, size=(5, 4, 3)).astype(np.uint8)
    g["in_rand1x9"] = rng.integers(0, 256, size=(1, 9, 3)).astype(np.uint8)
    g["in_rand8x1"] = rng.integers(0, 256, size=(8, 1, 3)).astype(np.uint8)
    g["in_gray31x33"] = np.repeat(rng.integers(0, 256, size=(31, 33, 1)).astype(np.uint8), 1, axis=2)
    for key in [k for k in list(g) if k.startswith("in_")]:
        img = g[key]
        for model, lut, linear in (("g", lut_g, False), ("l", lut_l, True)):
            feat, hyper = ref_lut_stages(img.astype(np.float32), lut, 1 if linear else 3)
            g["feat_%s_%s" % (model, key[3:])] = feat.astype(np.uint8)
            g["codes_%s_%s" % (model, key[3:])] = codes_of(hyper)
    np.savez_compressed(os.path.join(HERE, "lut_stages.npz"), **g)

    # ---- 3. resamplers alone: random image + random hyper codes, many scales -----------------
    g = {}
    rng = np.random.default_rng(303)
    H, W = 13, 17
    img = rng.integers(0, 256, size=(3, H, W)).astype(np.float32)
    codes = rng.integers(0, 256, size=(9, H, W)).astype(np.uint8)
    hyper = codes.astype(np.float32) / float(255)
    g["img"], g["codes"] = img.astype(np.uint8), codes
    scales = [(2, 2), (3, 3), (4, 4), (8, 8), (3.5, 3.5), (1.5, 1.5), (2, 3), (2.4, 1.7), (1, 1), (4, 2.5)]
    g["scales"] = np.array(scales, dtype=np.float64)
    for i, (sh, sw) in enumerate(scales):
        rs = SteeringGaussianResize2dNumpy(support_sz=2, max_sigma=10)
        rs.set_shape(img.shape, scale_factors=[sh, sw])
        g["gauss_%d" % i] = rs.resize(img, hyper[0::3], hyper[1::3], hyper[2::3])
        rl = AmplifiedLinearResize2dNumpy()
        rl.set_shape(img.shape, scale_factors=[sh, sw])
        g["linear_%d" % i] = rl.resize(img, hyper[0:3])
    # max_sigma variants (train scripts use other values; the kernel takes it as a parameter)
    rs = SteeringGaussianResize2dNumpy(support_sz=2, max_sigma=4)
    rs.set_shape(img.shape, scale_factors=[3, 3])
    g["gauss_ms4"] = rs.resize(img, hyper[0::3], hyper[1::3], hyper[2::3])
    np.savez_compressed(os.path.join(HERE, "resize_sr.npz"), **g)

    # ---- 4. warps alone: random image + codes, shipped homographies + synthetic ones ---------
    g = {}
    rng = np.random.default_rng(404)
    H, W = 24, 20
    img = rng.integers(0, 256, size=(3, H, W)).astype(np.float32)
    codes = rng.integers(0, 256, size=(9, H, W)).astype(np.uint8)
    hyper = codes.astype(np.float32) / float(255)
    g["img"], g["codes"] = img.astype(np.uint8), codes
    mats = []
    for s in ("isc", "osc"):
        for n in NAMES:
            g["set5_%s_%s" % (s, n)] = torch.load(os.path.join(REF, "data/WarpBenchmark/Set5", s, n + ".pth")).numpy()
    # synthetic homographies input(24x20) -> canvas(70x64)
    mats.append(np.array([[3.0, 0.2, 2.0], [-0.1, 2.7, 3.0], [0.0004, -0.0003, 1.0]]))
    mats.append(np.array([[2.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.0, 0.0, 1.0]]))         # pure x2, exact-integer coordinates
    mats.append(np.array([[1.7, -0.6, 20.0], [0.5, 2.2, -4.0], [0.002, 0.001, 0.95]]))  # partly outside the canvas
    mats.append(np.array([[4.5, 0.3, -30.0], [0.2, 5.0, -25.0], [-0.001, 0.0015, 1.1]]))  # input larger than canvas
    g["mats"] = np.stack(mats)
    oshape = (3, 70, 64)
    g["out_shape"] = np.array(oshape)
    for i, M in enumerate(mats):
        rs = SteeringGaussianWarp2dNumpy(support_sz=2, max_sigma=10)
        rs.set_shape(img.shape, M, oshape)
        g["gauss_%d" % i] = rs.warp(img, hyper[0::3], hyper[1::3], hyper[2::3])
        rl = AmplifiedLinearWarp2dNumpy()
        rl.set_shape(img.shape, M, oshape)
        g["linear_%d" % i] = rl.warp(img, hyper[0:3])
        white = np.zeros_like(img)
        white[:, 4:H - 4, 4:W - 4] = 255
        nn = NearestWarp2dNumpy()
        nn.set_shape(img.shape, M, oshape)
        g["mask_%d" % i] = (nn.warp(white) == 255)
        g["nearest_%d" % i] = nn.warp(img)
        g["pad_%d" % i] = np.array(rs.pad_vec)
    np.savez_compressed(os.path.join(HERE, "warp.npz"), **g)

    # ---- 5. whole path on Set5 fixtures (known answers of scripts.sh:33-47) -------------------
    g = {}
    hr = {n: np.array(Image.open(os.path.join(REF, "data/rrBenchmark/Set5/HR", n + ".png"))) for n in NAMES}
    g["hr_butterfly"] = hr["butterfly"]
    # SR: butterfly x4 (LeRF-G), bird x2 on the x4 LR input (LeRF-G), woman x3.5 (LeRF-L)
    for tag, n, lut, linear, sh, sw in (("g_butterfly_x4", "butterfly", lut_g, False, 4, 4),
                                        ("g_bird_x2", "bird", lut_g, False, 2, 2),
                                        ("l_woman_x3p5", "woman", lut_l, True, 3.5, 3.5)):
        out, feat, hyper = ref_sr(set5[n], lut, linear, sh, sw)
        g["sr_" + tag] = out
    # per-image PSNR-Y / SSIM of the x4 run for all five images, both models (mean = scripts.sh:33-38)
    for model, lut, linear in (("g", lut_g, False), ("l", lut_l, True)):
        ps = []
        for n in NAMES:
            out, _, _ = ref_sr(set5[n], lut, linear, 4, 4)
            u8 = np.clip(np.round(out).transpose((1, 2, 0)), 0, 255).astype(np.uint8)
            gt = hr[n]
            if gt.shape != u8.shape:
                gt = gt[:u8.shape[0], :u8.shape[1], :]
                u8 = u8[:gt.shape[0], :gt.shape[1], :]
            y_gt, y_out = _rgb2ycbcr(gt)[:, :, 0], _rgb2ycbcr(u8)[:, :, 0]
            ps.append([PSNR(y_gt, y_out, 4), cal_ssim(y_gt, y_out)])
        g["set5_x4_psnr_ssim_" + model] = np.array(ps)
        print("Set5 x4 %s mean PSNR/SSIM: %.2f/%.4f" % (model, np.mean(np.array(ps)[:, 0]), np.mean(np.array(ps)[:, 1])))
    # warp: butterfly isc + osc (LeRF-G), woman osc (LeRF-L)
    for tag, s, n, lut, linear in (("g_isc_butterfly", "isc", "butterfly", lut_g, False),
                                   ("g_osc_butterfly", "osc", "butterfly", lut_g, False),
                                   ("l_osc_woman", "osc", "woman", lut_l, True)):
        img = np.array(Image.open(os.path.join(REF, "data/WarpBenchmark/Set5", s, n + ".png")))
        M = torch.load(os.path.join(REF, "data/WarpBenchmark/Set5", s, n + ".pth")).numpy()
        gt = hr[n].transpose((2, 0, 1))
        out, mask, feat, hyper = ref_warp(img, lut, linear, M, gt.shape)
        g["warp_in_" + tag] = img
        g["warp_M_" + tag] = M
        g["warp_gt_shape_" + tag] = np.array(gt.shape)
        g["warp_out_" + tag] = out
        g["warp_mask_" + tag] = mask
        u8 = np.clip(np.round(out).transpose((1, 2, 0)), 0, 255).astype(np.uint8)
        mp = mPSNR(torch.Tensor(u8), torch.Tensor(hr[n]), torch.Tensor(np.array(mask.transpose((1, 2, 0)))))
        g["warp_mpsnr_" + tag] = np.array(float(mp))
        if n == "woman":
            g["hr_woman"] = hr[n]
        print("warp", tag, "mPSNR %.3f" % float(mp), "NaN samples:", int(np.isnan(out).sum()))
    np.savez_compressed(os.path.join(HERE, "set5_path.npz"), **g)
    copy_set5_fixtures()
    for f in sorted(os.listdir(HERE)):
        p = os.path.join(HERE, f)
        if os.path.isfile(p):
            print("%-24s %8.1f KB" % (f, os.path.getsize(p) / 1024))


def copy_set5_fixtures():
    """The Set5 benchmark files (image DATA, 2.4 MB) the eval-script adapters are tested on (tests/test_eval_adapters.py):
    data/rrBenchmark/Set5 and data/WarpBenchmark/Set5, plus the HR folder the warp script expects under WarpBenchmark
    (the reference repo does not ship it there; it is the same five HR images)."""
    dst = os.path.join(HERE, "data")
    for rel in ("rrBenchmark/Set5", "WarpBenchmark/Set5"):
        shutil.copytree(os.path.join(REF, "data", rel), os.path.join(dst, rel), dirs_exist_ok=True)
    shutil.copytree(os.path.join(REF, "data", "rrBenchmark/Set5/HR"), os.path.join(dst, "WarpBenchmark/Set5/HR"), dirs_exist_ok=True)


if __name__ == "__main__":
    main()
