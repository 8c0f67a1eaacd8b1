"""Shared helpers for the tests: golden fixtures, shipped LUTs, seeded synthetic inputs."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SET5 = ["baby", "bird", "butterfly", "head", "woman"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def lut_dir(model):
    return os.path.join(GOLDEN, "luts", model)


def sha12(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


def uniform_image(seed, h, w, c=3):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, c)).astype(np.uint8)


def natural_image(seed, h, w, c=3):
    """Smooth + texture + noise field (SURVEY.md 8d 'natural-like'), numpy only."""
    rng = np.random.default_rng(seed)

    def blur(x, sigma):
        r = max(1, int(3 * sigma))
        k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
        k /= k.sum()
        x = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 0, x)
        x = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 1, x)
        return x / (x.std() + 1e-12)

    out = np.empty((h, w, c), dtype=np.uint8)
    for ch in range(c):
        f = 128 + 60 * blur(rng.standard_normal((h, w)), 6.0) + 25 * blur(rng.standard_normal((h, w)), 1.5) \
            + 4 * rng.standard_normal((h, w))
        out[:, :, ch] = np.clip(np.round(f), 0, 255).astype(np.uint8)
    return out


def random_luts(seed, oC2=3, modes="sct"):
    """Full-range int8 tables (the shipped ones never hit -128)."""
    rng = np.random.default_rng(seed)
    luts = {}
    for m in modes:
        luts["s1_%sr0" % m] = rng.integers(-128, 128, size=(17 ** 4, 1)).astype(np.int8)
        for r in "01":
            luts["s2_%sr%s" % (m, r)] = rng.integers(-128, 128, size=(17 ** 4, oC2)).astype(np.int8)
    return luts


def psnr_y(gt_hwc_u8, out_hwc_u8, shave):
    """PSNR on the Y channel as the reference's eval computes it (common/utils.py:46-76 _rgb2ycbcr,
    :138-151 PSNR with shave_border = scale; eval_lut_sr.py:735-742 crops both to the common size)."""
    gt, out = np.asarray(gt_hwc_u8), np.asarray(out_hwc_u8)
    h, w = min(gt.shape[0], out.shape[0]), min(gt.shape[1], out.shape[1])
    gt, out = gt[:h, :w], out[:h, :w]
    T0 = np.array([0.256788235294118, 0.504129411764706, 0.097905882352941])

    def y(img):
        return np.dot(img.reshape(-1, 3), T0).reshape(img.shape[:2]) + 16

    d = (y(out).astype(np.float32) - y(gt).astype(np.float32))
    if shave > 0:
        d = d[shave:-shave, shave:-shave]
    return float(20 * np.log10(255.0 / np.sqrt(np.mean(np.power(d, 2)))))
