"""Runs the UNMODIFIED reference (ddlee-cn/LeRF-PyTorch) LUT SR path on host cores, for bench.py's CPU legs.

The reference is a script collection with no installer.  `stage()` copies the few files of the path from
/root/reference into baseline/_ref/ (git-ignored; it travels to the GPU box with the snapshot):
resample/eval_lut_sr.py, resize_right/*.py, common/{__init__,option,utils}.py.  Nothing here re-implements the
reference: `sr_frame` calls its FourSimplexInterpFaster (resample/eval_lut_sr.py:24-470) in the loops of eltr._worker
(:541-628), then SteeringGaussianResize2dNumpy.set_shape / .resize (:645-661) -- eltr itself cannot be imported as a
library (it reads module globals `opt` and `dataset`, :496-497).  The reference is single-threaded numpy;
`time_workers` runs one independent image band per process, the image-parallel form its `Pool` import hints at.
"""
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC = os.environ.get("LERF_REFERENCE", "/root/reference")
FILES = ["resample/eval_lut_sr.py", "resize_right/__init__.py", "resize_right/interp_methods.py",
         "resize_right/resize_right.py", "resize_right/resize_right2d_numpy.py", "resize_right/resize_right2d_torch.py",
         "common/__init__.py", "common/option.py", "common/utils.py"]


def stage():
    """Copy the reference files of this path into baseline/_ref (only where /root/reference exists: the build container)."""
    if not os.path.isdir(SRC):
        return os.path.isdir(REF_DIR)
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(REF_DIR, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copy2(src, dst)
    return True


_mod = None


def load():
    """Import the reference's modules (no __main__ runs: the guard is at eval_lut_sr.py:747).  None if absent."""
    global _mod
    if _mod is not None:
        return _mod
    root = REF_DIR if os.path.exists(os.path.join(REF_DIR, "resample", "eval_lut_sr.py")) else SRC
    if not os.path.exists(os.path.join(root, "resample", "eval_lut_sr.py")):
        return None
    cwd = os.getcwd()
    try:
        os.chdir(root)  # the reference modules do sys.path.insert(0, "./")
        sys.path.insert(0, root)
        import warnings
        warnings.filterwarnings("ignore")
        from resample.eval_lut_sr import FourSimplexInterpFaster, mode_pad_dict
        from resize_right.resize_right2d_numpy import SteeringGaussianResize2dNumpy
        _mod = (FourSimplexInterpFaster, mode_pad_dict, SteeringGaussianResize2dNumpy, root)
    except Exception:
        _mod = None
    finally:
        os.chdir(cwd)
        if sys.path and sys.path[0] == root:
            sys.path.pop(0)
    return _mod


def load_luts(exp_dir):
    """eval_lut_sr.py:750-775 (LeRF-G: oC = 3 for stage 2)."""
    lut = {}
    for stage_, rots, oC in ((1, "0", 1), (2, "01", 3)):
        for m in "sct":
            for r in rots:
                p = os.path.join(exp_dir, "LUTft_s%d_%sr%s.npy" % (stage_, m, r))
                lut["s%d_%sr%s" % (stage_, m, r)] = np.array(np.load(p)).astype(np.float32).reshape(-1, oC)
    return lut


def sr_frame(img_hwc_u8, lut, scale):
    """The body of eltr._worker for LeRF-G (eval_lut_sr.py:541-661), driven through the reference's own functions."""
    interp, pads, Resizer, _ = load()
    img_lr = np.asarray(img_hwc_u8).astype(np.float32)
    pred = 0
    for mode in "sct":
        for r in range(4):
            rot = np.rot90(img_lr, r)
            h, w, _c = rot.shape
            img_in = np.pad(rot, ((0, pads[mode]), (0, pads[mode]), (0, 0)), mode="edge").transpose((2, 0, 1))
            pred += interp(lut["s1_%sr0" % mode], img_in, h, w, 4, 4 - r, upscale=1, mode=mode, oC=1)
    img_lr = np.round(np.clip(pred / 3 + 0, 0, 255)).astype(np.float32).transpose((1, 2, 0))
    pred = 0
    for mode in "sct":
        for rs, key in (([0, 2], "s2_%sr0"), ([1, 3], "s2_%sr1")):
            for r in rs:
                rot = np.rot90(img_lr, r)
                h, w, _c = rot.shape
                img_in = np.pad(rot, ((0, pads[mode]), (0, pads[mode]), (0, 0)), mode="edge").transpose((2, 0, 1))
                pred += interp(lut[key % mode], img_in, h, w, 4, 4 - r, upscale=1, mode=mode, oC=3)
    hyper = np.round(np.clip(pred / 12 + 255 // 2, 0, 255)).astype(np.float32) / float(255)
    img = img_lr.transpose((2, 0, 1))
    rs_ = Resizer(support_sz=2, max_sigma=10)
    rs_.set_shape(list(img.shape), scale_factors=[scale, scale])
    return rs_.resize(img, hyper[0::3], hyper[1::3], hyper[2::3]), img_lr, hyper


def _work(job):
    os.environ["OMP_NUM_THREADS"] = "1"
    img, lut_dir, scale, repeats = job
    lut = load_luts(lut_dir)
    t0 = time.perf_counter()
    for _ in range(repeats):
        out, _, _ = sr_frame(img, lut, scale)
    return (time.perf_counter() - t0) / repeats, out.shape


def time_workers(imgs, lut_dir, scale, repeats=1):
    """One image per worker process, all at once.  Returns (seconds for the slowest worker, output pixels in total)."""
    import multiprocessing as mp
    if load() is None:
        raise RuntimeError("reference not available (baseline/_ref missing)")
    if len(imgs) == 1:
        dt, shp = _work((imgs[0], lut_dir, scale, repeats))
        return dt, shp[1] * shp[2]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(imgs)) as pool:
        t0 = time.perf_counter()
        res = pool.map(_work, [(im, lut_dir, scale, repeats) for im in imgs])
        wall = (time.perf_counter() - t0) / repeats
    return wall, sum(s[1] * s[2] for _, s in res)
