"""CPU check of the cell-packed LUT lookup's bit tricks: the product header lut_cell.cuh is compiled with g++
(host twins of prmt / dp4a) and run over images exactly like the kernel; results must equal the oracle's."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import lerf_oracle as orc
import util

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cell_emul.cpp")
LIB = os.path.join(HERE, "csrc", "libcell_emul.so")
HDR = os.path.join(os.path.dirname(HERE), "lerf_pytorch_b200", "csrc", "lut_cell.cuh")
HDR2 = os.path.join(os.path.dirname(HERE), "lerf_pytorch_b200", "csrc", "lut_mt.cuh")
HDR3 = os.path.join(os.path.dirname(HERE), "lerf_pytorch_b200", "csrc", "lut_pw.cuh")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2), os.path.getmtime(HDR3)):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call([gxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-x", "c++", SRC, "-o", LIB], env=env)
    return ctypes.CDLL(LIB)


def _run(emul, stage, tables, oC, img_chw, hash_w=(9, 5, 3), paired=0):
    P, H, W = img_chw.shape
    tabs = [np.ascontiguousarray(t, dtype=np.int8) for t in tables]
    arr = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    out = np.empty((P * oC, H, W), dtype=np.uint8)
    img = np.ascontiguousarray(img_chw, dtype=np.uint8)
    assert emul.emul_stage_cell(stage, arr, oC, ctypes.c_void_p(img.ctypes.data), P, H, W,
                                hash_w[0], hash_w[1], hash_w[2], paired, ctypes.c_void_p(out.ctypes.data)) == 0
    return out


@pytest.mark.parametrize("oC,kind", [(3, "shipped"), (1, "shipped"), (3, "random"), (1, "random")])
def test_cell_lookup_equals_oracle(emul, oC, kind):
    if kind == "shipped":
        luts = orc.load_luts(util.lut_dir("lerf-g" if oC == 3 else "lerf-l"), linear=(oC == 1))
    else:
        luts = util.random_luts(11 + oC, oC2=oC)  # full int8 range incl. -128
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(37, 29, 3)).astype(np.uint8)
    img[:6, :6] = 255  # msb 15 / lsb 15 corner: vertex rows at index 16
    img[6:10, :8] = (np.arange(8) * 16)[None, :, None]  # all-lsb-zero ties
    feat, codes, _ = orc.lut_stages(img, luts, oC=oC)
    chw = np.ascontiguousarray(np.transpose(img, (2, 0, 1)))
    t1 = [luts["s1_%sr0" % m] for m in "sct"]
    t2 = []
    for m in "sct":
        t2 += [luts["s2_%sr0" % m], luts["s2_%sr1" % m]]
    for hw in ((9, 5, 3), (0, 0, 0), (7, 11, 13)):
        for paired in (0, 1):  # 1: two lookups per 16x2 sorting network (the production form of the oC = 1 kernels)
            assert np.array_equal(_run(emul, 1, t1, 1, chw, hw, paired), feat), (hw, paired)
            assert np.array_equal(_run(emul, 2, t2, oC, feat, hw, paired), codes), (hw, paired)


@pytest.mark.parametrize("kind", ["shipped", "random"])
def test_maxtap_block_lookup_equals_oracle(emul, kind):
    """Stage 2 on the max-tap block format (lut_mt.cuh): one 32-byte block per lookup."""
    luts = orc.load_luts(util.lut_dir("lerf-g"), linear=False) if kind == "shipped" else util.random_luts(31, oC2=3)
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, size=(33, 27, 3)).astype(np.uint8)
    img[:6, :6] = 255
    img[6:10, :8] = (np.arange(8) * 16)[None, :, None]  # all-lsb-zero ties
    img[10:14, :8] = 7                                   # all-equal lsbs
    feat, codes, _ = orc.lut_stages(img, luts, oC=3)
    t2 = []
    for m in "sct":
        t2 += [luts["s2_%sr0" % m], luts["s2_%sr1" % m]]
    tabs = [np.ascontiguousarray(t, dtype=np.int8) for t in t2]
    arr = (ctypes.c_void_p * 6)(*[t.ctypes.data for t in tabs])
    out = np.empty_like(codes)
    f = np.ascontiguousarray(feat)
    for fn in (emul.emul_stage2_maxtap, emul.emul_stage2_maxtap1):  # uint2 taps (r1d) and single-word taps (r1e)
        out[:] = 0
        assert fn(arr, ctypes.c_void_p(f.ctypes.data), 3, f.shape[1], f.shape[2], ctypes.c_void_p(out.ctypes.data)) == 0
        assert np.array_equal(out, codes)


@pytest.mark.parametrize("oC,kind", [(3, "shipped"), (1, "shipped"), (3, "random"), (1, "random")])
def test_paired_window_lookup_equals_oracle(emul, oC, kind):
    """Both stages on the paired-window format (lut_pw.cuh): one sort + one block per window serves two rotations."""
    if kind == "shipped":
        luts = orc.load_luts(util.lut_dir("lerf-g" if oC == 3 else "lerf-l"), linear=(oC == 1))
    else:
        luts = util.random_luts(41 + oC, oC2=oC)
    rng = np.random.default_rng(15)
    img = rng.integers(0, 256, size=(35, 41, 3)).astype(np.uint8)
    img[:6, :6] = 255
    img[6:10, :8] = (np.arange(8) * 16)[None, :, None]
    img[10:14, :8] = 7
    feat, codes, _ = orc.lut_stages(img, luts, oC=oC)
    chw = np.ascontiguousarray(np.transpose(img, (2, 0, 1)))
    t1 = [luts["s1_%sr0" % m] for m in "sct"]
    t2 = []
    for m in "sct":
        t2 += [luts["s2_%sr0" % m], luts["s2_%sr1" % m]]

    def run(stage, tables, oc, src, fold=0):
        tabs = [np.ascontiguousarray(t, dtype=np.int8) for t in tables]
        arr = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        out = np.empty((src.shape[0] * oc, src.shape[1], src.shape[2]), dtype=np.uint8)
        s = np.ascontiguousarray(src)
        assert emul.emul_stage_pw(stage, arr, oc, ctypes.c_void_p(s.ctypes.data), s.shape[0], s.shape[1], s.shape[2],
                                  fold, ctypes.c_void_p(out.ctypes.data)) == 0
        return out

    for fold in (0, 1):  # 1: folded tables (half the order planes, results swapped on a reversed lookup)
        assert np.array_equal(run(1, t1, 1, chw, fold), feat), fold
        assert np.array_equal(run(2, t2, oC, feat, fold), codes), fold
