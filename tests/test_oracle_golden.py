"""CPU: pin the oracle (oracle/lerf_oracle.{c,py}) to vectors produced by the reference's own code.

The goldens were written by tests/golden/make_golden.py, which imports and runs
ddlee-cn/LeRF-PyTorch (eval_lut_sr.py, resize_right2d_numpy.py, common/utils.py).
"""
import numpy as np
import pytest

from oracle import lerf_oracle as orc
from util import SET5, golden, lut_dir, sha12

MODE_PAD = {"s": 1, "d": 2, "y": 2, "c": 3, "t": 3}

# sha1[:12] of the uint8 stage outputs on rrLR_X4.00_4.00/<name>.png,
# computed in the survey session with the reference's code (independent of make_golden.py).
SURVEY_HASHES = {
    "g": {"baby": ("d82b654e9e44", "d4f52b53e0df"), "bird": ("0027ebafe4bd", "930b68cb5f49"),
          "butterfly": ("b83ab24faaa7", "a1bd0ee7d511"), "head": ("5cc86a0ae633", "c0a9a8c2db67"),
          "woman": ("ec20d339e49c", "9774945920ef")},
    "l": {"baby": ("d7b12f106f6f", "9fee4cbfbcc3"), "bird": ("d0e7e59f05a3", "9ff286db2068"),
          "butterfly": ("f216b391249d", "6ecae2efdbae"), "head": ("2f1b2d38d3c0", "ba4ee3ce6d6d"),
          "woman": ("d310ad022e96", "c9c63a5d0c70")},
}
LUT_HASHES = {
    "lerf-g": {"s1_c": "2b1207aed1cf", "s1_s": "4bd46ec51f61", "s1_t": "749de4454775", "s2_cr0": "fe0cf6bed09e",
               "s2_cr1": "05469c5ed7a4", "s2_sr0": "6488ea566a1e", "s2_sr1": "87d0b374efc4",
               "s2_tr0": "cefef21923a0", "s2_tr1": "0d5da1c0ab9b"},
    "lerf-l": {"s1_c": "4c23cb89854f", "s1_s": "60b0996980bc", "s1_t": "50b0b3f484d8", "s2_cr0": "1e4f84c8d0aa",
               "s2_cr1": "6b1b770dc296", "s2_sr0": "faecb4ed3ef6", "s2_sr1": "669f9f9dbc6b",
               "s2_tr0": "4ef868f7d508", "s2_tr1": "d4afe9877761"},
}


def test_shipped_lut_hashes():
    for model, hs in LUT_HASHES.items():
        luts = orc.load_luts(lut_dir(model), linear=(model == "lerf-l"))
        for key, h in hs.items():
            k = key if key.startswith("s2") else key + "r0"
            assert sha12(luts[k]) == h, (model, key)


@pytest.mark.parametrize("iname", ["uniform", "ties"])
def test_lut_pass_all_modes_bit_exact(iname):
    g = golden("lut_pass")
    h, w = 9, 11
    img = g["img_" + iname]
    for mode in "sdyct":
        pad = MODE_PAD[mode]
        for oC in (1, 3):
            for rot in range(4):
                got = orc.FourSimplexInterpFaster(g["table_oc%d" % oC], img[:, :h + pad, :w + pad], h, w, 4, rot,
                                                  upscale=1, mode=mode, oC=oC)
                want = g["out_%s_%s_oc%d_rot%d" % (iname, mode, oC, rot)]
                assert got.shape == want.shape
                assert np.array_equal(got, want), (mode, oC, rot)


def test_lut_pass_errors():
    g = golden("lut_pass")
    with pytest.raises(ValueError):
        orc.FourSimplexInterpFaster(g["table_oc1"], g["img_uniform"], 9, 11, 4, 0, mode="x", oC=1)


@pytest.mark.parametrize("model", ["g", "l"])
def test_stages_bit_exact_and_survey_hashes(model):
    g = golden("lut_stages")
    luts = orc.load_luts(lut_dir("lerf-" + model), linear=(model == "l"))
    oC = 3 if model == "g" else 1
    names = [k[3:] for k in g.files if k.startswith("in_")]
    assert len(names) >= 10
    for n in names:
        feat, codes, hyper = orc.lut_stages(g["in_" + n], luts, oC=oC)
        assert np.array_equal(feat, g["feat_%s_%s" % (model, n)]), n
        assert np.array_equal(codes, g["codes_%s_%s" % (model, n)]), n
        if n in SET5:
            assert (sha12(feat), sha12(codes)) == SURVEY_HASHES[model][n], n


def _maxabs(a, b):
    m = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), m)
    return float(np.max(np.abs(a[m] - b[m]))) if m.any() else 0.0


def _count_bad(a, b, tol):
    m = np.isfinite(b) & np.isfinite(a)
    return int(np.sum(np.isfinite(a) != np.isfinite(b)) + np.sum(np.abs(a[m] - b[m]) > tol))


def test_resize_sr_matches_reference_float64():
    g = golden("resize_sr")
    img = g["img"].astype(np.float32)
    hyper = g["codes"].astype(np.float32) / float(255)
    for i, (sh, sw) in enumerate(g["scales"]):
        rs = orc.SteeringGaussianResize2dNumpy(support_sz=2, max_sigma=10)
        rs.set_shape(img.shape, scale_factors=[sh, sw])
        got = rs.resize(img, hyper[0::3], hyper[1::3], hyper[2::3])
        assert got.shape == g["gauss_%d" % i].shape
        assert _maxabs(got, g["gauss_%d" % i]) < 1e-10, (sh, sw)
        rl = orc.AmplifiedLinearResize2dNumpy()
        rl.set_shape(img.shape, scale_factors=[sh, sw])
        got = rl.resize(img, hyper[0:3])
        assert _maxabs(got, g["linear_%d" % i]) < 1e-10, (sh, sw)
    rs = orc.SteeringGaussianResize2dNumpy(support_sz=2, max_sigma=4)
    rs.set_shape(img.shape, scale_factors=[3, 3])
    assert _maxabs(rs.resize(img, hyper[0::3], hyper[1::3], hyper[2::3]), g["gauss_ms4"]) < 1e-10


def test_warp_matches_reference_float64():
    g = golden("warp")
    img = g["img"].astype(np.float32)
    hyper = g["codes"].astype(np.float32) / float(255)
    oshape = tuple(int(v) for v in g["out_shape"])
    for i, M in enumerate(g["mats"]):
        rs = orc.SteeringGaussianWarp2dNumpy(support_sz=2, max_sigma=10)
        rs.set_shape(img.shape, M, oshape)
        assert _maxabs(rs.warp(img, hyper[0::3], hyper[1::3], hyper[2::3]), g["gauss_%d" % i]) < 1e-9, i
        rl = orc.AmplifiedLinearWarp2dNumpy()
        rl.set_shape(img.shape, M, oshape)
        # The linear kernel is discontinuous at |d| = 1 (interp of resize_right2d_numpy.py:587-589): an output
        # pixel that maps EXACTLY onto an input grid point (matrix 0 maps (0,0) -> (2,3)) gets d = 1 +- 1e-16
        # depending on the BLAS kernel np.dot picked, so the reference itself is not reproducible there.
        assert _count_bad(rl.warp(img, hyper[0:3]), g["linear_%d" % i], 1e-9) <= 3, i
        nn = orc.NearestWarp2dNumpy()
        nn.set_shape(img.shape, M, oshape)
        assert _maxabs(nn.warp(img), g["nearest_%d" % i]) == 0.0, i
        assert np.array_equal(orc.warp_mask(img.shape, M, oshape), g["mask_%d" % i]), i


def test_fixed_kernel_warps_match_reference_float64():
    """Bilinear / Bicubic / Lanczos2 / Lanczos3Warp2dNumpy goldens (tests/golden/make_golden_fixed.py)."""
    g = golden("fixed_warp")
    w = golden("warp")
    img = w["img"].astype(np.float32)
    oshape = tuple(int(v) for v in g["out_shape"])
    for i in g["which"]:
        for name, cls in (("bilinear", orc.BilinearWarp2dNumpy), ("bicubic", orc.BicubicWarp2dNumpy),
                          ("lanczos2", orc.Lanczos2Warp2dNumpy), ("lanczos3", orc.Lanczos3Warp2dNumpy)):
            rs = cls()
            rs.set_shape(img.shape, w["mats"][i], oshape)
            assert _maxabs(rs.warp(img), g["%s_%d" % (name, i)]) < 1e-9, (name, i)


def test_whole_path_set5():
    g = golden("set5_path")
    st = golden("lut_stages")
    lg = orc.load_luts(lut_dir("lerf-g"), linear=False)
    ll = orc.load_luts(lut_dir("lerf-l"), linear=True)
    for tag, n, luts, linear, s in (("g_butterfly_x4", "butterfly", lg, False, 4), ("g_bird_x2", "bird", lg, False, 2),
                                    ("l_woman_x3p5", "woman", ll, True, 3.5)):
        out, _, _ = orc.lerf_sr(st["in_" + n], luts, s, s, linear=linear)
        want = g["sr_" + tag]
        assert _maxabs(out, want) < 1e-9, tag
        assert np.array_equal(orc.to_uint8_hwc(out), orc.to_uint8_hwc(want)), tag
    for tag, luts, linear in (("g_isc_butterfly", lg, False), ("g_osc_butterfly", lg, False), ("l_osc_woman", ll, True)):
        out, mask, _, _ = orc.lerf_warp(g["warp_in_" + tag], luts, g["warp_M_" + tag],
                                        tuple(int(v) for v in g["warp_gt_shape_" + tag]), linear=linear)
        assert np.array_equal(mask, g["warp_mask_" + tag]), tag
        assert _maxabs(out, g["warp_out_" + tag]) < 1e-8, tag


def test_resize_non_default_parameters_match_reference_float64():
    """Support sizes 1 / 3 / 4 / 6, np.pad modes of the image and the antialias branch of a height factor below 1
    (resize_right2d_numpy.py:51-55, :186-193, :208): the oracle against goldens generated by the reference
    (tests/golden/make_golden_general.py)."""
    G = golden("resize_general")
    img, hy = G["img"], [G["h0"], G["h1"], G["h2"]]
    for i, case in enumerate(G["cases"]):
        supp, sh, sw, pm = str(case).split("|")
        g = orc.SteeringGaussianResize2dNumpy(support_sz=int(supp), max_sigma=10, pad_mode=pm)
        g.set_shape(list(img.shape), scale_factors=[float(sh), float(sw)])
        assert g.support_sz == int(G["supp_after_%d" % i]), case
        ref = G["gauss_%d" % i]
        assert np.max(np.abs(g.resize(img, *hy) - ref)) <= 1e-9, case
        lin = orc.AmplifiedLinearResize2dNumpy(support_sz=int(supp), pad_mode=pm)
        lin.set_shape(list(img.shape), scale_factors=[float(sh), float(sw)])
        got, ref = lin.resize(img, hy[0]), G["linear_%d" % i]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), case
        m = np.isfinite(ref)
        assert np.max(np.abs(got[m] - ref[m])) <= 1e-9, case


def _warp_close(got, ref, tol, jumps=0):
    """Equal NaN patterns and values within tol, except at most ``jumps`` pixels: the linear kernel 1 - |d|/alpha... is cut
    at |d| = 1 (resize_right2d_numpy.py:590-600), and a tap whose projected distance is 1 to the last ulp falls on either
    side of the cut depending on the BLAS that multiplied the homography."""
    assert got.shape == ref.shape
    bad = np.isnan(got) != np.isnan(ref)
    m = ~np.isnan(got) & ~np.isnan(ref)
    bad |= m & (np.abs(np.where(m, got, 0) - np.where(m, ref, 0)) > tol)
    assert int(bad.any(axis=0).sum()) <= jumps, int(bad.any(axis=0).sum())


def test_warp_non_default_parameters_match_reference_float64():
    """support_sz 1 / 3 / 4 / 6 and np.pad modes for the WARP classes (resize_right2d_numpy.py:363-369, :397-398, :559):
    the oracle against goldens generated by the reference (tests/golden/make_golden_general.py)."""
    G = golden("resize_general")
    img, hy = G["img"], [G["h0"], G["h1"], G["h2"]]
    oshape = [3] + [int(v) for v in G["warp_out_hw"]]
    for i, case in enumerate(G["warp_cases"]):
        supp, pm, mi = str(case).split("|")
        g = orc.SteeringGaussianWarp2dNumpy(support_sz=int(supp), max_sigma=10, pad_mode=pm)
        g.set_shape(list(img.shape), G["warp_M"][int(mi)], oshape)
        _warp_close(g.warp(img, *hy), G["warp_gauss_%d" % i], 1e-9)
        lin = orc.AmplifiedLinearWarp2dNumpy(support_sz=int(supp), pad_mode=pm)
        lin.set_shape(list(img.shape), G["warp_M"][int(mi)], oshape)
        _warp_close(lin.warp(img, hy[0]), G["warp_linear_%d" % i], 1e-9, jumps=3)
