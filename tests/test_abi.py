"""CPU: the C-ABI library builds, loads, and exports every symbol include/lerf_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import lerf_pytorch_b200 as lp
    return lp


def _declared_symbols(header="lerf_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lerf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from lerf_pytorch_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 18
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "liblerf_b200.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "python binding lacks a prototype for %s" % n
    assert sorted(_lib.PROTOTYPES) == names
    assert not [n for n in names if "debug" in n], "test hooks belong in lerf_b200_testing.h"
    tnames = _declared_symbols("lerf_b200_testing.h")
    for n in tnames:
        assert hasattr(L, n), "liblerf_b200.so does not export %s" % n
    assert sorted(_lib.TESTING_PROTOTYPES) == tnames
    assert L.lerf_build_has_experiments() == 0, "the product library must be built without -DLERF_EXPERIMENTS"


def test_abi_version_and_error_string(built):
    L = built.lib()
    assert L.lerf_abi_version() == 1
    # argument validation happens before any CUDA call, so these run without a GPU
    assert L.lerf_lut_pass(None, None, 1, 4, 4, b"x", 1, None, None) == 1
    assert b"Mode x not implemented" in L.lerf_last_error_string()
    assert L.lerf_sr_scratch_bytes(3, 3, 10, 10) >= 3 * 100 * 4
    assert L.lerf_sr_scratch_bytes(3, 3, 0, 10) == 0


def test_no_cpu_fallback_when_library_missing(built, monkeypatch):
    from lerf_pytorch_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(ImportError):
        _lib.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lerf_pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "lerf_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_host_geometry_tables_match_reference_formulas(built):
    """sr_axis_tables against a direct transcription of the closed forms of SURVEY.md A.6."""
    from lerf_pytorch_b200 import sr_axis_tables
    eps = np.float64(1.1920928955078125e-07)
    for in_sz, scale in ((13, 2), (17, 3), (64, 4), (7, 3.5), (9, 1.5), (30, 8), (5, 1), (11, 2.4)):
        out_sz = int(np.ceil(scale * in_sz))
        left, dist, pad = sr_axis_tables(in_sz, out_sz, scale)
        o = np.arange(out_sz)
        p = o / float(scale) + (in_sz - 1) / 2 - (out_sz - 1) / (2 * float(scale))
        want_left = np.ceil(p - 1 - eps).astype(np.int64)
        assert np.array_equal(left, want_left)
        assert pad[0] == -want_left[0] and pad[1] == want_left[-1] + 1 - in_sz + 1
        assert np.allclose(dist[:, 0], p - want_left, atol=1e-12) and np.allclose(dist[:, 1], p - want_left - 1, atol=1e-12)
        assert left.min() >= -1 and left.max() <= in_sz - 1


def test_sr_set_shape_mirrors_reference_attributes(built):
    r = built.SteeringGaussianResize2dNumpy(support_sz=2, max_sigma=10)
    r.set_shape([3, 20, 30], scale_factors=[3.5, 2])
    assert r.out_shape == [3, 70, 60] and r.out_sz == [70, 60] and r.scale_factors == [1.0, 3.5, 2.0]
    assert r.pad_vec == ((0, 0), (1, 1), (1, 1))
    r.set_shape([3, 20, 30], out_shape=[3, 40, 90])
    assert r.scale_factors == [1.0, 2.0, 3.0]
    # a height factor below 1 latches antialiasing on and grows support_sz for good (resize_right2d_numpy.py:51-55)
    r.set_shape([3, 20, 30], scale_factors=[0.5, 0.5])
    assert r.antialias and r.support_sz == 4 and r.pad_vec == ((0, 0), (1, 1), (1, 1))
    r.set_shape([3, 20, 30], scale_factors=[0.5, 0.5])
    assert r.support_sz == 8  # it compounds, like in the reference
    r4 = built.SteeringGaussianResize2dNumpy(support_sz=4, pad_mode="reflect")
    r4.set_shape([3, 8, 8], scale_factors=[2, 2])
    assert r4.pad_vec == ((0, 0), (2, 2), (2, 2)) and not r4.antialias
    with pytest.raises(NotImplementedError):
        built.SteeringGaussianResize2dNumpy(pad_mode="linear_ramp")
