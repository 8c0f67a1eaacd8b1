"""GPU: the eval-script adapters reproduce the known answers the reference publishes in scripts.sh:33-47 on the
shipped Set5 fixtures (copied as data under tests/golden/data by make_golden.py's rules)."""
import os

import pytest

from util import GOLDEN, lut_dir

pytestmark = pytest.mark.gpu
DATA = os.path.join(GOLDEN, "data")

# scripts.sh:33-38 (SR, PSNR-Y/SSIM at x2, x3, x4) and :42-47 (warp, mPSNR isc / osc)
SR_PINS = {"lerf-g": ["35.71/0.9475", "32.02/0.8980", "30.15/0.8548"], "lerf-l": ["34.84/0.9432", "30.72/0.8773", "29.13/0.8270"]}
WARP_PINS = {"lerf-g": ["33.81", "27.89"], "lerf-l": ["32.90", "27.13"]}


@pytest.mark.parametrize("model", ["lerf-g", "lerf-l"])
def test_eval_lut_sr_adapter_reproduces_published_table(model, tmp_path, capsys):
    from lerf_pytorch_b200 import eval_lut_sr
    argv = ["-e", lut_dir(model), "--testDir", os.path.join(DATA, "rrBenchmark"), "--resultRoot", str(tmp_path)]
    if model == "lerf-l":
        argv.append("--linear")
    lines, _ = eval_lut_sr.main(argv)
    assert lines[0].split("\t")[0] == "Scale".ljust(15, " ")
    assert lines[1].split("\t") == ["Set5".ljust(15, " ")] + SR_PINS[model]
    assert capsys.readouterr().out.strip().splitlines()[-1] == lines[1]
    out_dir = os.path.join(str(tmp_path), model, "X4.00_4.00", "Set5")
    names = sorted(os.listdir(out_dir))
    assert "baby_LUTft.png" in names and "baby_lr.png" in names and "baby_gt.png" in names and "baby_LUTft_hyper.npy" in names


@pytest.mark.parametrize("model", ["lerf-g", "lerf-l"])
def test_eval_lut_warp_adapter_reproduces_published_table(model, tmp_path):
    from lerf_pytorch_b200 import eval_lut_warp
    argv = ["-e", lut_dir(model), "--testDir", os.path.join(DATA, "WarpBenchmark"), "--resultRoot", str(tmp_path), "--no-save"]
    if model == "lerf-l":
        argv.append("--linear")
    lines, _ = eval_lut_warp.main(argv)
    assert lines[1].split("\t") == ["Set5".ljust(15, " ")] + WARP_PINS[model]


def test_adapters_with_gpu_metrics_print_the_same_tables(tmp_path):
    """--gpu-metrics: PSNR-Y / SSIM / mPSNR computed on the device (metrics_gpu.py); the printed tables do not change."""
    from lerf_pytorch_b200 import eval_lut_sr, eval_lut_warp
    lines, _ = eval_lut_sr.main(["-e", lut_dir("lerf-g"), "--testDir", os.path.join(DATA, "rrBenchmark"), "--resultRoot",
                                 str(tmp_path), "--no-save", "--gpu-metrics"])
    assert lines[1].split("\t") == ["Set5".ljust(15, " ")] + SR_PINS["lerf-g"]
    lines, _ = eval_lut_warp.main(["-e", lut_dir("lerf-g"), "--testDir", os.path.join(DATA, "WarpBenchmark"), "--resultRoot",
                                   str(tmp_path), "--no-save", "--gpu-metrics"])
    assert lines[1].split("\t") == ["Set5".ljust(15, " ")] + WARP_PINS["lerf-g"]


def test_sr_adapter_with_gpu_png_writes_the_same_images(tmp_path):
    """--gpu-png: the result PNGs are assembled on the device (png_gpu.py); they decode to the pixels PIL's files hold and
    the table does not change."""
    import numpy as np
    from PIL import Image
    from lerf_pytorch_b200 import eval_lut_sr
    base = ["-e", lut_dir("lerf-g"), "--testDir", os.path.join(DATA, "rrBenchmark")]
    a, b = os.path.join(str(tmp_path), "pil"), os.path.join(str(tmp_path), "gpu")
    lines_a, _ = eval_lut_sr.main(base + ["--resultRoot", a])
    lines_b, _ = eval_lut_sr.main(base + ["--resultRoot", b, "--gpu-png"])
    assert lines_a == lines_b
    sub = os.path.join("lerf-g", "X4.00_4.00", "Set5")
    for name in sorted(os.listdir(os.path.join(a, sub))):
        if name.endswith("_LUTft.png"):
            assert np.array_equal(np.asarray(Image.open(os.path.join(a, sub, name))), np.asarray(Image.open(os.path.join(b, sub, name)))), name


def test_warp_adapter_with_gpu_png_writes_the_same_images(tmp_path):
    """--gpu-png in the warp adapter: masking to white and the PNG file on the device; same pixels as PIL's files."""
    import numpy as np
    from PIL import Image
    from lerf_pytorch_b200 import eval_lut_warp
    base = ["-e", lut_dir("lerf-g"), "--testDir", os.path.join(DATA, "WarpBenchmark")]
    a, b = os.path.join(str(tmp_path), "pil"), os.path.join(str(tmp_path), "gpu")
    lines_a, _ = eval_lut_warp.main(base + ["--resultRoot", a])
    lines_b, _ = eval_lut_warp.main(base + ["--resultRoot", b, "--gpu-png"])
    assert lines_a == lines_b
    n = 0
    for root, _, files in os.walk(a):
        for name in files:
            if name.endswith("_LUTft.png"):
                other = os.path.join(b, os.path.relpath(root, a), name)
                assert np.array_equal(np.asarray(Image.open(os.path.join(root, name))), np.asarray(Image.open(other))), name
                n += 1
    assert n >= 2
