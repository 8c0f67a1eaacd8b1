"""CPU: the metrics module (lerf_pytorch_b200/metrics.py) against values the reference's own common/utils.py
functions produced (stored by tests/golden/make_golden.py), and the adapters' option parsing / table format."""
import numpy as np

from util import golden


def test_psnr_ssim_match_reference_values():
    from lerf_pytorch_b200 import metrics
    g = golden("set5_path")
    want = g["sr_g_butterfly_x4"]
    out_u8 = np.clip(np.round(want).transpose((1, 2, 0)), 0, 255).astype(np.uint8)
    psnr, ssim = metrics.psnr_y_ssim(g["hr_butterfly"], out_u8, 4, 4)
    ref_psnr, ref_ssim = g["set5_x4_psnr_ssim_g"][2]  # butterfly is the third Set5 image
    assert abs(psnr - ref_psnr) < 1e-4 and abs(ssim - ref_ssim) < 1e-9


def test_mpsnr_matches_reference_value():
    from lerf_pytorch_b200 import metrics
    g = golden("set5_path")
    for tag in ("g_isc_butterfly", "g_osc_butterfly"):
        out = np.clip(np.round(g["warp_out_" + tag]).transpose((1, 2, 0)), 0, 255).astype(np.uint8)  # NaN -> 0 like astype
        mask = g["warp_mask_" + tag].transpose((1, 2, 0))
        got = metrics.mpsnr(out, g["hr_butterfly"], mask)
        assert abs(got - float(g["warp_mpsnr_" + tag])) < 1e-3, tag


def test_adapter_options_and_table_format():
    from lerf_pytorch_b200 import eval_lut_sr, eval_lut_warp
    from lerf_pytorch_b200.eval_common import build_parser, check_supported
    opt = build_parser("x", "./data/rrBenchmark").parse_args(["-e", "models/lerf-g"])
    assert (opt.modes, opt.modes2, opt.interval, opt.suppSize, opt.maxSigma, opt.stages, opt.lutName) == \
        ("sct", "sct", 4, 2, 10, 2, "LUTft")  # common/option.py defaults
    assert opt.testDir == "./data/rrBenchmark" and opt.resultRoot == "./results" and not opt.linear
    check_supported(opt)
    res = {("Set5", (2.0, 2.0)): [[35.0, 0.9], [36.0, 0.95]], ("Set5", (4.0, 4.0)): [[30.125, 0.85]]}
    lines = eval_lut_sr.format_table(["Set5"], [(2.0, 2.0), (4.0, 4.0)], res)
    assert lines[0] == "Scale          \t2.0x2.0\t\t4.0x4.0\t"
    assert lines[1] == "Set5           \t35.50/0.9250\t30.12/0.8500"
    wl = eval_lut_warp.format_table(["Set5"], ["isc", "osc"], {("Set5", "isc"): [[33.806]], ("Set5", "osc"): [[27.894]]})
    assert wl == ["Scale          \tisc\t\tosc\t", "Set5           \t33.81\t27.89"]


def test_gpu_metrics_module_equals_the_host_restatement_on_cpu_tensors():
    """metrics_gpu.py (torch ops; runs on any device) against metrics.py on the same inputs."""
    import torch
    from lerf_pytorch_b200 import metrics, metrics_gpu
    rng = np.random.default_rng(3)
    gt = rng.integers(0, 256, (75, 62, 3)).astype(np.uint8)
    out = np.clip(gt.astype(int) + rng.integers(-12, 13, gt.shape), 0, 255).astype(np.uint8)
    for sc in (2, 4):
        a = metrics.psnr_y_ssim(gt, out[:-3, :-1], sc, sc)
        b = metrics_gpu.psnr_y_ssim(torch.from_numpy(gt), torch.from_numpy(np.ascontiguousarray(out[:-3, :-1])), sc, sc)
        assert abs(float(a[0]) - b[0]) < 1e-4 and abs(float(a[1]) - b[1]) < 1e-9
    m = rng.random(gt.shape) < 0.6
    assert abs(metrics.mpsnr(out, gt, m) - metrics_gpu.mpsnr(torch.from_numpy(out), torch.from_numpy(gt), torch.from_numpy(m))) < 1e-4
