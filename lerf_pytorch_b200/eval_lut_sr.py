#!/usr/bin/env python
"""Drop-in for ``python resample/eval_lut_sr.py -e <expDir> [--linear]`` of the reference, on the B200 path.

Same options (common/option.py), same directory layout (``<testDir>/<dataset>/HR/*.png`` and
``LR_bicubic/rrLR_X{sh:.2f}_{sw:.2f}/*.png``), same result files and the same printed table as
resample/eval_lut_sr.py:747-811; the per-image body (eltr._worker, :514-744) runs as three CUDA kernels through
``LerfSR``.  The reference hard-codes ``all_datasets = ["Set5"]`` and scales 2, 3, 4 (:777-791); here they are
``--datasets`` and ``--scales`` with those defaults.

    python -m lerf_pytorch_b200.eval_lut_sr -e models/lerf-g
    python -m lerf_pytorch_b200.eval_lut_sr -e models/lerf-l --linear
"""
import os
import sys

import numpy as np

from . import metrics
from .eval_common import IoPipeline, build_parser, check_supported, list_pngs, load_lut_dict_like_reference, load_rgb


class Evaluator(object):
    """``eltr`` of the reference without its module globals (eval_lut_sr.py:473-512)."""

    def __init__(self, opt, lut_dict):
        import torch
        from . import LerfSR, LutSet
        self.opt = opt
        self.torch = torch
        self.device = torch.device(opt.device)
        self.luts = LutSet(lut_dict, linear=opt.linear, device=self.device)
        self._sr = {}
        self._LerfSR = LerfSR

    def engine(self, scale_h, scale_w):
        key = (float(scale_h), float(scale_w))
        if key not in self._sr:
            self._sr[key] = self._LerfSR(self.luts, scale_h, scale_w, max_sigma=self.opt.maxSigma,
                                         support_sz=self.opt.suppSize)
        return self._sr[key]

    def run(self, dataset, scale_h, scale_w):
        """One dataset at one scale.  Decode (next images), GPU work (this image) and encode + metrics (previous images)
        overlap through an IoPipeline; results come back in file order."""
        opt = self.opt
        files = list_pngs(os.path.join(opt.testDir, dataset, "HR"))
        result_path = os.path.join(opt.resultRoot, opt.expDir.rstrip("/").split("/")[-1],
                                   "X{:.2f}_{:.2f}".format(scale_h, scale_w), dataset)
        if opt.save and not os.path.isdir(result_path):
            os.makedirs(result_path)
        io = IoPipeline(getattr(opt, "io_threads", 0))
        try:
            def load(fname):
                lr = load_rgb(os.path.join(opt.testDir, dataset, "LR_bicubic/rrLR_X{:.2f}_{:.2f}".format(scale_h, scale_w), fname))
                return fname, lr, load_rgb(os.path.join(opt.testDir, dataset, "HR", fname))

            scores = [self._worker(io, fname, lr, gt, scale_h, scale_w, result_path) for fname, lr, gt in io.prefetch(load, files)]
            io.drain()
            return [f.result() for f in scores]
        finally:
            io.close()

    def _worker(self, io, fname, img_lr, img_gt, scale_h, scale_w, result_path):
        opt, torch = self.opt, self.torch
        # eval_lut_sr.py:630-641: result paths of pre-upscaled inputs (RRDB x4, LUT x2, down2 / down4) resample by scale / post
        post = 1
        if "rrdb" in result_path:
            post = 4
        elif "lutx2" in result_path:
            post = 2
        elif "down2" in result_path:
            post = 2
        elif "down4" in result_path:
            post = 4
        sr = self.engine(scale_h / post, scale_w / post)
        gpu_score = None
        with torch.cuda.device(self.device):
            d_in = torch.from_numpy(np.ascontiguousarray(img_lr.astype(np.uint8))).to(self.device)
            d_out = sr(d_in, out_format="u8_hwc")                            # :541-665 incl. the uint8 epilogue
            if getattr(opt, "gpu_metrics", False):
                from . import metrics_gpu
                gpu_score = metrics_gpu.psnr_y_ssim(torch.from_numpy(np.ascontiguousarray(img_gt)).to(self.device), d_out,
                                                    scale_h, scale_w)
            img_out = d_out.cpu().numpy() if (opt.save or gpu_score is None) else None
            png_file = None
            if opt.save and getattr(opt, "gpu_png", False):  # the PNG FILE of the result is assembled on the device
                from . import png_gpu
                png_file = png_gpu.encode_png(d_out).cpu().numpy()
            if opt.save:
                feat, codes = sr.stages(d_in)
                feat = feat.cpu().numpy()
                img_hyper = codes.cpu().numpy().astype(np.float32) / float(255)  # :623-628
        if opt.save:
            stem = fname.split("/")[-1][:-4]
            if png_file is not None:
                io.submit(_write_bytes, png_file, os.path.join(result_path, "{}_{}.png".format(stem, opt.lutName)))
            else:
                io.submit(_save_png, img_out, os.path.join(result_path, "{}_{}.png".format(stem, opt.lutName)))
            io.submit(_save_png, np.ascontiguousarray(feat.transpose((1, 2, 0))), os.path.join(result_path, "{}_lr.png".format(stem)))
            io.submit(_save_png, img_gt, os.path.join(result_path, "{}_gt.png".format(stem)))
            io.submit(np.save, os.path.join(result_path, "{}_{}_hyper.npy".format(fname.split("_")[-1][:-4], opt.lutName)), img_hyper)
        if gpu_score is not None:
            return io.submit(lambda: gpu_score)
        return io.submit(metrics.psnr_y_ssim, img_gt, img_out, scale_h, scale_w)


def _write_bytes(arr, path):
    with open(path, "wb") as f:
        f.write(arr.tobytes())


def _save_png(arr, path):
    from PIL import Image
    Image.fromarray(arr).save(path)


def format_table(all_datasets, all_scales, results):
    """The table of eval_lut_sr.py:793-811; results[(dataset, (sh, sw))] = list of [psnr, ssim]."""
    lines = []
    head = ["Scale".ljust(15, " ")]
    for sh, sw in all_scales:
        head.append("{:.1f}x{:.1f}\t".format(sh, sw))
    lines.append("\t".join(head))
    for dataset in all_datasets:
        row = [dataset.ljust(15, " ")]
        for sc in all_scales:
            ps = np.asarray(results[(dataset, tuple(sc))])
            row.append("{:.2f}/{:.4f}".format(np.mean(ps[:, 0]), np.mean(ps[:, 1])))
        lines.append("\t".join(row))
    return lines


def main(argv=None):
    p = build_parser(__doc__.splitlines()[0], './data/rrBenchmark')
    p.add_argument('--scales', type=str, default='2x2,3x3,4x4', help='comma-separated HxW scale pairs, e.g. 2x2,2x3,3.5x3.5')
    opt = p.parse_args(argv)
    check_supported(opt)
    lut_dict = load_lut_dict_like_reference(opt)
    ev = Evaluator(opt, lut_dict)
    all_datasets = [d for d in opt.datasets.split(",") if d]
    all_scales = [tuple(float(v) for v in s.lower().split("x")) for s in opt.scales.split(",") if s]
    results = {}
    for dataset in all_datasets:
        for sh, sw in all_scales:
            results[(dataset, (sh, sw))] = ev.run(dataset, sh, sw)
    lines = format_table(all_datasets, all_scales, results)
    print("\n".join(lines))
    return lines, results


if __name__ == "__main__":
    main()
    sys.exit(0)
