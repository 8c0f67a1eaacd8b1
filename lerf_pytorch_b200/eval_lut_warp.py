#!/usr/bin/env python
"""Drop-in for ``python resample/eval_lut_warp.py -e <expDir> [--linear] --testDir data/WarpBenchmark`` on the B200 path.

Same options, same layout (``<testDir>/<dataset>/{isc,osc}/<name>.{png,pth}`` + ``HR/<name>.png``), same result
files and the same printed table as resample/eval_lut_warp.py:305-355; the per-image body (:72-233) -- LUT stages,
homographic warp, nearest-neighbour validity mask -- runs in CUDA through ``LerfWarp``; the masked PSNR is
common/utils.py:168-175 restated in metrics.py.
"""
import os
import sys

import numpy as np

from . import metrics
from .eval_common import IoPipeline, build_parser, check_supported, list_pngs, load_lut_dict_like_reference, load_rgb


class Evaluator(object):
    """``eltr`` of eval_lut_warp.py:27-70 without its module globals."""

    def __init__(self, opt, lut_dict):
        import torch
        from . import LerfWarp, LutSet
        self.opt = opt
        self.torch = torch
        self.device = torch.device(opt.device)
        self.luts = LutSet(lut_dict, linear=opt.linear, device=self.device)
        self.border = 4  # eval_lut_warp.py:36
        self.warp = LerfWarp(self.luts, max_sigma=opt.maxSigma, support_sz=opt.suppSize, border=self.border)

    def run(self, dataset, scale_p):
        """One dataset at one scale class; decode / GPU work / encode + metrics overlap through an IoPipeline."""
        opt, torch = self.opt, self.torch
        files = list_pngs(os.path.join(opt.testDir, dataset, "HR"))
        result_path = os.path.join(opt.resultRoot, opt.expDir.rstrip("/").split("/")[-1], dataset, scale_p)
        if opt.save and not os.path.isdir(result_path):
            os.makedirs(result_path)
        io = IoPipeline(getattr(opt, "io_threads", 0))
        try:
            def load(fname):
                lr = load_rgb(os.path.join(opt.testDir, dataset, scale_p, fname))
                matrix = torch.load(os.path.join(opt.testDir, dataset, scale_p, fname.replace("png", "pth"))).numpy()
                return fname, lr, matrix, load_rgb(os.path.join(opt.testDir, dataset, "HR", fname))

            scores = [self._worker(io, fname, lr, m, gt, result_path) for fname, lr, m, gt in io.prefetch(load, files)]
            io.drain()
            return [f.result() for f in scores]
        finally:
            io.close()

    def _worker(self, io, fname, img_lr, matrix, img_gt, result_path):
        opt, torch = self.opt, self.torch
        with torch.cuda.device(self.device):
            d_in = torch.from_numpy(np.ascontiguousarray(img_lr.astype(np.uint8))).to(self.device)
            out, mask = self.warp(d_in, matrix, img_gt.shape[:2], out_format="u8_hwc", with_mask=True)
            if getattr(opt, "gpu_metrics", False) and not opt.save:  # mPSNR on the device: nothing is copied back
                from . import metrics_gpu
                d_gt = torch.from_numpy(np.ascontiguousarray(img_gt)).to(self.device)
                valid3 = mask.to(torch.bool).unsqueeze(-1).expand_as(d_gt)
                score = [metrics_gpu.mpsnr(out, d_gt, valid3)]
                return io.submit(lambda: score)
            img_out = out.cpu().numpy()
            valid = mask.cpu().numpy().astype(bool)                        # mask_output == 255 (:229)
            feat = None
            png_file = None
            if opt.save:
                from .lut_interp import lut_stage1
                feat = lut_stage1(self.luts, d_in, "HWC").cpu().numpy()
                if getattr(opt, "gpu_png", False):  # non valid pixels white (:259-261), then the PNG file, both on the device
                    from . import png_gpu
                    shown = torch.where(mask.to(torch.bool).unsqueeze(-1), out, torch.full_like(out, 255))
                    png_file = png_gpu.encode_png(shown).cpu().numpy()
        return io.submit(_score_and_save, img_out, img_gt, valid, feat, fname, result_path, opt.lutName, opt.save, png_file)


def _score_and_save(img_out, img_gt, valid, feat, fname, result_path, lut_name, save, png_file=None):
    """Host side of eltr._worker after the warp (eval_lut_warp.py:226-261): mPSNR inside the mask, then the four PNGs."""
    from PIL import Image
    valid3 = np.repeat(valid[:, :, None], img_gt.shape[2], axis=2)
    mpsnr = metrics.mpsnr(img_out, img_gt, valid3)                          # :226-233
    if save:
        stem = fname.split("/")[-1][:-4]
        Image.fromarray(np.ascontiguousarray(feat.transpose((1, 2, 0)))).save(os.path.join(result_path, "{}_lr.png".format(stem)))
        Image.fromarray((valid3 * 255).astype(np.uint8)).save(os.path.join(result_path, "{}_mask.png".format(stem)))
        white = (np.ones_like(img_gt) * 255).astype(np.uint8)              # non valid pixels leave as white (:259-261)
        if png_file is not None:
            with open(os.path.join(result_path, "{}_{}.png".format(stem, lut_name)), "wb") as f:
                f.write(png_file.tobytes())
        else:
            Image.fromarray(img_out * valid3 + (~valid3) * white).save(os.path.join(result_path, "{}_{}.png".format(stem, lut_name)))
        Image.fromarray(img_gt * valid3 + (~valid3) * white).save(os.path.join(result_path, "{}_gt.png".format(stem)))
    return [mpsnr]


def format_table(all_datasets, all_scales, results):
    """The table of eval_lut_warp.py:341-355."""
    lines = []
    head = ["Scale".ljust(15, " ")]
    for sp in all_scales:
        head.append("{}\t".format(sp))
    lines.append("\t".join(head))
    for dataset in all_datasets:
        row = [dataset.ljust(15, " ")]
        for sp in all_scales:
            row.append("{:.2f}".format(np.mean(np.asarray(results[(dataset, sp)])[:, 0])))
        lines.append("\t".join(row))
    return lines


def main(argv=None):
    p = build_parser(__doc__.splitlines()[0], './data/WarpBenchmark')
    p.add_argument('--scales', type=str, default='isc,osc', help='comma-separated sub-folders (in-scale / out-of-scale)')
    opt = p.parse_args(argv)
    check_supported(opt)
    lut_dict = load_lut_dict_like_reference(opt)
    ev = Evaluator(opt, lut_dict)
    all_datasets = [d for d in opt.datasets.split(",") if d]
    all_scales = [s for s in opt.scales.split(",") if s]
    results = {}
    for dataset in all_datasets:
        for sp in all_scales:
            results[(dataset, sp)] = ev.run(dataset, sp)
    lines = format_table(all_datasets, all_scales, results)
    print("\n".join(lines))
    return lines, results


if __name__ == "__main__":
    main()
    sys.exit(0)
