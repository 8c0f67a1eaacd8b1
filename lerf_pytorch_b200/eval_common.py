"""Shared pieces of the two eval-script adapters: options and LUT loading with the reference's names.

Options mirror common/option.py:13-41 (BaseOptions) and :210-218 (TestOptions) of the reference: same flags, same
defaults; flags that only matter for training / network evaluation are accepted and ignored so the command lines of
scripts.sh:33-47 work unchanged.
"""
import argparse
import os

import numpy as np


def build_parser(description, default_test_dir):
    p = argparse.ArgumentParser(description=description)
    # BaseOptions (common/option.py:13-41)
    p.add_argument('--name', type=str, default='lerf')
    p.add_argument('--model', type=str, default='SRNetsSWF2')
    p.add_argument('--scale', '-r', type=str, default='4')
    p.add_argument('--nsigma', type=int, default=-1)
    p.add_argument('--nf', type=int, default=64)
    p.add_argument('--modes', type=str, default='sct')
    p.add_argument('--modes2', type=str, default='sct')
    p.add_argument('--interval', type=int, default=4, help='N bit uniform sampling')
    p.add_argument('--norm', type=int, default=255)
    p.add_argument('--suppSize', type=int, default=2)
    p.add_argument('--inC', type=int, default=1)
    p.add_argument('--outC', type=int, default=3)
    p.add_argument('--featC', type=int, default=1)
    p.add_argument('--maxSigma', type=int, default=10)
    p.add_argument('--stages', type=int, default=2)
    p.add_argument('--twoStage', action='store_true', default=False)
    p.add_argument('--linear', action='store_true', default=False, help='linear resampling function (LeRF-L)')
    p.add_argument('--modelRoot', type=str, default='./models')
    p.add_argument('--expDir', '-e', type=str, default='')
    p.add_argument('--load_from_opt_file', action='store_true', default=False)
    p.add_argument('--debug', default=False, action='store_true')
    # TestOptions (common/option.py:210-218)
    p.add_argument('--testDir', type=str, default=default_test_dir)
    p.add_argument('--resultRoot', type=str, default='./results')
    p.add_argument('--loadIter', type=int, default=50000)
    p.add_argument('--lutName', type=str, default='LUTft')
    # additions of this adapter
    p.add_argument('--datasets', type=str, default='Set5', help='comma-separated dataset folders under testDir')
    p.add_argument('--device', type=str, default='cuda:0')
    p.add_argument('--no-save', dest='save', action='store_false', default=True,
                   help='skip writing PNG / npy results (metrics only)')
    p.add_argument('--gpu-metrics', dest='gpu_metrics', action='store_true', default=False,
                   help='compute PSNR-Y / SSIM / mPSNR on the device (metrics_gpu.py); with --no-save the result image never '
                        'leaves the GPU')
    p.add_argument('--gpu-png', dest='gpu_png', action='store_true', default=False,
                   help='write the result image with the device PNG writer (png_gpu.py: stored deflate blocks, lossless, '
                        'uncompressed) instead of PIL on the host')
    p.add_argument('--io-threads', dest='io_threads', type=int, default=4,
                   help='host threads that decode the next images and encode / score finished ones while the GPU works '
                        '(SURVEY.md 8f item 2); 0 = everything inline like the reference')
    return p


class IoPipeline(object):
    """The image I/O step either side of the hot path (reference: PIL decode eval_lut_sr.py:517-534,
    PNG / npy writes :667-708, metrics :735-742, all inline on one thread).

    ``prefetch(load, items)`` yields ``load(item)`` in order while up to ``depth`` later items are being decoded on worker
    threads; ``submit(fn, *args)`` runs encodes / metrics on the same workers and returns a future; ``drain()`` waits
    for everything submitted and re-raises the first failure.  PIL's zlib work releases the GIL, so the workers run
    in parallel with each other and with the CUDA launches of the main thread.  With ``threads=0`` every call runs
    inline, which is the reference's behaviour.
    """

    def __init__(self, threads=4, depth=None):
        self.threads = max(0, int(threads))
        self.depth = max(1, self.threads if depth is None else depth)
        self._pool = None
        self._pending = []
        if self.threads:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.threads, thread_name_prefix="lerf-io")

    def prefetch(self, load, items):
        items = list(items)
        if not self._pool:
            for it in items:
                yield load(it)
            return
        futs = []
        nxt = 0
        for i in range(len(items)):
            while nxt < len(items) and nxt <= i + self.depth:
                futs.append(self._pool.submit(load, items[nxt]))
                nxt += 1
            yield futs[i].result()
            futs[i] = None  # drop the decoded image as soon as it is consumed

    def submit(self, fn, *args):
        if not self._pool:
            return _Done(fn(*args))
        f = self._pool.submit(fn, *args)
        self._pending.append(f)
        return f

    def drain(self):
        pending, self._pending = self._pending, []
        for f in pending:
            f.result()

    def close(self):
        try:
            self.drain()
        finally:
            if self._pool:
                self._pool.shutdown(wait=True)
                self._pool = None


class _Done(object):
    def __init__(self, value):
        self._value = value

    def result(self):
        return self._value


def check_supported(opt):
    if opt.modes != 'sct' or opt.modes2 != 'sct' or opt.interval != 4 or opt.stages != 2:
        raise NotImplementedError("the B200 path implements the shipped configuration: --modes sct --modes2 sct "
                                  "--interval 4 --stages 2 (other modes go through FourSimplexInterpFaster)")
    if opt.suppSize != 2:
        raise NotImplementedError("--suppSize %d: only support size 2 is implemented" % opt.suppSize)


def load_lut_dict_like_reference(opt):
    """resample/eval_lut_sr.py:750-775 (= eval_lut_warp.py:308-333): same keys, float32 [17^4, oC] tables."""
    lut = dict()
    for s in range(opt.stages):
        stage = s + 1
        cur_modes, rots, oC = opt.modes, ["0"], 1
        if stage == opt.stages:  # hyper stage
            cur_modes, rots, oC = opt.modes2, ["0", "1"], (1 if opt.linear else 3)
        for mode in cur_modes:
            for r in rots:
                key = "s{}_{}r{}".format(str(s + 1), mode, r)
                path = os.path.join(opt.expDir, "{}_s{}_{}r{}.npy".format(opt.lutName, str(stage), mode, r))
                lut[key] = np.array(np.load(path)).astype(np.float32).reshape(-1, oC)
    return lut


def load_rgb(path):
    """PIL decode; gray images are replicated to three channels (eval_lut_sr.py:527-529)."""
    from PIL import Image
    img = np.array(Image.open(path))
    if len(img.shape) == 2:
        img = np.expand_dims(img, axis=2)
        img = np.concatenate([img, img, img], axis=2)
    return img


def list_pngs(folder):
    files = [f for f in os.listdir(folder) if "png" in f]
    files.sort()
    return files
