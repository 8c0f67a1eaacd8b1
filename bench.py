#!/usr/bin/env python
"""Headline benchmark: output MPix/s of LeRF-G x4 SR on synthetic 2040x1356 frames (BASELINE.json cfg-3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the whole LUT inference path (stage 1 -> stage 2 -> resampling) over a batch of
`--frames` resident uint8 frames per GPU (per-image sharding, no collective on the data path; "scaling":
"weak").  Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, C, SCALE = 1356, 2040, 3, 4  # cfg-3: DIV2K-sized 2040x1356 (WxH) frame, x4
LUT_DIR = os.path.join(ROOT, "tests", "golden", "luts", "lerf-g")
METRIC = "output MPix/s (LeRF-G x4 SR, 2040x1356 frames)"
WORKLOAD = "cfg-3: LeRF-G LUT x4 SR of synthetic 2040x1356 frames, uint8 in -> float32 planar out"


def config_block(frames, inp):
    """The `config` of BOTH arms (ours and --impl reference): same workload, so the driver's same_config holds.  What an
    arm actually ran per step (the reference arm times a bounded sample of it) is in its cpu_baseline.sample."""
    return {"workload": WORKLOAD, "frames_per_gpu_per_step": frames, "input": inp, "sharding": "per image, no collective",
            "l2": "each step writes %.2f GB per GPU (>> 126 MB L2) and alternates between two sets of input frames"
                  % (frames * C * H * SCALE * W * SCALE * 4 / 1e9)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# synthetic frames
# ---------------------------------------------------------------------------------------------------
def natural_frames_gpu(n, seed, device):
    """Seeded 'natural-like' uint8 frames [n,H,W,C]: smooth + texture + noise, made on the GPU."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator(device=device).manual_seed(seed)

    def blur(x, sigma):
        r = int(3 * sigma)
        k = torch.exp(-0.5 * (torch.arange(-r, r + 1, device=device, dtype=torch.float32) / sigma) ** 2)
        k = k / k.sum()
        x = F.conv2d(F.pad(x, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
        x = F.conv2d(F.pad(x, (0, 0, r, r), mode="reflect"), k.view(1, 1, -1, 1))
        return x / x.std(dim=(2, 3), keepdim=True)

    out = torch.empty((n, H, W, C), dtype=torch.uint8, device=device)
    for i in range(n):
        z = torch.randn((3, C, 1, H, W), generator=g, device=device)
        f = 128 + 60 * blur(z[0], 6.0) + 25 * blur(z[1], 1.5) + 4 * z[2]
        out[i] = f.round().clamp(0, 255).to(torch.uint8)[:, 0].permute(1, 2, 0)
    return out


def uniform_frames_gpu(n, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randint(0, 256, (n, H, W, C), generator=g, device=device, dtype=torch.uint8)


def natural_frame_numpy(seed, h, w):
    """CPU twin of the generator for the reference arm (same recipe; numpy RNG)."""
    import cv2
    rng = np.random.default_rng(seed)
    out = np.empty((h, w, C), dtype=np.uint8)
    for c in range(C):
        def blur(x, s):
            y = cv2.GaussianBlur(x, (0, 0), s, borderType=cv2.BORDER_REFLECT)
            return y / y.std()
        f = 128 + 60 * blur(rng.standard_normal((h, w)), 6.0) + 25 * blur(rng.standard_normal((h, w)), 1.5) \
            + 4 * rng.standard_normal((h, w))
        out[:, :, c] = np.clip(np.round(f), 0, 255).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period; the timed
    region of the default run is ~140 ms, shorter than nvidia-smi's start-up), nvidia-smi -lms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = self.f = self.thread = None
        self.sm, self.reasons, self.mx, self.stop_flag = [], set(), None, False

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            return int(vis.split(",")[self.idx]) if vis else self.idx
        except (ValueError, IndexError):
            return self.idx

    def _poll(self, nv, h):
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(h))
                for name, bit in self.BITS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.mx, reasons=sorted(self.reasons),
                           samples=len(self.sm), source="nvml")
            return out
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       source="nvidia-smi")
        return out


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------
def bind_near_gpu(gpu_index):
    """Pin this process (and therefore its pinned host buffers, by first touch) to the CPUs NVML reports as local to
    the GPU -- what `numactl` would do for one rank per GPU.  End-to-end numbers move host memory at PCIe rate on every
    rank at once; with 4 ranks on one box the unbound run reached 66 GB/s aggregate against 101 GB/s with 2 ranks.
    Returns a short description for the JSON line; does nothing when the affinity is unknown or not allowed."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[gpu_index]) if vis else gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(near & allowed)
        if not use or set(use) == set(allowed):
            return "unchanged (%d cpus allowed, %d local to the GPU)" % (len(allowed), len(near & allowed))
        os.sched_setaffinity(0, use)
        return "bound to %d of %d cpus (NVML cpu affinity of GPU %d)" % (len(use), len(allowed), phys)
    except Exception as ex:  # pragma: no cover
        return "unchanged (%s)" % type(ex).__name__


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_rate(rows, seed, repeats=1, threads=None):
    """Time the oracle C port (stage 1 + stage 2 + set_shape/resize, float64 out) on a `rows` x 2040 band of a
    cfg-3 frame.  Returns (out MPix/s, seconds per run, threads)."""
    from oracle import lerf_oracle as orc
    orc.build()
    # all the host threads this process may use -- torchrun exports OMP_NUM_THREADS=1, which would make this a 1-core run
    orc.set_threads(threads or host_threads())
    nthreads = orc.max_threads()
    luts = orc.load_luts(LUT_DIR, linear=False)
    img = natural_frame_numpy(seed, rows, W)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out, _, _ = orc.lerf_sr(img, luts, SCALE, SCALE, linear=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    mpix = out.shape[1] * out.shape[2] / 1e6
    return mpix / best, best, nthreads


def reference_rate(rows, workers, seed=3000, repeats=1):
    """Time the UNMODIFIED reference (baseline/_ref, numpy, single-threaded by construction) on `workers` independent
    `rows` x 2040 bands of cfg-3 frames, one process each.  Returns (out MPix/s, seconds, workers) or None if the
    reference is not staged on this box."""
    try:
        from baseline import ref_runner as rr
        if rr.load() is None:
            return None
        imgs = [natural_frame_numpy(seed + i, rows, W) for i in range(workers)]
        secs, opix = rr.time_workers(imgs, LUT_DIR, SCALE, repeats)
        return opix / 1e6 / secs, secs, workers
    except Exception as ex:  # pragma: no cover
        sys.stderr.write("reference_rate: %r\n" % (ex,))
        return None


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  The real thing
    (numpy, one process per host thread, each on its own band) when baseline/_ref is staged, the C/OpenMP oracle port
    otherwise; the port is always timed beside it.  Each step is a bounded sample sized so the run ends in ~2.5 minutes."""
    if rank != 0:
        return 0
    nthr = host_threads()
    steps_total = max(1, args.steps + args.warmup)
    line_extra = {}
    kind = "port"
    ref = reference_rate(8, nthr)  # calibration: 8-row bands
    if ref is not None:
        kind = "reference"
        budget = 110.0 / steps_total
        rows = int(max(8, min(64, 8 * budget / ref[1])))
        if args.sample_rows > 0:
            rows = int(min(H, args.sample_rows))
        from baseline import ref_runner as rr
        imgs = [natural_frame_numpy(3000 + i, rows, W) for i in range(nthr)]
        for _ in range(args.warmup):
            rr.time_workers(imgs, LUT_DIR, SCALE)
        t = time.perf_counter()
        opix = 0
        for _ in range(args.steps):
            _, n = rr.time_workers(imgs, LUT_DIR, SCALE)
            opix += n
        dt = (time.perf_counter() - t) / args.steps
        mpix = opix / args.steps / 1e6
        val = mpix / dt
        sample = ("per step: %d processes x one %dx%d band of a cfg-3 frame each (x4 -> %.2f out MPix in all), natural-like seeds "
                  "3000.., through the reference's FourSimplexInterpFaster + SteeringGaussianResize2dNumpy (baseline/_ref)" % (nthr, rows, W, mpix))
        prate, psecs, pthr = cpu_port_rate(H // 4, 3000)
        line_extra["port"] = {"value": prate, "unit": "MPix/s", "cores": pthr, "kind": "port",
                              "sample": "one %dx%d band (x4 -> %.1f out MPix) in %.2f s, oracle/lerf_oracle.c (OpenMP)" %
                                        (H // 4, W, (H // 4) * SCALE * W * SCALE / 1e6, psecs)}
        dtype = "f64 (numpy reference)"
    else:
        # bounded sample: calibrate on a 64-row band, then size the band so (steps + warmup) runs end in ~2.5 minutes
        rate0, t0, nthr = cpu_port_rate(64, 3000)
        budget = 150.0 / steps_total
        rows = int(max(32, min(H, 64 * budget / t0)))
        if args.sample_rows > 0:
            rows = int(min(H, args.sample_rows))
        from oracle import lerf_oracle as orc
        luts = orc.load_luts(LUT_DIR, linear=False)
        img = natural_frame_numpy(3000, rows, W)
        for _ in range(args.warmup):
            out, _, _ = orc.lerf_sr(img, luts, SCALE, SCALE, linear=False)
        t = time.perf_counter()
        for _ in range(args.steps):
            out, _, _ = orc.lerf_sr(img, luts, SCALE, SCALE, linear=False)
        dt = (time.perf_counter() - t) / args.steps
        mpix = out.shape[1] * out.shape[2] / 1e6
        val = mpix / dt
        sample = "per step: %dx%d band of a cfg-3 frame (x4 -> %.2f out MPix), natural-like seed 3000, C/OpenMP oracle port" % (rows, W, mpix)
        dtype = "f64 (CPU oracle port)"
    cpu = {"value": val, "unit": "MPix/s", "cores": nthr, "kind": kind, "sample": sample}
    cpu.update(line_extra)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": config_block(args.frames, args.input),
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def rowband_arm(lp, luts, dev, rank, n_gpus, dist, torch, steps=5, warmup=3):
    """BASELINE.json cfg-5: ONE synthetic 3840x2160 frame x8 -> 30720x17280; rank g computes the output rows of its band
    from the whole input (the C ABI derives the input rows + 7-row halo the band needs; band edges sit on cell
    boundaries).  No collective: the bands stay sharded.  Strong scaling: total work is fixed, value = the frame's
    output MPix / max-over-ranks time.  Every rank also checks its band, bit for bit, against the same rows computed
    from the input band + 7-row halo alone (what a rank would be sent)."""
    global H, W
    h5, w5, s5 = 2160, 3840, 8
    saved = H, W
    H, W = h5, w5
    try:
        frames = natural_frames_gpu(2, 5000, dev)  # the SAME two frames on every rank; steps alternate
    finally:
        H, W = saved
    sr5 = lp.LerfSR(luts, s5)
    oH5, oW5 = sr5.set_shape(h5, w5, C)
    y0, y1 = lp.row_bands(oH5, n_gpus, align=s5)[rank]
    out5 = torch.empty((1, C, oH5, oW5), dtype=torch.float32, device=dev)  # only rows [y0, y1) are ever written

    def step(i):
        sr5(frames[i % 2], out_format="f32", rows=(y0, y1), out=out5)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(i)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    last = (steps - 1) % 2
    r0, r1 = max(y0 // s5 - 7, 0), min((y1 + s5 - 1) // s5 + 7, h5)
    crop = lp.LerfSR(luts, s5)(frames[last][r0:r1].contiguous(), out_format="f32")
    ok = bool(torch.equal(crop[:, y0 - r0 * s5:y1 - r0 * s5], out5[0, :, y0:y1]))
    flag = torch.tensor([1 if ok else 0], device=dev)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    del out5, crop
    torch.cuda.empty_cache()
    mpix = oH5 * oW5 / 1e6
    bytes_ = C * h5 * w5 + C * oH5 * oW5 * 4
    peak, _ = measured_peak_gbs()
    return {"workload": "cfg-5: LeRF-G x8 SR of one synthetic 3840x2160 frame -> 30720x17280, uint8 in -> float32 planar out, output "
                        "row bands across the ranks (7-input-row halo), no collective",
            "scaling": "strong", "n_gpus": n_gpus, "value": mpix / (ms * 1e-3), "unit": "MPix/s", "ms_per_frame": ms, "steps": steps,
            "band_rows_per_rank": y1 - y0, "path_frac_per_gpu": bytes_ / n_gpus / (ms * 1e-3) / 1e9 / peak,
            "band_equals_halo_crop_on_every_rank": bool(flag.item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8, help="frames per GPU per step")
    ap.add_argument("--input", default="natural", choices=["natural", "uniform"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-arms", action="store_true", help="skip the uniform-input, uint8 and cfg-5 row-band arms")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--sample-rows", type=int, default=0,
                    help="reference arm: rows of the frame band timed per step (0 = sized so the run ends in ~2.5 minutes)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    # stdout carries exactly ONE line, the JSON: libraries that write to fd 1 behind Python's back (NCCL prints
    # "NCCL version ..." there at communicator creation) are sent to stderr for the whole run
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "lerf_pytorch_b200", "liblerf_b200.so")):
        ge.build()
    import lerf_pytorch_b200 as lp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_binding = bind_near_gpu(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    lut_dict = lp.load_lut_dict(LUT_DIR, linear=False)
    luts = lp.LutSet(lut_dict, linear=False, device=dev)
    sr = lp.LerfSR(luts, SCALE)
    B = args.frames
    gen = natural_frames_gpu if args.input == "natural" else uniform_frames_gpu
    pool = gen(2 * B, 3000 + 1000 * rank, dev)          # 2B distinct frames; steps alternate halves
    sr.set_shape(H, W, C)
    oH, oW = sr.out_sz
    out = sr.alloc_out(B, C, "f32", dev)                # 4.25 GB at B=8: every step streams far more than the 126 MB L2
    out_mpix_step = B * oH * oW / 1e6

    names = ("lut_stage1", "lut_stage2", "resize_sr")

    def timed_arm(frames2, fmt, out_buf, steps, warmup, record_kernels, sample_clocks=False):
        """W warm-up steps, then K timed steps bracketed by barrier + synchronize; max over ranks.  Returns
        (ms per step, per-kernel ms or None, launches, clocks or None)."""
        ev = []

        def step(i, record):
            frames = frames2[(i % 2) * B:(i % 2 + 1) * B]
            if record:
                marks = [torch.cuda.Event(enable_timing=True)]
                marks[0].record()

                def rec(_name):
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    marks.append(e)
                sr(frames, out_format=fmt, out=out_buf, record=rec)
                ev.append(marks)
            else:
                sr(frames, out_format=fmt, out=out_buf)

        for i in range(warmup):
            step(i, False)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(local_rank)
            sampler.start()
        lp.lib().lerf_launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i, record_kernels)
        e1.record()
        torch.cuda.synchronize()
        n_launch = int(lp.lib().lerf_launch_count())
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        clk_ = sampler.stop() if sampler is not None else None
        t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        per_ = None
        if record_kernels:
            per_ = {n: 0.0 for n in names}
            for marks in ev:
                for k, n in enumerate(names):
                    per_[n] += marks[k].elapsed_time(marks[k + 1])
            per_ = {n: v / len(ev) for n, v in per_.items()}
        return float(t_ms.item()) / steps, per_, n_launch, clk_

    ms_step, per, launches, clk = timed_arm(pool, "f32", out, args.steps, args.warmup, True, sample_clocks=True)
    value = n_gpus * out_mpix_step / (ms_step * 1e-3)

    top = max(per, key=per.get)
    P = B * C
    kbytes = {  # algorithmic bytes per launch (DESIGN.md "Kernels"): compulsory reads + writes of each kernel
        "lut_stage1": P * H * W * 1 + P * H * W * 1,
        "lut_stage2": P * H * W * 1 + P * H * W * 3,
        "resize_sr": P * H * W * (1 + 3) + P * oH * oW * 4,
    }
    peak, peak_src = measured_peak_gbs()
    path_bytes = P * H * W * 1 + P * oH * oW * 4          # C*H*W*b_in + C*oH*oW*b_out (u8 -> f32)
    achieved = kbytes[top] / (per[top] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None, "peak_source": peak_src,
        "kernel_ms": per, "kernel_share": {n: per[n] / sum(per.values()) for n in names},
        "path": {"bytes_per_step": path_bytes, "achieved": path_bytes / (ms_step * 1e-3) / 1e9,
                 "frac": path_bytes / (ms_step * 1e-3) / 1e9 / peak, "frac_of_8TBs": path_bytes / (ms_step * 1e-3) / 8e12},
    }
    tfile = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tfile):
        try:
            t = json.load(open(tfile)).get(top)
            roofline["traffic"] = t.get("dram_bytes_per_launch") if isinstance(t, dict) else t
            roofline["traffic_note"] = ("DRAM bytes of one %s launch at this launch size, ncu --set full (profiles/traffic.json); "
                                        "algorithmic bytes per launch: %d" % (top, kbytes[top]))
        except Exception:
            pass

    lfile = os.path.join(ROOT, "profiles", "limiters.json")  # what bounds the kernel instead of HBM, from the same ncu capture
    if os.path.exists(lfile):
        try:
            roofline["limiter"] = json.load(open(lfile)).get(top)
        except Exception:
            pass

    # the other input class of SURVEY.md 8d: uniform-random pixels (adversarial for the tables), same run, same kernels
    other = None
    if not args.no_extra_arms:
        okind = "uniform" if args.input == "natural" else "natural"
        opool = (uniform_frames_gpu if okind == "uniform" else natural_frames_gpu)(2 * B, 3000 + 1000 * rank, dev)
        oms, oper, _, _ = timed_arm(opool, "f32", out, max(3, args.steps // 2), 3, True)
        other = {"input": okind, "value": n_gpus * out_mpix_step / (oms * 1e-3), "unit": "MPix/s", "ms_per_step": oms,
                 "kernel_ms": oper, "path_frac": path_bytes / (oms * 1e-3) / 1e9 / peak}
        del opool
    # secondary mode of SURVEY.md 8d: uint8 in -> uint8 out (the fused a10 epilogue; what run_host and the eval adapters use)
    path_u8 = None
    if not args.no_extra_arms:
        path_u8 = {}
        for fmt in ("u8", "u8_hwc"):
            obuf = sr.alloc_out(B, C, fmt, dev)
            ums, uper, _, _ = timed_arm(pool, fmt, obuf, max(3, args.steps // 2), 3, True)
            ubytes = P * H * W * 1 + P * oH * oW * 1
            path_u8[fmt] = {"bytes_per_step": ubytes, "ms_per_step": ums, "value": n_gpus * out_mpix_step / (ums * 1e-3),
                            "achieved": ubytes / (ums * 1e-3) / 1e9, "frac": ubytes / (ums * 1e-3) / 1e9 / peak, "kernel_ms": uper}
            del obuf
        roofline["path_u8"] = path_u8
        roofline["path_u8_note"] = ("3.19 B per output pixel (141.1 MB per frame): a quarter of the float32 path's bytes for the same "
                                    "instruction-bound kernels, so the HBM fraction is about a quarter of path.frac by construction")

    # what was timed is what the reference computes: one frame of the timed batch against the oracle, outside the timing
    parity = None
    if rank == 0 and not args.no_parity_check:
        try:
            from oracle import lerf_oracle as orc
            orc.build()
            orc.set_threads(host_threads())
            fr = pool[:1]
            got = sr(fr, out_format="f32")[0].cpu().numpy().astype(np.float64)
            gfeat, gcodes = sr.stages(fr)
            gu8 = sr(fr, out_format="u8_hwc")[0].cpu().numpy()
            ref, rfeat, rcodes = orc.lerf_sr(fr[0].cpu().numpy(), orc.load_luts(LUT_DIR, linear=False), SCALE, SCALE, linear=False)
            err = float(np.max(np.abs(got - ref)))
            lsb = int(np.max(np.abs(gu8.astype(np.int32) - orc.to_uint8_hwc(ref).astype(np.int32))))
            exact = bool(np.array_equal(gfeat.cpu().numpy(), rfeat) and np.array_equal(gcodes.cpu().numpy(), rcodes))
            parity = {"checked": bool(exact and err <= 1e-4 and lsb <= 1), "frame": "frame 0 of the timed batch, %dx%d, all %d output samples" % (W, H, ref.size),
                      "feat_codes_bit_exact": exact, "f32_max_abs_err": err, "f32_tolerance": 1e-4, "u8_max_lsb": lsb,
                      "against": "oracle/lerf_oracle.c (float64), pinned to reference-generated goldens by tests/test_oracle_golden.py"}
        except Exception as ex:  # pragma: no cover
            parity = {"checked": False, "error": repr(ex)}

    # cfg-5 (one 3840x2160 frame x8, output row bands across the ranks): the strong-scaling case of SURVEY.md 8e
    rowband = None
    if not args.no_extra_arms:
        try:
            rowband = rowband_arm(lp, luts, dev, rank, n_gpus, dist, torch)
        except Exception as ex:  # pragma: no cover
            rowband = {"value": None, "error": repr(ex)}

    # end to end through the public host API: pinned uint8 frames in, pinned uint8 HWC frames out, copies inside
    e2e = None
    try:
        if args.e2e_steps <= 0:
            raise RuntimeError("skipped (--e2e-steps 0)")
        host_in = torch.empty((B, H, W, C), dtype=torch.uint8).pin_memory()
        host_in.copy_(pool[:B].cpu())
        host_out = torch.empty((B, oH, oW, C), dtype=torch.uint8).pin_memory()
        sr.run_host(host_in, host_out)  # warm-up (allocates the slots)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            sr.run_host(host_in, host_out)
        b.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item()) / args.e2e_steps
        chk = int(host_out[0, oH // 2, oW // 2].sum())  # touch the result on the host
        # what the D2H link alone delivers for the same bytes (one plain pinned copy): the ceiling of this e2e figure
        dbuf = torch.empty((B, oH, oW, C), dtype=torch.uint8, device=dev)
        host_out.copy_(dbuf, non_blocking=True)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()  # every rank copies at the same time: the ceiling of N ranks sharing the host, not of one alone
        torch.cuda.synchronize()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        for _ in range(3):
            host_out.copy_(dbuf, non_blocking=True)
        b2.record()
        torch.cuda.synchronize()
        t3 = torch.tensor([a2.elapsed_time(b2) / 3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        link_gbs = dbuf.numel() / (float(t3.item()) * 1e-3) / 1e9
        del dbuf
        e2e = {"value": n_gpus * out_mpix_step / (ms2 * 1e-3), "unit": "MPix/s", "h2d_bytes_per_step": B * H * W * C,
               "d2h_bytes_per_step": B * oH * oW * C, "ms_per_step": ms2, "steps": args.e2e_steps,
               "api": "LerfSR.run_host(pinned uint8 HWC in, pinned uint8 HWC out), 3 streams x 4 row bands per frame", "checksum": chk,
               "cpu_binding": cpu_binding, "d2h_link_GBps_plain_copy": link_gbs,
               "d2h_link_note": "per rank, all ranks copying at once after a barrier (max over ranks)",
               "d2h_GBps_achieved": B * oH * oW * C / (ms2 * 1e-3) / 1e9}
    except Exception as ex:  # pragma: no cover
        e2e = {"value": None, "error": repr(ex)}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            rate, secs, nthreads = cpu_port_rate(H // 2, 3000)
            port = {"value": rate, "unit": "MPix/s", "cores": nthreads, "kind": "port",
                    "sample": "one %dx%d half frame of the same workload (x4 -> %.1f out MPix) in %.1f s, oracle/lerf_oracle.c (OpenMP)" %
                              (H // 2, W, (H // 2) * SCALE * W * SCALE / 1e6, secs)}
            ref = reference_rate(24, host_threads())  # the unmodified reference (baseline/_ref), one process per host thread
            if ref is not None:
                cpu = {"value": ref[0], "unit": "MPix/s", "cores": ref[2], "kind": "reference",
                       "sample": "%d processes x one 24x%d band of a cfg-3 frame each (x4 -> %.1f out MPix in all) in %.1f s, the "
                                 "reference's own numpy functions (baseline/_ref)" % (ref[2], W, ref[2] * 24 * SCALE * W * SCALE / 1e6, ref[1]),
                       "port": port}
            else:
                cpu = port
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 LUT stages (dp4a on int8 tables); f64 exponent + f32 ex2/accumulate resampling",
            "data": "synthetic",
            "config": config_block(B, args.input),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
            "parity_checked": bool(parity and parity.get("checked")), "parity": parity,
            "value_uniform" if args.input == "natural" else "value_natural": other["value"] if other else None,
            "other_input": other, "rowband_cfg5": rowband,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
