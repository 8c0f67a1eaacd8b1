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
    """Time the oracle C port (stage 1 + stage 2 + set_shape/resize + uint8 epilogue) on a `rows` x 2040 band of a
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
        orc.to_uint8_hwc(out)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    mpix = out.shape[1] * out.shape[2] / 1e6
    return mpix / best, best, nthreads


def run_reference_arm(args, rank):
    if rank != 0:
        return 0
    # bounded sample: calibrate on a 64-row band, then size the band so (steps + warmup) runs end in ~2.5 minutes
    rate0, t0, nthreads = cpu_port_rate(64, 3000)
    budget = 150.0 / max(1, args.steps + args.warmup)
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
        orc.to_uint8_hwc(out)
    dt = (time.perf_counter() - t) / args.steps
    mpix = out.shape[1] * out.shape[2] / 1e6
    val = mpix / dt
    sample = "per step: %dx%d band of a cfg-3 frame (x4 -> %.2f out MPix), natural-like seed 3000" % (rows, W, mpix)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (CPU oracle port)", "data": "synthetic",
        "config": {"workload": "cfg-3: LeRF-G LUT x4 SR of synthetic 2040x1356 frames", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "MPix/s", "cores": nthreads, "kind": "port", "sample": sample,
                         "note": "the reference is pure Python/numpy and cannot travel to the GPU box; this is the C oracle "
                                 "port of its algorithm (oracle/lerf_oracle.c, OpenMP). The numpy reference itself measured "
                                 "0.075 MPix/s on one core in the build container (BASELINE.md)."},
        "e2e": {"value": val, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
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
    ev = []

    def step(i, record=True):
        frames = pool[(i % 2) * B:(i % 2 + 1) * B]
        if record:
            marks = [torch.cuda.Event(enable_timing=True)]
            marks[0].record()

            def rec(_name):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
            sr(frames, out_format="f32", out=out, record=rec)
            ev.append(marks)
        else:
            sr(frames, out_format="f32", out=out)

    for i in range(args.warmup):
        step(i, record=False)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    lp.lib().lerf_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    launches = int(lp.lib().lerf_launch_count())
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    value = n_gpus * out_mpix_step / (ms_step * 1e-3)

    # per-kernel device time over the timed region (CUDA events on the launching stream)
    per = {n: 0.0 for n in names}
    for marks in ev:
        for k, n in enumerate(names):
            per[n] += marks[k].elapsed_time(marks[k + 1])
    per = {n: v / len(ev) for n, v in per.items()}
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
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        host_out.copy_(dbuf, non_blocking=True)
        b2.record()
        torch.cuda.synchronize()
        link_gbs = dbuf.numel() / (a2.elapsed_time(b2) * 1e-3) / 1e9
        del dbuf
        e2e = {"value": n_gpus * out_mpix_step / (ms2 * 1e-3), "unit": "MPix/s", "h2d_bytes_per_step": B * H * W * C,
               "d2h_bytes_per_step": B * oH * oW * C, "ms_per_step": ms2, "steps": args.e2e_steps,
               "api": "LerfSR.run_host(pinned uint8 HWC in, pinned uint8 HWC out), 3 streams x 4 row bands per frame", "checksum": chk,
               "cpu_binding": cpu_binding, "d2h_link_GBps_plain_copy": link_gbs,
               "d2h_GBps_achieved": B * oH * oW * C / (ms2 * 1e-3) / 1e9}
    except Exception as ex:  # pragma: no cover
        e2e = {"value": None, "error": repr(ex)}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            rate, secs, nthreads = cpu_port_rate(H // 2, 3000)
            cpu = {"value": rate, "unit": "MPix/s", "cores": nthreads, "kind": "port",
                   "sample": "one %dx%d half frame of the same workload (x4 -> %.1f out MPix) in %.1f s" %
                             (H // 2, W, (H // 2) * SCALE * W * SCALE / 1e6, secs)}
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 LUT stages (dp4a on int8 tables); f64 exponent + f32 ex2/accumulate resampling",
            "data": "synthetic",
            "config": {"workload": "cfg-3: LeRF-G LUT x4 SR of synthetic 2040x1356 frames, uint8 in -> float32 planar out",
                       "frames_per_gpu_per_step": B, "input": args.input, "sharding": "per image, no collective",
                       "l2": "each step writes %.2f GB per GPU (>> 126 MB L2) and alternates between two sets of input frames"
                             % (B * C * oH * oW * 4 / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
