#!/usr/bin/env python
"""Row-band sharding on real GPUs (SURVEY.md 8e case 2, BASELINE.json cfg-5): LeRF-G x8 SR of ONE synthetic 3840x2160 frame
to 30720x17280, rank g computing output rows [g*oH/G, (g+1)*oH/G).  No collective on the data path: every rank holds the
25 MB input (it needs only its band + a 7-row halo, which the C ABI derives from the band) and keeps its output band.

    python scripts/bench_rowband.py                                   # one GPU, whole frame
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_rowband.py --steps K --warmup W

Prints one JSON line on rank 0 in bench.py's format with "scaling": "strong" (the total work is fixed), the time being
the max over ranks.  Each rank also checks its band against the same rows computed from the input band + halo alone.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

H, W, C, S = 2160, 3840, 3, 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gpus", type=int, default=None)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")  # stdout carries only the JSON line; NCCL's version banner on fd 1 goes to stderr
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
    saved = bench.H, bench.W
    bench.H, bench.W = H, W
    frames = bench.natural_frames_gpu(2, 5000, dev)  # the SAME two frames on every rank (seed 5000); steps alternate
    bench.H, bench.W = saved
    sr = lp.LerfSR(luts, S)
    oH, oW = sr.set_shape(H, W, C)
    y0, y1 = lp.row_bands(oH, world, align=S)[rank]  # band edges on cell boundaries
    out = torch.empty((1, C, oH, oW), dtype=torch.float32, device=dev)  # only rows [y0, y1) are ever written

    def step(i):
        sr(frames[i % 2], out_format="f32", rows=(y0, y1), out=out)

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps

    # parity of the band: the same rows from the input band + 7-row halo only (what a rank would be sent)
    last = (args.steps - 1) % 2
    r0, r1 = max(y0 // S - 7, 0), min((y1 + S - 1) // S + 7, H)
    crop = lp.LerfSR(luts, S)(frames[last][r0:r1].contiguous(), out_format="f32")
    ok = bool(torch.equal(crop[:, y0 - r0 * S:y1 - r0 * S], out[0, :, y0:y1]))
    flag = torch.tensor([1 if ok else 0], device=dev)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        peak, src = bench.measured_peak_gbs()
        byt = C * H * W + C * oH * oW * 4
        json_out.write(json.dumps({
            "metric": "output MPix/s (LeRF-G x8 SR, one 3840x2160 frame, row-band sharded)", "value": oH * oW / 1e6 / (ms * 1e-3),
            "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "data": "synthetic",
            "config": {"workload": "cfg-5: LeRF-G x8 SR of one synthetic 3840x2160 frame -> 30720x17280, uint8 in -> float32 planar out",
                       "sharding": "output row bands, 7-input-row halo, no collective", "band_rows_per_rank": y1 - y0,
                       "l2": "each step writes %.2f GB per GPU and alternates between two input frames" % (C * (y1 - y0) * oW * 4 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": byt / (ms * 1e-3) / 1e9 / world / peak, "peak_source": src,
                         "note": "whole path, algorithmic bytes C*H*W + 4*C*oH*oW, per GPU"},
            "band_equals_halo_crop_on_every_rank": bool(flag.item())}) + "\n")
        json_out.flush()
    if dist is not None:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
