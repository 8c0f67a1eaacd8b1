#!/usr/bin/env python
"""Multi-GPU box: how the host-to-host path (LerfSR.run_host) behaves when every rank moves its results over PCIe at
once.  For several (streams, row bands) settings: barrier, 3 steps of 8 frames, max over ranks; and, for the ceiling, a
barrier-synchronised plain pinned D2H copy of the same bytes on every rank.  torchrun-launched; rank 0 prints a table."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lerf_pytorch_b200 as lp  # noqa: E402

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
sys.stdout.flush()
real_out = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
luts = lp.LutSet(lp.load_lut_dict(bench.LUT_DIR), device=dev)
sr = lp.LerfSR(luts, 4)
B = 8
frames = bench.natural_frames_gpu(B, 3000 + rank, dev)
oH, oW = sr.set_shape(bench.H, bench.W, 3)
host_in = torch.empty((B, bench.H, bench.W, 3), dtype=torch.uint8).pin_memory()
host_in.copy_(frames.cpu())
host_out = torch.empty((B, oH, oW, 3), dtype=torch.uint8).pin_memory()
dbuf = torch.empty((B, oH, oW, 3), dtype=torch.uint8, device=dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


gb = host_out.numel() / 1e9
rows = []
ms = timed(lambda: (host_out.copy_(dbuf, non_blocking=True), torch.cuda.synchronize()), 3)
rows.append(("plain pinned D2H copy, all ranks at once", ms, gb / (ms * 1e-3)))


# r2: can the host side absorb more with another kind of pinned destination?
def cudart():
    import ctypes
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + ["libcudart.so.12", "libcudart.so"]
    for c in cands:
        try:
            return ctypes.CDLL(c)
        except OSError:
            pass
    return None


def host_tensor(ptr, n):
    import ctypes
    return torch.frombuffer((ctypes.c_uint8 * n).from_address(ptr), dtype=torch.uint8)


def try_variant(name, make):
    try:
        t, free = make()
        ms_ = timed(lambda: (t.copy_(dbuf.view(-1), non_blocking=True), torch.cuda.synchronize()), 3)
        rows.append((name, ms_, gb / (ms_ * 1e-3)))
        free()
    except Exception as ex:  # pragma: no cover
        rows.append((name + " -- unavailable: %r" % (ex,), float("nan"), float("nan")))


def make_wc():
    import ctypes
    rt = cudart()
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(host_out.numel()), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    if rc != 0:
        raise RuntimeError("cudaHostAlloc(WriteCombined) -> %d" % rc)
    return host_tensor(p.value, host_out.numel()), lambda: rt.cudaFreeHost(p)


def make_huge(flags_huge):
    import ctypes
    import mmap
    n = (host_out.numel() + (1 << 21) - 1) >> 21 << 21
    fl = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | (getattr(mmap, "MAP_HUGETLB", 0x40000) if flags_huge else 0)
    m = mmap.mmap(-1, n, flags=fl)
    if not flags_huge:
        m.madvise(getattr(mmap, "MADV_HUGEPAGE", 14))
    buf = (ctypes.c_uint8 * n).from_buffer(m)
    ptr = ctypes.addressof(buf)
    ctypes.memset(ptr, 0, n)  # touch: the pages exist (and are huge where the kernel agrees) before they are registered
    rt = cudart()
    rc = rt.cudaHostRegister(ctypes.c_void_p(ptr), ctypes.c_size_t(n), ctypes.c_uint(0))
    if rc != 0:
        raise RuntimeError("cudaHostRegister -> %d" % rc)
    t = torch.frombuffer(buf, dtype=torch.uint8)[:host_out.numel()]
    return t, lambda: rt.cudaHostUnregister(ctypes.c_void_p(ptr))


try_variant("  ... into cudaHostAlloc(WriteCombined)", make_wc)
try_variant("  ... into MAP_HUGETLB + cudaHostRegister", lambda: make_huge(True))
try_variant("  ... into THP (madvise) + cudaHostRegister", lambda: make_huge(False))
# half of the ranks at a time: is the ceiling per box (shared) or per rank?
if world > 1:
    def halves():
        for par in (0, 1):
            if rank % 2 == par:
                host_out.copy_(dbuf, non_blocking=True)
            torch.cuda.synchronize()
            dist.barrier()
    ms = timed(halves, 3)
    rows.append(("plain copy, even ranks then odd ranks", ms, gb / (ms * 1e-3)))
for depth, bands in ((3, 4), (3, 1), (2, 4), (2, 1), (1, 1), (4, 8)):
    sr._slots = None
    ms = timed(lambda: sr.run_host(host_in, host_out, depth=depth, bands=bands), 3)
    rows.append(("run_host depth %d, bands %d" % (depth, bands), ms, gb / (ms * 1e-3)))
if rank == 0:
    real_out.write("N = %d GPUs; per rank: 8 frames, %.2f GB D2H per step\n" % (world, gb))
    for name, ms, rate in rows:
        real_out.write("%-42s %8.2f ms per step  %6.1f GB/s per rank  %7.1f GB/s box  %9.0f MPix/s\n" % (
            name, ms, rate, rate * world, world * B * oH * oW / 1e6 / (ms * 1e-3)))
    real_out.flush()
if world > 1:
    dist.destroy_process_group()
