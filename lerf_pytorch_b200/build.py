"""Builds liblerf_b200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build())."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "liblerf_b200.so")
SOURCES = ["lut.cu", "lut_cell.cu", "lut_pw.cu", "resample.cu", "resample_int.cu", "resample_tile.cu", "warp_fixed.cu", "fused.cu", "pipeline.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, "csrc", s) for s in SOURCES + ["common.cuh", "lut_cell.cuh", "lut_rm.cuh", "lut_cell_body.cuh", "lut_mix.cuh", "lut_mt.cuh", "lut_pw.cuh", "resample_int.cuh"]] + [os.path.join(ROOT, "include", "lerf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + \
          [os.path.join(HERE, "csrc", s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a wrapper gcc that nvcc must not pick up
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
