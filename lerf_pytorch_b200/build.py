"""Builds liblerf_b200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build()).

`python -m lerf_pytorch_b200.build --experiments` also builds liblerf_b200_exp.so: the same sources with
-DLERF_EXPERIMENTS, i.e. with every tuning variant and superseded kernel compiled in (scripts/kbench.py and the variant
tests load it through LERF_B200_EXPERIMENTS=1).  The product library carries the production kernels plus the second
implementations the parity tests need."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "liblerf_b200.so")
LIB_EXP = os.path.join(HERE, "liblerf_b200_exp.so")
SOURCES = ["lut.cu", "lut_cell.cu", "lut_pw.cu", "resample.cu", "resample_int.cu", "resample_tile.cu", "warp_fixed.cu",
           "fused.cu", "pipeline.cu", "lut_ft.cu", "png.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, experiments=False):
    lib = LIB_EXP if experiments else LIB
    if not force and not needs_build(lib):
        return lib
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DLERF_EXPERIMENTS"] if experiments else []) + \
        ["-I", os.path.join(ROOT, "include"), "-o", lib] + [os.path.join(HERE, "csrc", s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a wrapper gcc that nvcc must not pick up
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
    if "--experiments" in sys.argv:
        print(build(force=True, verbose="-v" in sys.argv, experiments=True))
