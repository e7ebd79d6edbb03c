"""Build librovr_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Run as a script or call build(). nvcc cross-compiles without a GPU. The library is a single
translation unit (csrc/rovr_b200.cu) so there is no link step beyond cudart.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rovr_b200.cu")
OUT = os.path.join(HERE, "librovr_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
DEPS.append(os.path.join(HERE, "..", "include", "rovr_b200.h"))


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build librovr_b200.so")
    return exe


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC", "--cudart", "static", "-o", OUT, SRC,
    ]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building librovr_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
