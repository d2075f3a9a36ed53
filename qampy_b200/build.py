"""Build the CUDA library (sm_100a only) in-tree: ``qampy_b200/lib/libqampy_b200.so``.

    python -m qampy_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is plain C ABI (``include/qampy_b200.h``); Python
binds it with ctypes (``qampy_b200/_lib.py``).  The built file is git-ignored but travels with the
repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqampy_b200.so")
SOURCES = ["cabi.cu", "eq_apply.cu", "eq_train.cu", "eq_train_gla.cu", "eq_train_fast.cu", "eq_train_fast_l8.cu", "eq_train_fast_l8a.cu", "eq_train_fast_l8b.cu", "eq_train_fast_l8c.cu", "eq_train_fast_l16.cu", "eq_train_la_l8.cu", "eq_train_la_l8a.cu", "eq_train_la_l32.cu", "eq_train_la_l32a.cu",
           "bps.cu", "bps_fast.cu", "pilot_ops.cu", "synth_ops.cu", "decision.cu", "vv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--threads", "8"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libqampy_b200.so")


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "qampy_b200.h"))
    return deps


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC=/opt/gcc/bin/gcc is not nvcc's host compiler
    env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as fh:
        fh.write(res.stdout + res.stderr)
    os.replace(tmp, LIB)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
