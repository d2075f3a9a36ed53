"""Build the CUDA library (sm_100a only) in-tree: ``qampy_b200/lib/libqampy_b200.so``.

    python -m qampy_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is plain C ABI (``include/qampy_b200.h``); Python
binds it with ctypes (``qampy_b200/_lib.py``).  The built file is git-ignored but travels with the
repo snapshot to the GPU box.
"""
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqampy_b200.so")
SOURCES = ["cabi.cu", "eq_apply.cu", "eq_train.cu", "eq_train_gla.cu", "eq_train_fast.cu", "eq_train_fast_l8.cu", "eq_train_fast_l8a.cu", "eq_train_fast_l8b.cu", "eq_train_fast_l8c.cu", "eq_train_fast_l16.cu", "eq_train_la_l8.cu", "eq_train_la_l8a.cu", "eq_train_la_l32.cu", "eq_train_la_l32a.cu",
           "bps.cu", "bps_fast.cu", "bps_par.cu", "pilot_ops.cu", "synth_ops.cu", "decision.cu", "vv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
OBJDIR = os.path.join(LIBDIR, "obj")    # one object per translation unit: a change rebuilds only what includes it
EXTRA_FLAGS = os.environ.get("QB_NVCC_EXTRA", "").split()   # e.g. -DQB_BPS_NR=16 for tuning builds


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libqampy_b200.so")


_INCLUDE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _includes(path, seen=None):
    """The files a source includes with quotes, recursively (the library's own headers)."""
    seen = set() if seen is None else seen
    try:
        text = open(path).read()
    except OSError:
        return seen
    for inc in _INCLUDE.findall(text):
        dep = os.path.normpath(os.path.join(os.path.dirname(path), inc))
        if dep not in seen and os.path.exists(dep):
            seen.add(dep)
            _includes(dep, seen)
    return seen


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _flags_stamp():
    return " ".join(NVCC_FLAGS + EXTRA_FLAGS)


def _stale_sources():
    stamp = os.path.join(OBJDIR, "flags.txt")
    same_flags = os.path.exists(stamp) and open(stamp).read() == _flags_stamp()
    me = os.path.getmtime(os.path.abspath(__file__))
    out = []
    for src in SOURCES:
        path, obj = os.path.join(CSRC, src), _obj(src)
        if not same_flags or not os.path.exists(obj):
            out.append(src)
            continue
        t = os.path.getmtime(obj)
        if any(os.path.getmtime(d) > t for d in [path] + sorted(_includes(path))) or me > t:
            out.append(src)
    return out


def is_stale():
    """The library is older than one of its sources (the objects are a cache: they do not travel to the GPU box)."""
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = set()
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        deps.add(path)
        deps |= _includes(path)
    if any(os.path.getmtime(d) > t for d in deps):
        return True
    stamp = os.path.join(OBJDIR, "flags.txt")
    return os.path.exists(stamp) and open(stamp).read() != _flags_stamp()


def build(force=False, verbose=False, jobs=None):
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC=/opt/gcc/bin/gcc is not nvcc's host compiler
    env.pop("CXX", None)
    todo = list(SOURCES) if force else _stale_sources()
    logs = {}

    def compile_one(src):
        cmd = [_nvcc()] + NVCC_FLAGS + EXTRA_FLAGS + ["-c", "-o", _obj(src) + ".tmp", os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        os.replace(_obj(src) + ".tmp", _obj(src))
        logs[src] = res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=jobs or max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        list(pool.map(compile_one, todo))
    with open(os.path.join(OBJDIR, "flags.txt"), "w") as fh:
        fh.write(_flags_stamp())
    for src, text in logs.items():                      # ptxas -v output per translation unit
        with open(os.path.join(OBJDIR, os.path.splitext(src)[0] + ".ptxas.log"), "w") as fh:
            fh.write(text)
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", tmp] + [_obj(s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as fh:
        for src in SOURCES:
            logp = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".ptxas.log")
            if os.path.exists(logp):
                fh.write(open(logp).read())
    if verbose:
        print("compiled: %s" % (", ".join(todo) or "nothing"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
