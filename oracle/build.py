"""Build recipe for the ORACLE (test infrastructure, not product code).

Two builds of ``qampy_oracle.c``:

* ``_build/libqampy_oracle.so``       strict (``-O2 -ffp-contract=off``): the parity oracle;
* ``_build/libqampy_oracle_fast.so``  the reference's own flags (``setup.py:24-31``:
  ``-O3 -ffast-math -march=native -funroll-loops -fopenmp -DNDEBUG``): only ever *timed*
  (bench.py ``cpu_baseline`` / ``--impl reference``).  ``-march=native`` is host specific, so
  bench.py rebuilds this one on the box it runs on (``native=True``); the copy built in the
  development container uses ``-march=x86-64-v3`` so that it can travel.

The reference's own implementation of the path is Pythran source; Pythran is not present in
the image and cannot be installed (no network), so there is no ``oracle/_ref`` build.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "qampy_oracle.c")
DEPS = [SRC, os.path.join(HERE, "qo_kernels.inc")]

STRICT_FLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math"]
FAST_FLAGS = ["-O3", "-ffast-math", "-funroll-loops", "-DNDEBUG"]


def _gcc():
    # the image exports CC=/opt/gcc/bin/gcc, which has no libgomp spec; prefer the system gcc
    for cand in ("/usr/bin/gcc", shutil.which("gcc"), shutil.which("cc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no C compiler found for the oracle build")


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in DEPS)


def _compile(target, flags):
    os.makedirs(BUILD, exist_ok=True)
    tmp = target + ".tmp%d" % os.getpid()
    cmd = [_gcc()] + flags + ["-fopenmp", "-fPIC", "-shared", "-o", tmp, SRC, "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    os.replace(tmp, target)
    return target


def build_strict(force=False):
    target = os.path.join(BUILD, "libqampy_oracle.so")
    if force or _stale(target):
        _compile(target, STRICT_FLAGS)
    return target


def build_fast(native=False, force=False):
    """native=True: -march=native into a host-specific file (rebuilt where it is timed)."""
    if native:
        target = os.path.join(BUILD, "libqampy_oracle_fast_native.so")
        try:
            if force or _stale(target):
                _compile(target, FAST_FLAGS + ["-march=native"])
            return target
        except Exception:
            pass  # fall back to the portable build below
    target = os.path.join(BUILD, "libqampy_oracle_fast.so")
    if force or _stale(target):
        _compile(target, FAST_FLAGS + ["-march=x86-64-v3"])
    return target


def build_all(force=False):
    return build_strict(force), build_fast(False, force)


if __name__ == "__main__":
    print(build_all(force=True))
