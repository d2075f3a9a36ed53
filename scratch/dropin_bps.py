"""phaserec.bps through the drop-in API (host arrays in and out) on ONE capture of two polarisations, complex64 and
complex128, against the oracle port built with the reference's flags on all host cores (index search only: the CPU
side leaves out the unwrap and the rotation)."""
import json, os, sys, time
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import cpu_oracle as co
from qampy_b200 import phaserecovery as pr, theory

def tm(f, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r

for M, A, N, n in ((64, 64, 45, 2 ** 17), (16, 32, 21, 10 ** 6), (64, 64, 45, 10 ** 6)):
    al64 = theory.normalised_symbols(M)
    rng = np.random.default_rng(1)
    x = al64[rng.integers(0, M, (2, n))] * np.exp(1j * 0.1) + 0.03 * (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n)))
    for dt in (np.complex128, np.complex64):
        E, al = x.astype(dt), al64.astype(dt)
        ang = np.linspace(-np.pi / 4, np.pi / 4, A, endpoint=False, dtype=E.real.dtype).reshape(1, -1)
        tg, (out, ph) = tm(lambda: pr.bps(E, A, al, N))
        tc, idx = tm(lambda: co.bps_streams(E, ang, al, N, kind="fast_native"), n=1)
        print(json.dumps({"call": "bps", "dtype": np.dtype(dt).name, "M": M, "angles": A, "N": N, "symbols": n, "gpu_s": tg, "cpu_s": tc,
                          "speedup": tc / tg}), flush=True)
