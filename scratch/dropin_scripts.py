"""The equaliser calls of the reference's Scripts/*_equalisation.py (their dtype, length, taps, methods, step-size rule)
through the drop-in API on the GPU against the oracle port built with the reference's flags (all host threads; the
reference itself trains its modes on 2 threads).  Prints one line per shape and dtype."""
import json
import os
import sys
import time
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
import torch
import cpu_oracle as co
from qampy_b200 import equalisation as eq, synth

SHAPES = (
    # name, M, nsym, ntaps, mu, methods, adaptive
    ("64_qam_equalisation.py", 64, 2 ** 17, 13, (0.19e-2, 0.19e-2), ("mcma", "mddma"), (True, True)),
    ("mrde_equaliser.py", 16, 2 ** 18, 30, (1e-3, 0.5e-3), ("mcma", "mrde"), (False, False)),
    ("32_qam_equalisation.py", 32, 10 ** 6, 11, (1e-3, 1e-3), ("mcma", "sbd"), (False, False)),
)


def tm(f, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


out = []
for name, M, nsym, ntaps, mu, methods, adaptive in SHAPES:
    E64, _ = synth.synth_signal(M, nsym, seed=3, snr_db=25.0, beta=0.01, theta=np.pi / 3, dgd=30e-12, device='cuda')
    for dt in (np.complex128, np.complex64):
        E = E64.cpu().numpy().astype(dt)
        g = lambda: eq.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods, adaptive_stepsize=adaptive)
        c = lambda: co.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods, adaptive_stepsize=adaptive,
                                              kind="fast_native")
        tg, rg = tm(g)
        tc, rc = tm(c, 1)
        d = float(np.sqrt(np.mean(np.abs(rg[0] - rc[0]) ** 2)))
        rec = {"script": name, "dtype": np.dtype(dt).name, "symbols": nsym, "ntaps": ntaps, "methods": methods,
               "adaptive": adaptive, "gpu_s": tg, "cpu_s": tc, "speedup": tc / tg, "rms_diff_vs_fast_oracle": d}
        out.append(rec)
        print(json.dumps(rec), flush=True)
