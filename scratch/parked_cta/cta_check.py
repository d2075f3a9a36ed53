"""One-stream-per-CTA trainer (csrc/eq_train_cta.cu) against the strict oracle over shapes and edge cases, then its
cycles per symbol.  Run on the GPU box under `timeout` (a deadlock in the warp pipeline would otherwise hang):
    timeout 300 python scratch/cta_check.py
"""
import os
import sys
import time

sys.path.insert(0, '.')
sys.path.insert(0, 'oracle')
import numpy as np
import torch

import cpu_oracle as co
from qampy_b200 import device, synth, theory

dev = torch.device('cuda', 0)


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2))) if np.size(a) else 0.0


def one(M, ntaps, nmodes, nsym, method, nseg=1, adaptive=False, niter=1, mu0=2e-3, odd=False, seed=0):
    E, _ = synth.synth_numpy(M, nsym + 64, nmodes=nmodes, seed=seed + M + ntaps, snr_db=24.0)
    if odd:
        E = np.ascontiguousarray(E[:, :E.shape[1] - 1])
    S = (E.shape[1] - ntaps + 1) // 2 // nseg
    Ed = torch.from_numpy(E).to(dev)
    Ev = device.segment_view(Ed, nseg, S, 2, ntaps)
    L_seg = Ev.shape[2]
    tr = theory.cal_training_symbol_len(2, ntaps, L_seg)
    sy = theory.reshape_symbols(None, method, M, np.complex64, nmodes)
    w = torch.from_numpy(np.tile(theory.init_taps(ntaps, nmodes, np.complex64), (nseg, 1, 1, 1))).to(dev)
    mu = torch.full((nseg, nmodes), mu0, dtype=torch.float32, device=dev)
    err = torch.zeros((nseg, nmodes, tr * niter), dtype=torch.complex64, device=dev)
    device.train_equaliser(Ev, tr, niter, 2, mu, w, None, adaptive, torch.from_numpy(sy).to(dev), method, err, layout="cta")
    torch.cuda.synchronize()
    Es = np.stack([E[:, s * S * 2: s * S * 2 + L_seg] for s in range(nseg)])
    wr = np.tile(theory.init_taps(ntaps, nmodes, np.complex64), (nseg, 1, 1, 1))
    er, wr, mur = co.train_segments(Es, tr, niter, 2, mu0, wr, np.arange(nmodes), adaptive, sy, method, mu_shared=False)
    de = rms(err.cpu().numpy() - er) / max(1.0, rms(er))
    dw = float(np.max(np.abs(w.cpu().numpy() - wr)))
    ok = de < 1e-5 and dw < 2e-5
    print("%s M%d ntaps%d nm%d nsym%d nseg%d %s%s%s: err rms %.2e taps %.2e  %s" % (
        method, M, ntaps, nmodes, nsym, nseg, "adaptive " if adaptive else "", "niter%d " % niter if niter > 1 else "",
        "oddL " if odd else "", de, dw, "ok" if ok else "FAIL"), flush=True)
    return ok


def timing():
    for name, M, nsym, ntaps, method in (("C2", 16, 10 ** 6, 21, "mcma"), ("C3", 64, 10 ** 6, 45, "mcma"),
                                         ("C3", 64, 10 ** 6, 45, "mrde"), ("C3", 64, 10 ** 6, 45, "cma")):
        E, _ = synth.synth_signal(M, nsym, seed=3, snr_db=28.0, device=dev)
        Ev = E[None]
        tr = theory.cal_training_symbol_len(2, ntaps, E.shape[1])
        sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex64, 2)).to(dev)
        for layout in ("latency", "cta"):
            ts = []
            for _ in range(3):
                w = torch.from_numpy(theory.init_taps(ntaps, 2, np.complex64)[None]).to(dev)
                mu = torch.full((1, 2), 1e-3, dtype=torch.float32, device=dev)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy, method, None, layout=layout)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = min(ts)
            print("%s %s ntaps %d %-8s packed=%s: %.2f ms = %.1f cycles/symbol" % (
                name, method, ntaps, layout, os.environ.get("QB_CTA_PACKED", "1"), t, t * 1e-3 * 1.965e9 / tr), flush=True)


if __name__ == "__main__":
    if "--time" in sys.argv:
        timing()
        sys.exit(0)
    t0 = time.time()
    ok = True
    ok &= one(64, 45, 2, 3000, "mcma")
    ok &= one(64, 45, 2, 3000, "mrde")
    ok &= one(16, 21, 2, 3001, "mcma", nseg=3)
    ok &= one(4, 11, 1, 2000, "cma")
    ok &= one(16, 17, 2, 2500, "rde")
    ok &= one(64, 64, 2, 2600, "mrde")
    ok &= one(64, 45, 2, 40, "mcma")                 # shorter than the pipeline is deep
    ok &= one(64, 45, 2, 5, "mcma")
    ok &= one(64, 45, 2, 3000, "mcma", odd=True)     # rows not 16-byte aligned: no TMA
    ok &= one(64, 45, 2, 3000, "mcma", adaptive=True, mu0=5e-3)
    ok &= one(64, 19, 2, 3000, "mrde", adaptive=True, niter=2, mu0=5e-3)
    ok &= one(16, 3, 2, 1000, "mcma")
    ok &= one(16, 7, 4, 1500, "mcma")
    ok &= one(64, 45, 2, 70000, "mcma", nseg=2)      # many ring wraps, re-basing of the running sums
    print("all ok" if ok else "FAILURES", "%.1f s" % (time.time() - t0))
    sys.exit(0 if ok else 1)
