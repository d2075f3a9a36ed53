"""BPS on ONE capture, device tensors: producer / chain split against the phase-parallel form (qb_set_option BPS_SPLIT)."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, theory
dev = torch.device('cuda', 0)
M, A, N = 64, 64, 45
al = theory.normalised_symbols(M).astype(np.complex64)
tabs = device.BpsTables(A, al, np.complex64, dev)
for n in (2 ** 17, 10 ** 6, 10 ** 7):
    rng = np.random.default_rng(1)
    x = (al[rng.integers(0, M, (2, n))] * np.exp(1j * 0.1) + 0.03 * (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n)))).astype(np.complex64)
    xd = torch.from_numpy(x).to(dev)
    ref = None
    for mode in ("1", "2"):
        device.set_option("BPS_SPLIT", mode)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out, ph, idx = device.bps(xd, tabs, N); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        if ref is None: ref = (ph.clone(), idx.clone())
        same = bool((ref[0] == ph).all()) and bool((ref[1] == idx).all())
        t = min(ts)
        print("rows %8d  %s: %8.2f ms = %6.1f cycles/row  identical %s" % (n, {"1": "producer/chain", "2": "phase-parallel"}[mode], t, t * 1e-3 * 1.965e9 / n, same), flush=True)
    device.set_option("BPS_SPLIT", None)
