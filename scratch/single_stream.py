"""Device time of ONE capture (reference call shape: 1 segment, 2 streams) per stage and training layout."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
def ev(f, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for name, M, nsym, ntaps, method, A, N in (("C2", 16, 10**6, 21, "mcma", 32, 21), ("C3", 64, 2 * 10**6, 45, "mcma", 64, 45), ("C3", 64, 2 * 10**6, 45, "mrde", 64, 45)):
    E, _ = synth.synth_signal(M, nsym, seed=3, snr_db=28.0, device=dev)
    Ev = E[None]
    tr = theory.cal_training_symbol_len(2, ntaps, E.shape[1])
    sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex64, 2)).to(dev)
    for layout in ("throughput", "latency"):
        def f():
            w = torch.from_numpy(theory.init_taps(ntaps, 2, np.complex64)[None]).to(dev)
            mu = torch.full((1, 2), 1e-3, dtype=torch.float32, device=dev)
            device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy, method, None, layout=layout)
            f.w = w
        t = ev(f)
        print("%s %s ntaps %d %s: %.1f ms = %.0f cycles/symbol" % (name, method, ntaps, layout, t, t * 1e-3 * 1.965e9 / tr), flush=True)
    t = ev(lambda: device.apply_filter_to_signal(Ev, 2, f.w))
    print("   apply %.2f ms" % t)
    eq = device.apply_filter_to_signal(Ev, 2, f.w)[0]
    tabs = device.BpsTables(A, theory.normalised_symbols(M).astype(np.complex64), np.complex64, dev)
    t = ev(lambda: device.bps(eq, tabs, N, want_idx=False))
    print("   bps (2 streams) %.1f ms = %.0f cycles/row" % (t, t * 1e-3 * 1.965e9 / eq.shape[1]), flush=True)
