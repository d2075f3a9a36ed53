"""Per-stage kernel time of ONE capture (latency layout) for the script shapes: cycles per symbol by method / taps / step rule."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
for M, nsym, ntaps, cases in ((64, 2 ** 17, 13, (("mcma", True), ("mddma", True), ("mcma", False), ("mddma", False), ("sbd", False), ("dd", False))),
                              (16, 2 ** 18, 30, (("mcma", False), ("mrde", False))),
                              (32, 10 ** 6, 11, (("mcma", False), ("sbd", False))),
                              (64, 10 ** 6, 45, (("mcma", False), ("mrde", False), ("sbd", False)))):
    E, _ = synth.synth_signal(M, nsym, seed=3, snr_db=25.0, device=dev)
    tr = theory.cal_training_symbol_len(2, ntaps, E.shape[1])
    for dt in (torch.complex64, torch.complex128):
        Ed = E.to(dt)
        npdt = np.complex64 if dt == torch.complex64 else np.complex128
        for method, adaptive in cases:
            sy = torch.from_numpy(np.ascontiguousarray(theory.reshape_symbols(None, method, M, npdt, 2))).to(dev)
            ts = []
            for r in range(2):
                w = torch.from_numpy(theory.init_taps(ntaps, 2, npdt)[None]).to(dev)
                mu = torch.full((1, 2), 1e-3, dtype=torch.float32 if dt == torch.complex64 else torch.float64, device=dev)
                torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                device.train_equaliser(Ed[None], tr, 1, 2, mu, w, None, adaptive, sy, method, None, layout="latency")
                b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            print("M %3d ntaps %2d %-10s %-6s adaptive %-5s: %8.2f ms = %5.0f cycles/symbol" % (M, ntaps, str(dt)[6:], method, adaptive, min(ts), min(ts) * 1e-3 * 1.965e9 / tr), flush=True)
