"""Training kernel alone at the C3 shape (1184 segments of 8454 symbols, 2 modes) -- for ncu and quick timing."""
import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps, S = 64, 45, 8454
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
E, _ = synth.synth_signal(M, nseg * S + 100, seed=1, device=dev)
Ev = device.segment_view(E, nseg, S, 2, ntaps)
tr = theory.cal_training_symbol_len(2, ntaps, Ev.shape[2])
for method in (sys.argv[2].split(',') if len(sys.argv) > 2 else ('mcma', 'mrde')):
    sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex64, 2)).to(dev)
    w0 = torch.from_numpy(np.tile(theory.init_taps(ntaps, 2, np.complex64), (nseg, 1, 1, 1))).to(dev)
    mu = torch.full((nseg, 2), 1e-3, dtype=torch.float32, device=dev)
    ts = []
    for r in range(3):
        w = w0.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy, method, None)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(method, 'streams', nseg * 2, '%.3f ms' % t, '%.0f cyc/sym/warp' % (t * 1e-3 * 1.965e9 / tr), flush=True)
