"""Static filter alone at the C3 shape (1184 segments of 8454 symbols) -- for ncu and quick timing."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps, S, nseg = 64, 45, 8454, 1184
E, _ = synth.synth_signal(M, nseg * S + 100, seed=1, device=dev)
Ev = device.segment_view(E, nseg, S, 2, ntaps)
rng = np.random.default_rng(0)
w = torch.from_numpy(((rng.standard_normal((nseg, 2, 2, ntaps)) + 1j * rng.standard_normal((nseg, 2, 2, ntaps))) / ntaps).astype(np.complex64)).to(dev)
ts = []
for r in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = device.apply_filter_to_signal(Ev, 2, w); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('apply ms', ['%.3f' % t for t in ts], flush=True)
