"""BPS alone at the C3 shape (2368 streams x 8454 rows, 64 angles, N = 45) -- for ncu and quick timing."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, S, NS = 64, 8454, int(sys.argv[1]) if len(sys.argv) > 1 else 2368
alphabet = theory.normalised_symbols(M).astype(np.complex64)
rng = np.random.default_rng(0)
x = alphabet[rng.integers(0, M, (NS, S))] + 0.03 * (rng.standard_normal((NS, S)) + 1j * rng.standard_normal((NS, S)))
x = torch.from_numpy(x.astype(np.complex64)).to(dev)
tables = device.BpsTables(64, alphabet, np.complex64, dev)
ts = []
for r in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); device.bps(x, tables, 45, want_idx=False); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('bps streams', NS, 'ms', ['%.3f' % t for t in ts], '%.1f cyc/row/SM-resident' % (min(ts) * 1e-3 * 1.965e9 / S), flush=True)
