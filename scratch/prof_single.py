"""ONE capture, latency layout (one stream per warp), short: for ncu source-level captures of train_la_kernel<32,...>."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps, nsym = 64, 45, int(sys.argv[1]) if len(sys.argv) > 1 else 60000
E, _ = synth.synth_signal(M, nsym, seed=3, snr_db=28.0, device=dev)
tr = theory.cal_training_symbol_len(2, ntaps, E.shape[1])
for method in ("mcma", "mrde"):
    sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex64, 2)).to(dev)
    for r in range(2):
        w = torch.from_numpy(theory.init_taps(ntaps, 2, np.complex64)[None]).to(dev)
        mu = torch.full((1, 2), 1e-3, dtype=torch.float32, device=dev)
        device.train_equaliser(E[None], tr, 1, 2, mu, w, None, False, sy, method, None, layout="latency")
    torch.cuda.synchronize()
print("ok")
