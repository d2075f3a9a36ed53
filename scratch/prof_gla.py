"""ONE capture, complex128, mcma (the mrde_equaliser.py shape): for ncu source-level captures of train_gla_kernel<double,...>."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps, nsym = 16, 30, 60000
E, _ = synth.synth_signal(M, nsym, seed=3, snr_db=25.0, device=dev, dtype=torch.complex128)
tr = theory.cal_training_symbol_len(2, ntaps, E.shape[1])
for method in ("mcma", "mrde"):
    sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex128, 2)).to(dev)
    for r in range(2):
        w = torch.from_numpy(theory.init_taps(ntaps, 2, np.complex128)[None]).to(dev)
        mu = torch.full((1, 2), 1e-3, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        device.train_equaliser(E[None], tr, 1, 2, mu, w, None, False, sy, method, None, layout="latency")
        b.record(); torch.cuda.synchronize()
    print(method, "%.2f ms = %.0f cycles/symbol" % (a.elapsed_time(b), a.elapsed_time(b) * 1e-3 * 1.965e9 / tr))
