"""Where the time of ONE drop-in equaliser call goes (script shape 64_qam_equalisation.py, complex64 and complex128):
cProfile of the host side + the device time of the kernels (launch list by events is not needed: total - host)."""
import cProfile, pstats, io, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import equalisation as eq, synth
M, nsym, ntaps, mu, methods, adaptive = 64, 2 ** 17, 13, (0.19e-2, 0.19e-2), ("mcma", "mddma"), (True, True)
E64, _ = synth.synth_signal(M, nsym, seed=3, snr_db=25.0, beta=0.01, theta=np.pi / 3, dgd=30e-12, device='cuda')
for dt in (np.complex64, np.complex128):
    E = E64.cpu().numpy().astype(dt)
    f = lambda: eq.dual_mode_equalisation(E, 2, mu, M, Ntaps=ntaps, methods=methods, adaptive_stepsize=adaptive)
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("== %s: %.1f ms per call" % (np.dtype(dt).name, (t1 - t0) * 1e3))
    pr = cProfile.Profile(); pr.enable(); f(); torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(18); print(s.getvalue()[:3500])
