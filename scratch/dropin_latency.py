"""Single-capture (reference semantics, ONE segment) timing through the NumPy drop-in API vs the oracle port
built with the reference's flags.  C2: 16-QAM MCMA ntaps 21, 1e6 symbols, BPS(32, 21); C3 strict: 64-QAM
MCMA->MRDE ntaps 45, nsym symbols, BPS(64, 45)."""
import os, sys, time
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import cpu_oracle as co
from qampy_b200 import equalisation as eq, phaserecovery as ph, synth, theory

def tm(f, n=2):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r

for name, M, nsym, ntaps, methods, A, N in (("C2", 16, 10**6, 21, ("mcma",), 32, 21), ("C3-2e6", 64, 2 * 10**6, 45, ("mcma", "mrde"), 64, 45)):
    E, _ = synth.synth_signal(M, nsym, seed=3, snr_db=28.0, device='cuda')
    E = E.cpu().numpy()
    alphabet = theory.normalised_symbols(M).astype(np.complex64)
    if len(methods) == 1:
        g = lambda: eq.equalise_signal(E, 2, 1e-3, M, Ntaps=ntaps, method=methods[0], apply=True)
        c = lambda: co.equalise_signal(E, 2, 1e-3, M, Ntaps=ntaps, method=methods[0], apply=True, kind="fast_native")
    else:
        g = lambda: eq.dual_mode_equalisation(E, 2, (1e-3, 1e-3), M, Ntaps=ntaps, methods=methods)
        c = lambda: co.dual_mode_equalisation(E, 2, (1e-3, 1e-3), M, Ntaps=ntaps, methods=methods, kind="fast_native")
    tg, rg = tm(g)
    tc, rc = (0.0, rg) if os.environ.get('SKIP_CPU') else tm(c, 1)
    Eo = rg[0]
    tgb, _ = tm(lambda: ph.bps(Eo, A, alphabet, N))
    tcb = 0.0 if os.environ.get('SKIP_CPU') else tm(lambda: co.bps_driver(Eo, A, alphabet, N, kind="fast_native"), 1)[0]
    d = float(np.sqrt(np.mean(np.abs(rg[0] - rc[0]) ** 2)))
    np.save('gpurun_out/lat_%s_%s.npy' % (name, os.environ.get('QB_TRAIN_LA_LPS', '8')), rg[0][:, :200000])
    print("%s nsym %d: equaliser GPU %.3f s CPU %.3f s (rms diff %.1e) | bps GPU %.3f s CPU %.3f s" % (name, nsym, tg, tc, d, tgb, tcb), flush=True)
