"""cProfile of pilots.pilot_receiver at BASELINE config C4 (capture resident on the GPU)."""
import sys, time, cProfile, pstats, warnings
sys.path.insert(0, '.'); warnings.filterwarnings("ignore")
import numpy as np, torch
from qampy_b200 import pilots, synth
dev = torch.device('cuda', 0)
M, fl, sl, rat, nfr = 256, 2 ** 16, 2 ** 10, 32, 61
d = synth.synth_pilot_signal(M, fl, sl, rat, nfr, snr_db=35, freq_off=100e6, linewidth=100e3, delay=4000, seed=11, device=dev)
seq, php, idx_pil = d["pilot_seq"].cpu().numpy(), d["ph_pilots"].cpu().numpy(), d["idx_pil"]
idx = np.nonzero(idx_pil)[0][sl:]
frames = list(range(59))
Ed = d["E"]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pilots.pilot_receiver(Ed, seq, php[:, :idx.size], idx_pil, fl, 2, frames, to_host=False)
    torch.cuda.synchronize(); print('run %.3f s' % (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
res = pilots.pilot_receiver(Ed, seq, php[:, :idx.size], idx_pil, fl, 2, frames, to_host=False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
# per-call timing of the batched trainings
import qampy_b200.equalisation as be
orig = be.equalise_windows
def timed(E, starts, window, os, mu, M, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(E, starts, window, os, mu, M, **kw)
    torch.cuda.synchronize()
    print('equalise_windows nwin %3d window %5d method %-4s Niter %2d adaptive %s symbols %s modes %s: %.1f ms' % (
        np.size(starts), window, kw.get('method'), kw.get('Niter', 1), kw.get('adaptive_stepsize'), 'yes' if kw.get('symbols') is not None else 'no', kw.get('modes'), 1e3 * (time.perf_counter() - t0)))
    return r
be.equalise_windows = timed
res = pilots.pilot_receiver(Ed, seq, php[:, :idx.size], idx_pil, fl, 2, frames, to_host=False)
