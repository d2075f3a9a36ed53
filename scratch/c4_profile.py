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
