import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import cpu_oracle as co
from qampy_b200 import device, pipeline, synth, theory
dev = torch.device('cuda', 0)
rms = lambda a: float(np.sqrt(np.mean(np.abs(a) ** 2)))
M, ntaps, S = 64, 45, 8192
E, syms = synth.synth_numpy(M, 30000, seed=7, snr_db=28.0)
Ed = torch.from_numpy(E).to(dev)
nseg = 3
Ev = device.segment_view(Ed, nseg, S, 2, ntaps)
L_seg = Ev.shape[2]
tr = theory.cal_training_symbol_len(2, ntaps, L_seg)
Es = np.stack([E[:, s * S * 2: s * S * 2 + L_seg] for s in range(nseg)])
for kern in ('fast', 'warp'):
    if kern == 'warp': os.environ['QB_TRAIN_KERNEL'] = 'warp'
    w = torch.from_numpy(np.tile(theory.init_taps(ntaps, 2, np.complex64), (nseg, 1, 1, 1))).to(dev)
    wr = np.tile(theory.init_taps(ntaps, 2, np.complex64), (nseg, 1, 1, 1))
    for method in ('mcma', 'mrde'):
        sy = theory.reshape_symbols(None, method, M, np.complex64, 2)
        mu = torch.full((nseg, 2), 1e-3, dtype=torch.float32, device=dev)
        err = torch.zeros((nseg, 2, tr), dtype=torch.complex64, device=dev)
        device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, torch.from_numpy(sy).to(dev), method, err)
        er, wr, _ = co.train_segments(Es, tr, 1, 2, 1e-3, wr, [0, 1], False, sy, method, mu_shared=False)
        e = err.cpu().numpy()
        print(kern, method, 'taps maxdiff per seg', [float(np.max(np.abs(w[s].cpu().numpy() - wr[s]))) for s in range(nseg)],
              'err rms', [rms(e[s] - er[s]) for s in range(nseg)])
        d = np.abs(e - er)
        bad = np.argwhere(d > 1e-4)
        print('   first bad err idx', bad[:5].tolist(), 'count', len(bad))
    out = device.apply_filter_to_signal(Ev, 2, torch.from_numpy(wr).to(dev)).cpu().numpy()
    ref = co.apply_segments(Es, 2, wr)
    d = np.abs(out - ref)
    print(kern, 'apply rms', rms(out - ref), 'bad idx', np.argwhere(d > 1e-4)[:10].tolist(), 'count', int((d > 1e-4).sum()))
