import sys, os, time
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps, S = 64, 45, 8192
E, _ = synth.synth_signal(M, 1220 * S + 100, seed=1, device=dev)
def time_train(nseg, method, reps=3):
    Ev = device.segment_view(E, nseg, S, 2, ntaps)
    tr = theory.cal_training_symbol_len(2, ntaps, Ev.shape[2])
    sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex64, 2)).to(dev)
    w0 = torch.from_numpy(np.tile(theory.init_taps(ntaps, 2, np.complex64), (nseg, 1, 1, 1))).to(dev)
    mu = torch.full((nseg, 2), 1e-3, dtype=torch.float32, device=dev)
    ts = []
    for r in range(reps + 1):
        w = w0.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy, method, None)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), tr
for la in ('1', '0'):
    os.environ['QB_TRAIN_LA'] = la
    for method in ('mcma', 'mrde'):
        row = []
        for nseg in (1, 4, 37, 74, 148, 296, 592, 1220):
            t, tr = time_train(nseg, method)
            row.append('%d:%.2fms(%.0fcyc/sym)' % (nseg * 2, t, t * 1e-3 * 1.965e9 / tr))
        print('LA=' + la, method, ' '.join(row), flush=True)
# BPS alone
alphabet = theory.normalised_symbols(M).astype(np.complex64)
tables = device.BpsTables(64, alphabet, np.complex64, dev)
x = E[:, :1220 * S].reshape(2 * 1220, S)[:, :S].contiguous()
for ns in (2, 8, 148, 592, 1184, 2440):
    xs = x[:ns].contiguous()
    ts = []
    for r in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); device.bps(xs, tables, 45, want_idx=False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('bps streams', ns, '%.3f ms' % min(ts), '%.0f cyc/row' % (min(ts) * 1e-3 * 1.965e9 / S), flush=True)
