"""Training kernel (throughput layout) at the C3 capture cut into shorter segments: more streams -> more resident warps
per SM sub-partition.  Timing (no profiler) per segment length."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps = 64, 45
E, _ = synth.synth_signal(M, 10 ** 7, seed=1, device=dev)
for S in [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "8454,4227,2818,2113").split(',')]:
    nseg = (E.shape[1] // 2 - 30) // S
    Ev = device.segment_view(E, nseg, S, 2, ntaps)
    tr = theory.cal_training_symbol_len(2, ntaps, Ev.shape[2])
    for method in ('mcma', 'mrde'):
        sy = torch.from_numpy(theory.reshape_symbols(None, method, M, np.complex64, 2)).to(dev)
        w0 = torch.from_numpy(np.tile(theory.init_taps(ntaps, 2, np.complex64), (nseg, 1, 1, 1))).to(dev)
        mu = torch.full((nseg, 2), 1e-3, dtype=torch.float32, device=dev)
        ts = []
        for r in range(3):
            w = w0.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy, method, None)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("S %5d streams %5d warps/SMSP %.2f %s: %.3f ms" % (S, nseg * 2, nseg * 2 / 4 / 592, method, min(ts)), flush=True)
