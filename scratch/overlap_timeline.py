"""Do the training passes of one capture and the FIR + phase search of another share the machine?  Stream A (high
priority): mcma, mrde x K; stream B: apply, bps x K, no dependencies between them.  Times: each alone, both together."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import device, synth, theory
dev = torch.device('cuda', 0)
M, ntaps, S, K = 64, 45, int(sys.argv[1]) if len(sys.argv) > 1 else 8454, 10
E, _ = synth.synth_signal(M, 10 ** 7, seed=1, device=dev)
nseg = (E.shape[1] // 2 - 30) // S
Ev = device.segment_view(E, nseg, S, 2, ntaps)
tr = theory.cal_training_symbol_len(2, ntaps, Ev.shape[2])
sy = [torch.from_numpy(theory.reshape_symbols(None, m, M, np.complex64, 2)).to(dev) for m in ('mcma', 'mrde')]
w = torch.from_numpy(np.tile(theory.init_taps(ntaps, 2, np.complex64), (nseg, 1, 1, 1))).to(dev)
w2 = w.clone()
mu = torch.full((nseg, 2), 1e-3, dtype=torch.float32, device=dev)
err = torch.empty((nseg, 2, tr), dtype=torch.complex64, device=dev)
tables = device.BpsTables(64, theory.normalised_symbols(M).astype(np.complex64), np.complex64, dev)
eq = device.apply_filter_to_signal(Ev, 2, w2)
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else (0, -1)
sa, sb = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)

def train():
    for k in range(2):
        device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy[k], ('mcma', 'mrde')[k], err)

def tail():
    e = device.apply_filter_to_signal(Ev, 2, w2)
    device.bps(e.reshape(nseg * 2, -1), tables, 45, want_idx=False)

def run(do_a, do_b):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
    for i in range(K):
        if do_a:
            with torch.cuda.stream(sa): train()
        if do_b:
            with torch.cuda.stream(sb): tail()
    torch.cuda.current_stream().wait_stream(sa); torch.cuda.current_stream().wait_stream(sb)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

for r in range(2):
    a, b, ab = run(True, False), run(False, True), run(True, True)
    print("S %d: train alone %.3f ms, apply+bps alone %.3f ms, both streams %.3f ms per step (serial sum %.3f)" % (S, a, b, ab, a + b), flush=True)

# timeline of one overlapped round: events around every launch on both streams
def timeline():
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True); base.record()
    sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
    ev = []
    def mark(name, stream):
        e = torch.cuda.Event(enable_timing=True); e.record(stream); ev.append((name, e))
    for i in range(3):
        with torch.cuda.stream(sa):
            for k in range(2):
                mark("A%d %s start" % (i, ('mcma', 'mrde')[k]), sa)
                device.train_equaliser(Ev, tr, 1, 2, mu, w, None, False, sy[k], ('mcma', 'mrde')[k], err)
                mark("A%d %s end" % (i, ('mcma', 'mrde')[k]), sa)
        with torch.cuda.stream(sb):
            mark("B%d apply start" % i, sb)
            e = device.apply_filter_to_signal(Ev, 2, w2)
            mark("B%d apply end" % i, sb)
            device.bps(e.reshape(nseg * 2, -1), tables, 45, want_idx=False)
            mark("B%d bps end" % i, sb)
    torch.cuda.synchronize()
    for name, e in sorted(ev, key=lambda x: base.elapsed_time(x[1])):
        print("  %7.3f ms  %s" % (base.elapsed_time(e), name))
timeline()
