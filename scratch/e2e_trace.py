"""Timeline of pipeline.run_host at the bench's C3 recipe (warm start, halo, errors): when every upload, training
stage, chain and download ends, ms from the start of the call.  argv: chunks [segment symbols]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import pipeline, synth
dev = torch.device('cuda', 0)
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg0 = pipeline.ReceiverConfig(M=64, ntaps=45, os=2)
S = int(sys.argv[2]) if len(sys.argv) > 2 else pipeline.balanced_segment_symbols(2 * 10 ** 7, cfg0, target=8192)
cfg = pipeline.ReceiverConfig(M=64, ntaps=45, os=2, seg_symbols=S, want_err=True, bps_halo=45)
rx = pipeline.SegmentedReceiver(cfg, dev); rx.want_idx = False
E, _ = synth.synth_signal(64, 10 ** 7, seed=1, snr_db=28.0, device=dev)
taps = rx.acquire(E)
Eh = torch.empty(E.shape, dtype=E.dtype, pin_memory=True); Eh.copy_(E)
Ed = torch.empty_like(E)
o = p = e = None
for it in range(4):
    rx.trace = [] if it == 3 else None
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    o, p, _ = pipeline.run_host(rx, Eh, o, p, nchunks=nch, E_dev=Ed, wxy0=taps, err_host=e)
    e = rx.err_host
    t1.record(); torch.cuda.synchronize()
    print("run %d: %.2f ms (S %d, %d chunks)" % (it, t0.elapsed_time(t1), S, nch))
base = rx.trace[0][1]
for label, ev in sorted(rx.trace[1:], key=lambda x: base.elapsed_time(x[1])):
    print("  %7.2f  %s" % (base.elapsed_time(ev), label))
