import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import pipeline, synth
dev = torch.device('cuda', 0)
cfg0 = pipeline.ReceiverConfig(M=64, ntaps=45, os=2)
S = pipeline.balanced_segment_symbols(2 * 10**7, cfg0, target=8192)
cfg = pipeline.ReceiverConfig(M=64, ntaps=45, os=2, seg_symbols=S)
rx = pipeline.SegmentedReceiver(cfg, dev); rx.want_idx = False
E, _ = synth.synth_signal(64, 10**7, seed=1, snr_db=28.0, device=dev)
Eh = torch.empty(E.shape, dtype=E.dtype, pin_memory=True); Eh.copy_(E)
Ed = torch.empty_like(E)
for nch in (4, 8, 12, 16, 24, 32):
    outs = (None, None)
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = pipeline.run_host(rx, Eh, outs[0], outs[1], nchunks=nch, E_dev=Ed)[:2]
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print('chunks %2d: enqueue %.2f ms, total %.2f ms' % (nch, (t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
