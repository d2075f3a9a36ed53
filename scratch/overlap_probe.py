"""Does capture k's BPS overlap capture k+1's training?  Two streams, K captures back to back (C3 shape)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import pipeline, synth, device, theory
dev = torch.device('cuda', 0)
cfg0 = pipeline.ReceiverConfig(M=64, ntaps=45, os=2)
S = pipeline.balanced_segment_symbols(2 * 10**7, cfg0, target=8192)
cfg = pipeline.ReceiverConfig(M=64, ntaps=45, os=2, seg_symbols=S)
rx = pipeline.SegmentedReceiver(cfg, dev); rx.want_idx = False
E, _ = synth.synth_signal(64, 10**7, seed=1, snr_db=28.0, device=dev)
K = 8
def serial():
    for _ in range(K): rx.run(E)
def timed(f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
print('serial  %.3f ms per capture' % timed(serial))
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
groups = pipeline.plan_segments(E.shape[1], cfg)
first, nsym, nseg, drop = groups[0]
Ev = device.segment_view(E[:, first * 2:], nseg, nsym, 2, 45)
trs = theory.cal_training_symbol_len(2, 45, Ev.shape[2])
def overlapped():
    main = torch.cuda.current_stream()
    sA.wait_stream(main); sB.wait_stream(main)
    keep = []
    for _ in range(K):
        with torch.cuda.stream(sA):
            w = rx.w0.unsqueeze(0).repeat(nseg, 1, 1, 1)
            for st in range(2):
                mu = torch.full((nseg, 2), 1e-3, dtype=torch.float32, device=dev)
                device.train_equaliser(Ev, trs, 1, 2, mu, w, None, False, rx.syms[st], cfg.methods[st], None)
            eq = device.apply_filter_to_signal(Ev, 2, w)
            ev = torch.cuda.Event(); ev.record()
        with torch.cuda.stream(sB):
            sB.wait_event(ev)
            eq.record_stream(sB)
            out = device.bps(eq.reshape(nseg * 2, nsym), rx.bps_tables, 45, want_idx=False)
        keep.append((w, eq, out))
    main.wait_stream(sA); main.wait_stream(sB)
    return keep
print('overlap %.3f ms per capture' % timed(overlapped))
