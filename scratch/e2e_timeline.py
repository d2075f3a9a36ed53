"""Timeline of pipeline.run_host at C3: per chunk, when its H2D, chain and D2H finish (ms from start)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from qampy_b200 import pipeline, synth, device
dev = torch.device('cuda', 0)
cfg0 = pipeline.ReceiverConfig(M=64, ntaps=45, os=2)
S = pipeline.balanced_segment_symbols(2 * 10**7, cfg0, target=8192)
cfg = pipeline.ReceiverConfig(M=64, ntaps=45, os=2, seg_symbols=S)
rx = pipeline.SegmentedReceiver(cfg, dev); rx.want_idx = False
E, _ = synth.synth_signal(64, 10**7, seed=1, snr_db=28.0, device=dev)
Eh = torch.empty(E.shape, dtype=E.dtype, pin_memory=True); Eh.copy_(E)
Ed = torch.empty_like(E)
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
outs = (None, None)
for it in range(3):
    outs = pipeline.run_host(rx, Eh, outs[0], outs[1], nchunks=nch, E_dev=Ed)[:2]
    torch.cuda.synchronize()
# instrumented copy of run_host
st = rx._streams
groups = pipeline.plan_segments(E.shape[1], cfg)
main = torch.cuda.current_stream()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t0.record()
for s_ in [st["h2d"], st["d2h"]] + st["comp"]:
    s_.wait_stream(main)
ev = []
copied_to = 0
L = E.shape[1]
tc0 = time.perf_counter()
keep = []
for ci, (first, nsym, nseg, drop, seg0) in enumerate(pipeline._host_chunks(groups, nch)):
    need = min(L, (first + nsym * nseg) * cfg.os + cfg.ntaps - 1)
    e_in = torch.cuda.Event(enable_timing=True); e_c = torch.cuda.Event(enable_timing=True); e_o = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st["h2d"]):
        if need > copied_to:
            for k in range(2):
                Ed[k, copied_to:need].copy_(Eh[k, copied_to:need], non_blocking=True)
            copied_to = need
        e_in.record()
    comp = st["comp"][ci % len(st["comp"])]
    with torch.cuda.stream(comp):
        comp.wait_event(e_in)
        res = rx._run_group(Ed, first, nsym, nseg, drop, None, None)
        e_c.record()
    with torch.cuda.stream(st["d2h"]):
        st["d2h"].wait_event(e_c)
        outs[0][seg0:seg0 + nseg].copy_(res["out"], non_blocking=True)
        outs[1][seg0:seg0 + nseg].copy_(res["ph"], non_blocking=True)
        e_o.record()
    keep.append(res)
    ev.append((e_in, e_c, e_o, (time.perf_counter() - tc0) * 1e3))
torch.cuda.synchronize()
for ci, (a, b, c, tq) in enumerate(ev):
    print('chunk %2d  enqueued %.2f  h2d done %.2f  chain done %.2f  d2h done %.2f' % (ci, tq, t0.elapsed_time(a), t0.elapsed_time(b), t0.elapsed_time(c)))
