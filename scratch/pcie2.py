"""What the host link gives the e2e path: 320 MB up in 20 MB pieces on one stream while 560 MB go down in 20 MB pieces
on 1 / 3 streams, with and without a kernel running beside the copies.  ms until each direction is done."""
import sys
sys.path.insert(0, '.')
import torch
dev = torch.device('cuda', 0)
MB = 1 << 20
up_h = torch.empty(320 * MB, dtype=torch.uint8, pin_memory=True); up_d = torch.empty(320 * MB, dtype=torch.uint8, device=dev)
dn_h = torch.empty(560 * MB, dtype=torch.uint8, pin_memory=True); dn_d = torch.empty(560 * MB, dtype=torch.uint8, device=dev)
big = torch.empty(256 * MB, dtype=torch.float32, device=dev)
su = torch.cuda.Stream(); sd = [torch.cuda.Stream() for _ in range(3)]; sk = torch.cuda.Stream()

def run(nd, piece, up=True, down=True, kernel=False, delay_down=0):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    for s in [su, sk] + sd: s.wait_stream(torch.cuda.current_stream())
    eu = ed = None
    if kernel:
        with torch.cuda.stream(sk):
            for _ in range(12): big.mul_(1.0001)
    if up:
        with torch.cuda.stream(su):
            for o in range(0, 320 * MB, piece * MB): up_d[o:o + piece * MB].copy_(up_h[o:o + piece * MB], non_blocking=True)
            eu = torch.cuda.Event(enable_timing=True); eu.record()
    eds = []
    if down:
        for k, o in enumerate(range(0, 560 * MB, piece * MB)):
            with torch.cuda.stream(sd[k % nd]):
                dn_h[o:o + piece * MB].copy_(dn_d[o:o + piece * MB], non_blocking=True)
        for s in sd[:nd]:
            e = torch.cuda.Event(enable_timing=True); e.record(s); eds.append(e)
    torch.cuda.synchronize()
    return (t0.elapsed_time(eu) if eu else 0.0), max([t0.elapsed_time(e) for e in eds] or [0.0])

for name, kw in (("up alone", dict(nd=1, piece=20, down=False)), ("down alone, 1 stream", dict(nd=1, piece=20, up=False)),
                 ("down alone, 3 streams", dict(nd=3, piece=20, up=False)),
                 ("both, down on 1 stream", dict(nd=1, piece=20)), ("both, down on 3 streams", dict(nd=3, piece=20)),
                 ("both, 1 stream, 5 MB pieces", dict(nd=1, piece=5)), ("both, 1 stream, 80 MB pieces", dict(nd=1, piece=80)),
                 ("both, 1 stream + kernel", dict(nd=1, piece=20, kernel=True)), ("both, 3 streams + kernel", dict(nd=3, piece=20, kernel=True))):
    r = [run(**kw) for _ in range(3)][-1]
    print("%-32s up done %6.2f ms  down done %6.2f ms" % (name, r[0], r[1]), flush=True)
