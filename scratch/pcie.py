import torch, time
dev = torch.device('cuda', 0)
n = 320_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(240_000_000, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(240_000_000, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); print('H2D 320MB %.2f ms %.1f GB/s' % (a * 1e3, n / a / 1e9))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print('D2H 240MB %.2f ms %.1f GB/s' % (b * 1e3, 240e6 / b / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print('both %.2f ms' % (c * 1e3))
# chunked H2D: 12 chunks
def chunked():
    k = n // 12
    for i in range(12): d[i*k:(i+1)*k].copy_(h[i*k:(i+1)*k], non_blocking=True)
e = t(chunked); print('H2D 12 chunks %.2f ms' % (e * 1e3))
