"""BPS on ONE capture (2 streams of nrows): the three kernels (QB_BPS_KERNEL = default fast / ws / simple)."""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "run":
    sys.path.insert(0, '.')
    import numpy as np, torch
    from qampy_b200 import device, synth, theory
    dev = torch.device('cuda', 0)
    M, A, N, n = 64, 64, 45, 2 * 10 ** 6
    al = theory.normalised_symbols(M).astype(np.complex64)
    rng = np.random.default_rng(1)
    x = (al[rng.integers(0, M, (2, n))] * np.exp(1j * 0.1) + 0.03 * (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n)))).astype(np.complex64)
    xd = torch.from_numpy(x).to(dev)
    tabs = device.BpsTables(A, al, np.complex64, dev)
    for nstream in (2, 1):
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); device.bps(xd[:nstream], tabs, N, want_idx=False); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        t = min(ts)
        print("%-8s streams %d: %.1f ms = %.0f cycles/row" % (os.environ.get("QB_BPS_KERNEL", "fast"), nstream, t, t * 1e-3 * 1.965e9 / n), flush=True)
else:
    for k in ("", "ws", "simple"):
        env = dict(os.environ)
        if k: env["QB_BPS_KERNEL"] = k
        subprocess.run([sys.executable, __file__, "run"], env=env)
