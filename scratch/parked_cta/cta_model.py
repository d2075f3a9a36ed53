"""NumPy model of the CTA-wide single-stream trainer (csrc/eq_train_cta.cuh): same data flow, fp32.

    y_i = X_i . S_n  +  sum_{j = s_n}^{i-1} c_j G[i, j]          (exact algebra: block-exact LMS)
    S_n = W_{s_n},  s_n = max(0, 8 (n - 3)),  n = i // 8          (taps as of 3 blocks ago)
    G[i, j] = sum_{k,t} x_k[i os + t] conj(x_k[j os + t])         (lags 1 .. 24 + i % 8 <= 31)

Checks the index rules (snapshot age, lag masks, running-sum Gram with re-basing) against the strict oracle.
    python scratch/cta_model.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

B = 8
NLAG = 31
f32 = np.float32
c64 = np.complex64


def gram_prefix(x, T, ntaps, rebase=512):
    """G[i, l-1] for l = 1..31 by per-lag running sums of lag products, like the Gram warps do."""
    nm, L = x.shape
    G = np.zeros((T + 64, NLAG), c64)
    for l in range(1, NLAG + 1):
        # p[m] = sum_k x_k[m + 2 l] conj(x_k[m]),  m >= 0
        n = L - 2 * l
        if n <= 0:
            continue
        p = (x[:, 2 * l:2 * l + n] * np.conj(x[:, :n])).sum(axis=0).astype(c64)
        # fp32 running sum with periodic re-basing (error bounded by the epoch length)
        P = np.zeros(n + 1, c64)
        acc = c64(0)
        out = np.empty(n + 1, c64)
        out[0] = 0
        # emulate sequential fp32 accumulate in epochs: cumsum in complex64 inside an epoch
        for e0 in range(0, n, 2 * rebase):
            seg = p[e0:e0 + 2 * rebase]
            cs = np.cumsum(seg, dtype=c64)
            out[e0 + 1:e0 + 1 + seg.size] = cs      # relative to epoch start
        # G[i, i-l] = sum_{t<ntaps} p[(i-l) os + t] = Pfull[(i-l)os + ntaps] - Pfull[(i-l)os]; differences are taken
        # inside an epoch (re-based ring), across an epoch boundary the kernel re-bases the ring: same value to rounding
        Pfull = np.zeros(n + 1, np.complex128)
        Pfull[1:] = np.cumsum(p.astype(np.complex128))
        for i in range(l, T + 64):
            m0 = (i - l) * 2
            if m0 + ntaps <= n:
                G[i, l - 1] = c64(Pfull[m0 + ntaps] - Pfull[m0])
    return G


def errfct(method, y, syms):
    if method == "mcma":
        R = syms[0]
        return c64((R.real - y.real * y.real) * y.real + 1j * ((R.imag - y.imag * y.imag) * y.imag))
    if method == "cma":
        R = syms[0].real
        return c64((R - (y.real * y.real + y.imag * y.imag)) * y)
    raise ValueError(method)


def train_cta_model(x, T, ntaps, mu, w0, mode, syms, method="mcma"):
    nm, L = x.shape
    x = x.astype(c64)
    G = gram_prefix(x, T, ntaps)
    W = w0[mode].astype(c64).copy()          # (nm, ntaps) current taps of the update warps
    nblk = (T + B - 1) // B
    snaps = {0: W.copy()}                    # snapshot v = taps W_{8 v}
    c = np.zeros(T + 64, c64)
    e = np.zeros(T, c64)
    mu = f32(mu)

    def window(i):
        lo = i * 2
        w = np.zeros((nm, ntaps), c64)
        hi = min(L, lo + ntaps)
        if hi > lo:
            w[:, :hi - lo] = x[:, lo:hi]
        return w

    for n in range(nblk):
        v = max(0, n - 3)
        S = snaps[v]
        s_n = 8 * v
        for b in range(B):
            i = n * B + b
            if i >= T:
                break
            base = c64((window(i) * S).sum())
            y = base
            for l in range(1, 24 + b + 1):          # allowed lags
                j = i - l
                if j < s_n or j < 0:
                    break
                y = c64(y + c[j] * G[i, l - 1])
            e[i] = errfct(method, y, syms)
            c[i] = c64(mu * e[i])
        # update warps: apply block n's steps, publish snapshot n + 1
        for b in range(B):
            i = n * B + b
            if i >= T:
                break
            W = (W + c[i] * np.conj(window(i))).astype(c64)
        snaps[n + 1] = W.copy()
        snaps.pop(n - 4, None)
    return W, e


def main():
    import cpu_oracle as co
    from qampy_b200 import synth, theory
    M, ntaps = 64, 45
    nsym = 6000
    E, _ = synth.synth_numpy(M, nsym, seed=3, snr_db=28.0)
    E = E.astype(c64)
    L = E.shape[1]
    T = theory.cal_training_symbol_len(2, ntaps, L)
    syms = theory.reshape_symbols(None, "mcma", M, c64, 2)
    w0 = theory.init_taps(ntaps, 2, c64)
    w_ref = w0.copy()
    err_ref, w_ref, _ = co.train_equaliser(E, T, 1, 2, 1e-3, w_ref, [0, 1], False, syms, "mcma")
    for mode in (0, 1):
        W, e = train_cta_model(E, T, ntaps, 1e-3, w0, mode, syms[mode], "mcma")
        dw = np.abs(W - w_ref[mode]).max()
        de = np.sqrt(np.mean(np.abs(e - err_ref[mode][:T]) ** 2))
        print("mode %d: max tap diff %.3e  err rms diff %.3e (err rms %.3e)" % (
            mode, dw, de, np.sqrt(np.mean(np.abs(e) ** 2))))


if __name__ == "__main__":
    main()
