"""Convergence study for the headline recipe (CPU, fast oracle): SER of the equalised output (final taps applied to
a held-out stretch) against training length, step size and warm start.  python scratch/conv_study.py"""
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
sys.path.insert(0, '.')
sys.path.insert(0, 'oracle')
import numpy as np

import cpu_oracle as co
from qampy_b200 import synth, theory

KIND = "fast_native"
M, NT = 64, 45
NSYM = int(sys.argv[1]) if len(sys.argv) > 1 else 1200000
E, syms = synth.synth_numpy(M, NSYM, seed=1000, snr_db=28.0)
s1 = theory.reshape_symbols(None, "mcma", M, np.complex64, 2)
s2 = theory.reshape_symbols(None, "mrde", M, np.complex64, 2)
TEST0, TESTN = NSYM - 200000, 190000      # held-out stretch for the SER of a tap set


def train(seg, w, mus=(1e-3, 1e-3), methods=("mcma", "mrde")):
    """seg: (nseg, 2, L) ; w: (nseg,2,2,NT) in place"""
    tr = theory.cal_training_symbol_len(2, NT, seg.shape[2])
    for mu, m in zip(mus, methods):
        co.train_segments(seg, tr, 1, 2, mu, w, [0, 1], False, s1 if m == "mcma" else s2, m, mu_shared=False, kind=KIND)
    return w


def ser_of_taps(w, n=TESTN):
    """w (2,2,NT) applied to the held-out stretch"""
    seg = E[None, :, TEST0 * 2: TEST0 * 2 + n * 2 + NT - 1]
    eq = co.apply_segments(seg, 2, w[None].copy(), None, kind=KIND)[0]
    return synth.ser(eq, syms[:, TEST0:TEST0 + n + 200], M)


def seg_of(first, S):
    return np.ascontiguousarray(E[None, :, first * 2: first * 2 + S * 2 + NT - 1])


def ser_own(w, first, S):
    """SER of a segment equalised with its own final taps (what the pipeline outputs)"""
    eq = co.apply_segments(seg_of(first, S), 2, w[None].copy(), None, kind=KIND)[0]
    return synth.ser(eq, syms[:, first:first + S + 200], M)


w0 = theory.init_taps(NT, 2, np.complex64)
print("== cold start, mu 1e-3/1e-3: SER(held-out) after training on S symbols")
for S in (8454, 32768, 65536, 131072, 262144, 524288):
    t0 = time.time()
    w = train(seg_of(0, S), w0[None].copy())[0]
    print("S %7d  ser %.2e   (%.1f s)" % (S, ser_of_taps(w), time.time() - t0), flush=True)

print("== cold start with larger acquisition step sizes")
for mus in ((4e-3, 1e-3), (4e-3, 4e-3), (8e-3, 2e-3), (2e-3, 2e-3)):
    for S in (16384, 32768, 65536, 131072):
        w = train(seg_of(0, S), w0[None].copy(), mus)[0]
        print("mu %s S %7d  ser %.2e" % (mus, S, ser_of_taps(w)), flush=True)

print("== warm start: acquisition (cold, A symbols) then short segments (S=8454, mcma->mrde from the acquired taps)")
for A, amus in ((65536, (1e-3, 1e-3)), (131072, (1e-3, 1e-3)), (262144, (1e-3, 1e-3)), (32768, (4e-3, 2e-3)), (65536, (4e-3, 2e-3))):
    wa = train(seg_of(0, A), w0[None].copy(), amus)[0]
    sa = ser_of_taps(wa)
    S = 8454
    firsts = np.linspace(0, NSYM - S - 300, 48).astype(int)
    for label, mus, methods in (("mcma+mrde", (1e-3, 1e-3), ("mcma", "mrde")), ("mrde only", (1e-3,), ("mrde",)),
                                ("mrde+mrde", (1e-3, 1e-3), ("mrde", "mrde"))):
        segs = np.concatenate([seg_of(f, S) for f in firsts])
        w = train(segs, np.tile(wa, (len(firsts), 1, 1, 1)), mus, methods)
        sers = [ser_own(w[i], firsts[i], S) for i in range(len(firsts))]
        print("A %6d mu %s acq-ser %.2e | %s: ser mean %.2e max %.2e" % (A, amus, sa, label, np.mean(sers), np.max(sers)), flush=True)
