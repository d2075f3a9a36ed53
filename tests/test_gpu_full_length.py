"""Full-length ONE-capture parity (BASELINE configs C2 and C3 in the reference's own call shape: one stream per trained
mode, 1e6 / 1e7 symbols deep) against the strict oracle -- what the short-segment tests cannot show: the drift of an
fp32 look-ahead recurrence over 1e7 sequential updates, and bit-exactness of the blind phase search's sequential fp32
running sum where it has grown to 2e5 and crossed many binades.

Everything goes through the public drop-in functions (qampy_b200.equalisation.* / phaserecovery.bps)."""
import os
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2)))


def _capture(M, nsym, seed):
    import torch
    from qampy_b200 import synth
    # synthesised on the device (a 1e7-symbol double-precision FFT is slow on the host), used from the host like a user
    E, syms = synth.synth_signal(M, nsym, seed=seed, snr_db=28.0, device=torch.device("cuda", 0))
    return E.cpu().numpy(), syms.cpu().numpy()


def test_c2_one_capture_1e6_symbols_equalise_signal_and_bps():
    """C2: dual-pol 16-QAM, MCMA ntaps 21, 1e6 symbols, then BPS(32, N = 21)."""
    import cpu_oracle as co
    from qampy_b200 import equalisation, phaserecovery, synth, theory
    M, ntaps, nsym = 16, 21, 10 ** 6
    E, syms = _capture(M, nsym, 21)
    Eo, w, err = equalisation.equalise_signal(E, 2, 1e-3, M, Ntaps=ntaps, method="mcma", apply=True)
    Er, wr, er = co.equalise_signal(E, 2, 1e-3, M, Ntaps=ntaps, method="mcma", apply=True)
    assert Eo.shape == Er.shape == (2, (E.shape[1] - ntaps + 1) // 2) and err.shape == er.shape
    assert rms(Eo - Er) < 1e-5 and np.max(np.abs(w - wr)) < 1e-5
    assert rms(err - er) < 1e-5 and rms(err[:, -10000:] - er[:, -10000:]) < 1e-5      # no drift towards the end
    al = theory.normalised_symbols(M).astype(np.complex64)
    Eb, ph = phaserecovery.bps(Er, 32, al, 21)
    Ebr, phr = co.bps_driver(Er, 32, al, 21)
    assert np.array_equal(ph, phr) and rms(Eb - Ebr) < 1e-6
    assert synth.ser(Eb[:, 1000:200000], syms[:, 1000:200300], M) < 1e-3         # MCMA alone on 16-QAM at 28 dB


def test_c3_one_capture_1e7_symbols_dual_mode_and_bps():
    """C3: dual-pol 64-QAM, MCMA -> MRDE ntaps 45, 1e7 symbols, then BPS(64, N = 45) with bit-exact indices over
    the whole 1e7-row running sum."""
    import cpu_oracle as co
    from qampy_b200 import equalisation, phaserecovery, pythran_dsp, synth, theory
    M, ntaps, nsym = 64, 45, int(os.environ.get("QB_FULL_NSYM", 10 ** 7))
    E, syms = _capture(M, nsym, 33)
    Eo, w, (e1, e2) = equalisation.dual_mode_equalisation(E, 2, (1e-3, 1e-3), M, Ntaps=ntaps, methods=("mcma", "mrde"))
    Er, wr, (r1, r2) = co.dual_mode_equalisation(E, 2, (1e-3, 1e-3), M, Ntaps=ntaps, methods=("mcma", "mrde"))
    assert Eo.shape == Er.shape == (2, (E.shape[1] - ntaps + 1) // 2)
    assert rms(Eo - Er) < 1e-5 and np.max(np.abs(w - wr)) < 1e-5
    for a, b in ((e1, r1), (e2, r2)):
        assert a.shape == b.shape and rms(a - b) < 1e-5
        assert rms(a[:, -100000:] - b[:, -100000:]) < 1e-5                           # the last 1e5 of 1e7 updates
    al = theory.normalised_symbols(M).astype(np.complex64)
    ang = theory.bps_test_angles(64, np.float32)
    # the L1 kernel's output: indices, bit exact on the array the oracle saw; then the L2 driver's phases
    for r in range(2):
        idx = pythran_dsp.bps(Er[r], ang, al, 45)
        idr = co.bps_streams(Er[r][None], ang, al, 45)[0]
        assert idx.shape == idr.shape == (Er.shape[1],) and np.array_equal(idx, idr)
    Eb, ph = phaserecovery.bps(Er, 64, al, 45)
    Ebr, phr = co.bps_driver(Er, 64, al, 45)
    assert np.array_equal(ph, phr) and rms(Eb - Ebr) < 1e-6
    # converged: the second half of the capture demodulates without errors to speak of
    h = Eb.shape[1] // 2
    assert synth.ser(Eb[:, h:h + 400000], syms[:, h:h + 400300], M) < 1e-5
