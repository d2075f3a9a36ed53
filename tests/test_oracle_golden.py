"""Pins the CPU oracle (oracle/) against vectors produced by executing the reference's own
Python sources (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import cpu_oracle as co


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2))) if np.size(a) else 0.0


def test_constants(golden):
    g = golden("g0_constants")
    for M in (4, 16, 32, 64, 128, 256):
        assert np.array_equal(co.cal_symbols_qam(M), g["syms_%d" % M])
        assert co.cal_scaling_factor_qam(M) == pytest.approx(float(g["scale_%d" % M]), rel=1e-15)
        for m in ("cma", "cma2", "sgncma", "mcma", "rde", "mrde", "sbd", "mddma", "dd"):
            mine = co.generate_symbols_for_eq(m, M, np.complex128)
            ref = g["eqsyms_%s_%d" % (m, M)]
            assert mine.shape == ref.shape
            np.testing.assert_allclose(mine, ref, rtol=1e-14, atol=1e-15)
            # after the cast the kernels see, the constants are bit-identical
            assert np.array_equal(mine.astype(np.complex64), ref.astype(np.complex64))
    for L, nt, os_ in ((20000, 11, 2), (2000000, 21, 2), (20000000, 45, 2), (6000, 11, 2), (5000, 7, 1)):
        assert co.cal_training_symbol_len(os_, nt, L) == int(g["trsyms_%d_%d_%d" % (L, nt, os_)])


def test_c1_single_pol_cma(golden):
    g = golden("g1_c1_cma")
    E, wxy, err = co.equalise_signal(g["E_in"], 2, 1e-3, 4, Ntaps=11, method="cma", apply=True)
    assert E.shape == g["E_out"].shape and err.shape == g["err"].shape
    assert rms(E - g["E_out"]) < 1e-6
    assert rms(err - g["err"]) < 5e-6   # fp32 round-off: the interpreted reference mixes FMA array ops and powf/hypot scalars
    assert np.max(np.abs(wxy - g["wxy"])) < 1e-6


@pytest.mark.parametrize("tag,tol", [("c64", 2e-6), ("c128", 1e-12)])
def test_dual_mode_16qam(golden, tag, tol):
    g = golden("g2_dual16_" + tag)
    E, wxy, (e1, e2) = co.dual_mode_equalisation(g["E_in"], 2, (1e-3, 1e-3), 16, Ntaps=11,
                                                 methods=("mcma", "mrde"))
    assert E.dtype == g["E_out"].dtype
    assert rms(E - g["E_out"]) < tol
    assert rms(e1 - g["err1"]) < tol and rms(e2 - g["err2"]) < tol
    assert np.max(np.abs(wxy - g["wxy"])) < tol
    # BPS: indices bit exact, phases exact, rotated output to rounding
    A, N = int(g["bps_A"]), int(g["bps_N"])
    dt = np.float32 if tag == "c64" else np.float64
    ang = np.linspace(-np.pi / 4, np.pi / 4, A, endpoint=False, dtype=dt).reshape(1, -1)
    idx = co.bps_streams(g["bps_in"], ang, g["coded"], N)
    assert np.array_equal(idx, g["bps_idx"])
    Eb, ph = co.bps_driver(g["bps_in"], A, g["coded"], N)
    assert np.array_equal(ph, g["bps_ph"])
    assert rms(Eb - g["bps_out"]) < (1e-6 if tag == "c64" else 1e-14)


def test_dual_mode_64qam(golden):
    g = golden("g3_dual64_c64")
    E, wxy, (e1, e2) = co.dual_mode_equalisation(g["E_in"], 2, (1e-3, 1e-3), 64, Ntaps=15,
                                                 methods=("mcma", "mrde"))
    assert rms(E - g["E_out"]) < 2e-6
    assert rms(e1 - g["err1"]) < 2e-6 and rms(e2 - g["err2"]) < 2e-6
    ang = np.linspace(-np.pi / 4, np.pi / 4, 64, endpoint=False, dtype=np.float32).reshape(1, -1)
    idx = co.bps_streams(g["bps_in"], ang, g["coded"], 20)
    assert np.array_equal(idx, g["bps_idx"])
    assert len(np.unique(idx)) > 32          # the phase walk exercises most test angles
    Eb, ph = co.bps_driver(g["bps_in"], 64, g["coded"], 20)
    assert np.array_equal(ph, g["bps_ph"])
    assert rms(Eb - g["bps_out"]) < 1e-6


@pytest.mark.parametrize("method", ["cma", "cma2", "sgncma", "mcma", "rde", "mrde", "sbd", "mddma",
                                    "dd", "sbd_data"])
def test_all_error_functions(golden, method):
    g = golden("g4_methods")
    for tag, dt, tol in (("c64", np.complex64, 3e-6), ("c128", np.complex128, 1e-12)):
        if "wxy_%s_%s" % (method, tag) not in g:
            continue
        sy = g["symbols_tx"] if method == "sbd_data" else g["coded"]
        wxy, err = co.equalise_signal(g["E_in"].astype(dt), 2, 2e-3, 16, Ntaps=7, Niter=2,
                                      method=method, symbols=sy.astype(dt))
        ref_w, ref_e = g["wxy_%s_%s" % (method, tag)], g["err_%s_%s" % (method, tag)]
        assert err.shape == ref_e.shape
        if method == "cma2":
            # cma2 is unstable on this input IN THE REFERENCE (non-finite from symbol ~200 on); parity is
            # only meaningful before the blow-up, and both must blow up
            assert not np.isfinite(ref_e).all() and not np.isfinite(err).all()
            assert rms(err[:, :100] - ref_e[:, :100]) < 1e-4 * rms(ref_e[:, :100])
            continue
        assert rms(err - ref_e) < tol * max(1.0, rms(ref_e)), method
        assert np.max(np.abs(wxy - ref_w)) < tol * 5, method


def test_adaptive_stepsize(golden):
    g = golden("g4_methods")
    sy = co._reshape_symbols(None, "mcma", 16, np.complex64, 2)
    # one selected mode: deterministic in the reference
    w0 = co.init_taps(7, 2, np.complex64)
    err, w, mu = co.train_equaliser(g["E_in"].copy(), 700, 2, 2, np.float32(1e-2), w0, np.array([1]),
                                    True, sy, "mcma")
    assert rms(err - g["ad_err_m1"]) < 3e-6
    assert np.max(np.abs(w - g["ad_wxy_m1"])) < 1e-5
    assert mu == pytest.approx(float(g["ad_mu_m1"]), rel=1e-4)
    # both modes: the interpreted reference carries mu from mode 0 into mode 1 (mu_shared)
    w0 = co.init_taps(7, 2, np.complex64)
    err, w, mu = co.train_equaliser(g["E_in"].copy(), 700, 2, 2, np.float32(1e-2), w0, np.array([0, 1]),
                                    True, sy, "mcma", mu_shared=True)
    assert rms(err - g["ad_err_m01"]) < 3e-6
    assert np.max(np.abs(w - g["ad_wxy_m01"])) < 1e-5
    assert mu == pytest.approx(float(g["ad_mu_m01"]), rel=1e-4)


@pytest.mark.parametrize("tag", list("abcde"))
def test_apply_filter_shapes(golden, tag):
    g = golden("g5_apply")
    modes = g["modes_" + tag]
    modes = None if modes[0] < 0 else modes
    out = co.apply_filter_to_signal(g["E_" + tag], int(g["os_" + tag]), g["w_" + tag], modes)
    assert out.shape == g["out_" + tag].shape
    assert rms(out - g["out_" + tag]) < 1e-6


def test_bps_known_answer(golden):
    g = golden("g6_bps_kat")
    for k in range(3):
        Eb, ph = co.bps_driver(g["in_%d" % k], 32, g["coded"], 11)
        assert np.array_equal(ph, g["ph_%d" % k])
        assert rms(Eb - g["out_%d" % k]) < 1e-6
        # test/test_phaserec.py:124-145: recovered phase = -angle within one test-angle step
        np.testing.assert_allclose(ph[0, 20:-20] + g["angle_%d" % k], 0, atol=np.pi / 4 / 32)
    Eb, ph = co.bps_driver(g["in_1d"], 16, g["coded_1d"], 8)
    assert Eb.ndim == 1 and ph.ndim == 1
    assert np.array_equal(ph, g["ph_1d"])
    assert rms(Eb - g["out_1d"]) < 1e-14


@pytest.mark.parametrize("tag", ["c64", "c128"])
def test_bps_twostage(golden, tag):
    """Two-stage BPS (phaserecovery.py:222-288): the per-symbol angle table form (p == L) of the index
    search is bit exact, and so is everything built on it."""
    g = golden("g7_bps_twostage")
    E, coded = g["in_" + tag], g["coded_" + tag]
    A, N, B = int(g["A_" + tag]), int(g["N_" + tag]), int(g["B_" + tag])
    idx2 = co.bps(E[0], g["phn_" + tag], coded, N)
    assert np.array_equal(idx2, g["idx2_" + tag])
    assert np.array_equal(co.select_angles(g["phn_" + tag], idx2), g["phf_" + tag])
    En, ph = co.bps_twostage_driver(E, A, coded, N, B=B)
    assert ph.dtype == g["ph_" + tag].dtype and np.array_equal(ph, g["ph_" + tag])
    assert rms(En - g["out_" + tag]) < (1e-6 if tag == "c64" else 1e-13)


REAL_CASES = [("cma_real", False), ("sgncma_real", False), ("dd_real", False), ("dd_data_real", False),
              ("cma_real", True), ("dd_real", True)]


@pytest.mark.parametrize("tag,tol", [("c64", 5e-6), ("c128", 1e-12)])
@pytest.mark.parametrize("method,adaptive", REAL_CASES)
def test_real_valued_methods(golden, tag, tol, method, adaptive):
    """equalise_signal with the REAL_VALUED methods (equalisation.py:529-565): the NumPy restatement of
    train_equaliser_realvalued against what the reference produced."""
    g = golden("g8_real_valued")
    E, M = g["E_" + tag], int(g["M_" + tag])
    key = "%s_%s%s" % (tag, method, "_ad" if adaptive else "")
    sy = g["tx_" + tag] if method == "dd_data_real" else None
    Eo, wxy, err = co.equalise_signal_real(E, 2, 2e-3, M, 7, method, adaptive=adaptive, symbols=sy)
    assert wxy.dtype == g["wxy_" + key].dtype and err.dtype == g["err_" + key].dtype
    assert Eo.dtype == g["out_" + key].dtype
    assert rms(Eo - g["out_" + key]) < tol
    assert np.max(np.abs(wxy - g["wxy_" + key])) < tol
    if method != "sgncma_real":                       # sign() of a value at rounding distance from 0 may flip
        assert rms(err - g["err_" + key]) < tol


@pytest.mark.parametrize("tag", ["c64", "c128"])
def test_decisions_and_metrics(golden, tag):
    """make_decision, the soft demappers and estimate_snr: NumPy restatements vs the reference."""
    g = golden("g10_decisions")
    rx, coded = g["rx_" + tag], g["coded_" + tag]
    det, dist, idx = co.make_decision(rx, coded)
    assert np.array_equal(idx, g["idx_" + tag]) and np.array_equal(det, g["det_" + tag])
    assert np.allclose(dist, g["dist_" + tag], rtol=1e-6, atol=0)
    nb, snr, bm = int(g["nbits_" + tag]), g["snr_" + tag], g["bitmap_" + tag]
    tol = 2e-4 if tag == "c64" else 1e-9
    assert np.max(np.abs(co.soft_l_value_demapper(rx, nb, snr, bm) - g["lv_" + tag])) < tol * max(1.0, np.max(np.abs(g["lv_" + tag])))
    assert np.max(np.abs(co.soft_l_value_demapper(rx, nb, snr, bm, minmax=True) - g["lvmm_" + tag])) < tol * max(1.0, np.max(np.abs(g["lvmm_" + tag])))
    assert np.allclose(co.estimate_snr(rx, g["tx_" + tag], coded), g["est_" + tag], rtol=1e-5 if tag == "c64" else 1e-12)


def test_viterbiviterbi(golden):
    """Viterbi-Viterbi (phaserecovery.py:40-79): the NumPy restatement follows the reference's dtype at every step,
    so it reproduces the reference's phases to rounding."""
    g = golden("g11_viterbi")
    E = g["in_c64"]
    for N in (10, 11):
        Eo, ph = co.viterbiviterbi(E, N, 4)
        assert ph.dtype == np.float32 and ph.shape == (2, E.shape[1] - N + 1)
        assert np.max(np.abs(ph[1] - g["ph_c64_%d" % N])) < 1e-6 and np.max(np.abs(ph[0] - g["ph0_c64_%d" % N])) < 1e-6
        assert Eo.dtype == np.complex64 and rms(Eo - g["out_c64_%d" % N]) < 1e-6
        o = (N - 1) // 2
        assert not Eo[:, :o].any() and not Eo[:, o + ph.shape[1]:].any()
    Eo, ph = co.viterbiviterbi(g["in_c128"], 7, 8)
    assert np.max(np.abs(ph - g["ph_c128"])) < 1e-13 and rms(Eo - g["out_c128"]) < 1e-13
