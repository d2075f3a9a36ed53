"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE.

Run in the development container only (``/root/reference`` is not present on the GPU box):

    python tests/golden/make_golden.py

The reference's hot-path kernels are Pythran sources, i.e. valid Python; they are executed
*interpreted* (NumPy semantics).  ``qampy/core/pythran_dsp.py:69`` asserts ``p == 0 or p == L``
which is false for the normal ``p == 1`` call and is compiled out by ``-DNDEBUG`` in the real
build (``setup.py:28``), so the two kernel modules are re-executed with ``optimize=1`` (asserts
stripped) before anything is called.  The reference has no golden vectors of its own for this
path (its tests are statistical), so these files ARE the parity pin.

Everything is seeded: ``np.random.seed`` for the impairments and ``seed=`` for the bit source.
"""
import importlib
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("QAMPY_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """Import the reference with the two Pythran-source modules stripped of asserts."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import qampy  # noqa: F401
    from qampy.core import pythran_dsp, phaserecovery
    from qampy.core.equalisation import pythran_equalisation
    for mod in (pythran_dsp, pythran_equalisation):
        with open(mod.__file__) as fh:
            src = fh.read()
        exec(compile(src, mod.__file__, "exec", optimize=1), mod.__dict__)
    phaserecovery._bps_idx_pyt = pythran_dsp.bps        # from-imported at import time (:28-29)
    phaserecovery.select_angles = pythran_dsp.select_angles
    return importlib.import_module("qampy")


def make_signal(M, nsym, nmodes, dtype, seed, snr, theta=None, dgd=None, beta=0.1, lw=None):
    from qampy import signals, impairments
    np.random.seed(seed)
    sig = signals.SignalQAMGrayCoded(M, nsym, nmodes=nmodes, fb=40e9, dtype=dtype,
                                     seed=[seed + 10 + i for i in range(nmodes)])
    s = sig.resample(2 * sig.fb, beta=beta, renormalise=True)
    s = impairments.change_snr(s, snr)
    if lw:
        s = impairments.apply_phase_noise(s, lw)
    if theta is not None and nmodes == 2:
        s = impairments.apply_PMD(s, theta, dgd)
    return sig, s


def main():
    load_reference()
    from qampy import equalisation, phaserec, impairments
    from qampy.core.equalisation import equalisation as ceq
    from qampy.core import phaserecovery as cph
    from qampy.core.equalisation import pythran_equalisation as pe
    from qampy.core import pythran_dsp as pd
    from qampy import theory

    # ---- G0: constants --------------------------------------------------------------------
    out = {}
    for M in (4, 16, 32, 64, 128, 256):
        out["syms_%d" % M] = theory.cal_symbols_qam(M)
        out["scale_%d" % M] = np.float64(theory.cal_scaling_factor_qam(M))
        for m in ("cma", "cma2", "sgncma", "mcma", "rde", "mrde", "sbd", "mddma", "dd"):
            out["eqsyms_%s_%d" % (m, M)] = ceq.generate_symbols_for_eq(m, M, np.complex128)
    for L, nt, os_ in ((20000, 11, 2), (2000000, 21, 2), (20000000, 45, 2), (6000, 11, 2), (5000, 7, 1)):
        out["trsyms_%d_%d_%d" % (L, nt, os_)] = np.int64(ceq._cal_training_symbol_len(os_, nt, L))
    np.savez_compressed(os.path.join(HERE, "g0_constants.npz"), **out)

    # ---- G1: BASELINE config 1 -- single-pol QPSK CMA ntaps=11, 1e4 symbols, c64 ------------
    sig, s = make_signal(4, 10 ** 4, 1, np.complex64, 1, 20)
    E, wxy, err = equalisation.equalise_signal(s, 1e-3, Ntaps=11, method="cma", apply=True)
    np.savez_compressed(os.path.join(HERE, "g1_c1_cma.npz"), E_in=np.asarray(s), os=2, M=4, mu=1e-3,
                        ntaps=11, E_out=np.asarray(E), wxy=wxy, err=err,
                        symbols_tx=np.asarray(sig.symbols), coded=np.asarray(sig.coded_symbols))

    # ---- G2: dual-pol 16-QAM MCMA->MRDE ntaps=11 + BPS(32, 10), both widths -----------------
    for dt, tag in ((np.complex64, "c64"), (np.complex128, "c128")):
        sig, s = make_signal(16, 3000, 2, dt, 2, 25, np.pi / 5, 30e-12)
        E, wxy, (e1, e2) = equalisation.dual_mode_equalisation(s, (1e-3, 1e-3), 11,
                                                               methods=("mcma", "mrde"))
        np.random.seed(22)
        Epn = impairments.apply_phase_noise(E, 100e3)
        Eb, ph = phaserec.bps(Epn, 32, 10)
        idx = np.array([pd.bps(np.asarray(Epn)[i].copy(),
                               np.linspace(-np.pi / 4, np.pi / 4, 32, endpoint=False,
                                           dtype=ph.dtype).reshape(1, -1),
                               np.asarray(sig.coded_symbols), 10) for i in range(2)])
        np.savez_compressed(os.path.join(HERE, "g2_dual16_%s.npz" % tag), E_in=np.asarray(s), os=2, M=16,
                            mu=(1e-3, 1e-3), ntaps=11, E_out=np.asarray(E), wxy=wxy, err1=e1, err2=e2,
                            coded=np.asarray(sig.coded_symbols), bps_in=np.asarray(Epn),
                            bps_out=np.asarray(Eb), bps_ph=ph, bps_idx=idx, bps_A=32, bps_N=10)

    # ---- G3: dual-pol 64-QAM MCMA->MRDE ntaps=15 + BPS(64, 20) (C3 recipe, small) -----------
    sig, s = make_signal(64, 5000, 2, np.complex64, 3, 28, np.pi / 5.6, 40e-12)
    E, wxy, (e1, e2) = equalisation.dual_mode_equalisation(s, (1e-3, 1e-3), 15, methods=("mcma", "mrde"))
    np.random.seed(33)
    Epn = impairments.apply_phase_noise(E, 100e3)
    Eb, ph = phaserec.bps(Epn, 64, 20)
    idx = np.array([pd.bps(np.asarray(Epn)[i].copy(),
                           np.linspace(-np.pi / 4, np.pi / 4, 64, endpoint=False,
                                       dtype=np.float32).reshape(1, -1),
                           np.asarray(sig.coded_symbols), 20) for i in range(2)])
    np.savez_compressed(os.path.join(HERE, "g3_dual64_c64.npz"), E_in=np.asarray(s), os=2, M=64,
                        mu=(1e-3, 1e-3), ntaps=15, E_out=np.asarray(E), wxy=wxy, err1=e1, err2=e2,
                        coded=np.asarray(sig.coded_symbols), bps_in=np.asarray(Epn), bps_out=np.asarray(Eb),
                        bps_ph=ph, bps_idx=idx, bps_A=64, bps_N=20)

    # ---- G4: every error function, core-level call, Niter=2, dual-pol 16-QAM ----------------
    sig, s = make_signal(16, 1500, 2, np.complex64, 4, 22, np.pi / 7, 20e-12)
    E_in = np.asarray(s)
    out = dict(E_in=E_in, os=2, M=16, ntaps=7, mu=2e-3, Niter=2, coded=np.asarray(sig.coded_symbols),
               symbols_tx=np.asarray(sig.symbols))
    for m in ("cma", "cma2", "sgncma", "mcma", "rde", "mrde", "sbd", "mddma", "dd", "sbd_data"):
        for dt, tag in ((np.complex64, "c64"), (np.complex128, "c128")):
            if tag == "c128" and m not in ("mcma", "mrde", "sbd", "rde"):
                continue
            sy = np.asarray(sig.symbols) if m == "sbd_data" else np.asarray(sig.coded_symbols)
            wxy, err = ceq.equalise_signal(E_in.astype(dt), 2, 2e-3, 16, Ntaps=7, Niter=2, method=m,
                                           symbols=sy.astype(dt))
            out["wxy_%s_%s" % (m, tag)] = wxy
            out["err_%s_%s" % (m, tag)] = err
    # adaptive step size: single selected mode (deterministic in the reference) and both modes
    # (mu carried from mode 0 into mode 1 by the interpreted reference)
    for modes, tag in (([1], "m1"), ([0, 1], "m01")):
        E_c = E_in.copy()
        w0 = ceq._init_taps(7, 2, 2, np.complex64)
        sy = ceq._reshape_symbols(None, "mcma", 16, np.complex64, 2)
        err, w, mu = pe.train_equaliser(E_c, 700, 2, 2, np.float32(1e-2), w0, np.array(modes), True,
                                        sy.copy(), "mcma")
        out["ad_wxy_" + tag] = w
        out["ad_err_" + tag] = err
        out["ad_mu_" + tag] = np.float32(mu)
    np.savez_compressed(os.path.join(HERE, "g4_methods.npz"), **out)

    # ---- G5: apply_filter_to_signal shapes: mode subset, os=1, nmodes=1/3 --------------------
    rng = np.random.default_rng(5)
    out = {}
    for tag, nm, L, nt, os_, modes in (("a", 2, 1001, 9, 2, None), ("b", 2, 1000, 8, 2, [1]),
                                       ("c", 3, 500, 5, 1, [2, 0]), ("d", 1, 64, 64, 2, None),
                                       ("e", 2, 300, 45, 3, None)):
        E = (rng.standard_normal((nm, L)) + 1j * rng.standard_normal((nm, L))).astype(np.complex64)
        w = (rng.standard_normal((nm, nm, nt)) + 1j * rng.standard_normal((nm, nm, nt))).astype(np.complex64) / nt
        o = pe.apply_filter_to_signal(E, os_, w, None if modes is None else np.array(modes))
        out.update({"E_" + tag: E, "w_" + tag: w, "os_" + tag: os_, "out_" + tag: o,
                    "modes_" + tag: np.array([-1] if modes is None else modes)})
    np.savez_compressed(os.path.join(HERE, "g5_apply.npz"), **out)

    # ---- G6: BPS known answer (test/test_phaserec.py:124-145 pattern) + 1-D + cross-QAM ------
    from qampy import signals
    out = {}
    for k, ang in enumerate(np.linspace(0.1, np.pi / 4.1, 3)):
        sg = signals.SignalQAMGrayCoded(32, 2 ** 11, nmodes=1, fb=40e9, dtype=np.complex64, seed=[60 + k])
        s2 = sg * np.exp(1j * ang)
        Eb, ph = phaserec.bps(s2, 32, 11)
        out.update({"in_%d" % k: np.asarray(s2), "out_%d" % k: np.asarray(Eb), "ph_%d" % k: ph,
                    "angle_%d" % k: ang, "coded": np.asarray(sg.coded_symbols)})
    sg = signals.SignalQAMGrayCoded(16, 1500, nmodes=1, fb=40e9, dtype=np.complex128, seed=[66])
    np.random.seed(66)
    s2 = impairments.apply_phase_noise(impairments.change_snr(sg, 18), 400e3)
    Eb, ph = cph.bps(np.asarray(s2)[0], 16, np.asarray(sg.coded_symbols), 8)
    out.update(in_1d=np.asarray(s2)[0], out_1d=Eb, ph_1d=ph, coded_1d=np.asarray(sg.coded_symbols))
    np.savez_compressed(os.path.join(HERE, "g6_bps_kat.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
