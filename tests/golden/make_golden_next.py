"""Golden fixtures for the rows SURVEY.md section 8f marks "next", generated like make_golden.py by
EXECUTING THE REFERENCE (interpreted Pythran sources, asserts stripped):

    python tests/golden/make_golden_next.py

G7  two-stage blind phase search (qampy/core/phaserecovery.py:222-288): the per-symbol test-angle table
    form of pythran_dsp.bps (p == L) and select_angles, c64 and c128.
G8  real-valued equaliser methods (equalisation.py:529-565 -> pythran_equalisation.py:80-128):
    cma_real, sgncma_real, dd_real, dd_data_real through equalise_signal(apply=True), fixed and adaptive
    step size, c64 and c128.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    load_reference()
    from qampy import signals, impairments
    from qampy.core import phaserecovery as cph
    from qampy.core import pythran_dsp as pd

    out = {}
    for tag, dt, M, nsym, A, N, B, seed in (("c64", np.complex64, 16, 1200, 16, 8, 4, 71),
                                            ("c128", np.complex128, 64, 900, 32, 10, 6, 72)):
        sg = signals.SignalQAMGrayCoded(M, nsym, nmodes=2, fb=40e9, dtype=dt, seed=[seed, seed + 1])
        np.random.seed(seed)
        s2 = impairments.apply_phase_noise(impairments.change_snr(sg, 22), 300e3)
        E = np.asarray(s2)
        coded = np.asarray(sg.coded_symbols)
        En, ph = cph.bps_twostage(E, A, coded, N, B=B)
        # the L1 calls of the second stage for mode 0, step by step (phaserecovery.py:271-281)
        rt = E.real.dtype
        angles = np.linspace(-np.pi / 4, np.pi / 4, A, endpoint=False, dtype=rt).reshape(1, -1)
        idx1 = pd.bps(np.copy(E[0]), angles, coded, N)
        ph1 = pd.select_angles(np.copy(angles), idx1)
        b = np.linspace(-B / 2, B / 2, B)
        phn = (ph1[:, np.newaxis] + b[np.newaxis, :] / (B * A) * np.pi / 2).astype(rt)
        idx2 = pd.bps(np.copy(E[0]), phn, coded, N)
        phf = pd.select_angles(np.copy(phn), idx2)
        out.update({"in_" + tag: E, "coded_" + tag: coded, "A_" + tag: A, "N_" + tag: N, "B_" + tag: B,
                    "out_" + tag: np.asarray(En), "ph_" + tag: ph, "idx1_" + tag: idx1, "phn_" + tag: phn,
                    "idx2_" + tag: idx2, "phf_" + tag: phf})
    np.savez_compressed(os.path.join(HERE, "g7_bps_twostage.npz"), **out)

    # ---- G8: real-valued equaliser methods ---------------------------------------------------------------
    from make_golden import make_signal
    from qampy.core.equalisation import equalisation as ceq
    out = {}
    for tag, dt, M, seed in (("c64", np.complex64, 4, 81), ("c128", np.complex128, 16, 82)):
        sig, s = make_signal(M, 700, 2, dt, seed, 22, theta=np.pi / 5.5, dgd=20e-12)
        E = np.asarray(s)
        out["E_" + tag] = E
        out["M_" + tag] = M
        out["tx_" + tag] = np.asarray(sig)
        for method, adaptive in (("cma_real", False), ("sgncma_real", False), ("dd_real", False),
                                 ("dd_data_real", False), ("cma_real", True), ("dd_real", True)):
            kw = {}
            if method == "dd_data_real":
                kw["symbols"] = np.asarray(sig)
            Eo, wxy, err = ceq.equalise_signal(E, 2, 2e-3, M, Ntaps=7, method=method, apply=True,
                                               adaptive_stepsize=adaptive, **kw)
            key = "%s_%s%s" % (tag, method, "_ad" if adaptive else "")
            out["out_" + key], out["wxy_" + key], out["err_" + key] = np.asarray(Eo), wxy, err
    np.savez_compressed(os.path.join(HERE, "g8_real_valued.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
