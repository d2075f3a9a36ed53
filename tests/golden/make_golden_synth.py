"""Golden vectors for the on-device generators, produced by the REFERENCE's own functions on fixed
symbols:  qampy.core.resample.rrcos_resample (core/resample.py:73-126) per mode, then
qampy.core.impairments.apply_PMD_to_field (core/impairments.py:94-131).  Deterministic stages only -- the reference's
noise generators draw from np.random and cannot be reproduced sample by sample.

    python tests/golden/make_golden_synth.py        ->  tests/golden/g13_synth.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from qampy.core import impairments, resample  # noqa: E402
from qampy import theory  # noqa: E402

if __name__ == "__main__":
    rng = np.random.RandomState(20261017)
    out = {}
    for tag, M, n, beta, taps, theta, dgd, fb in (("a", 16, 4096, 0.1, 4001, np.pi / 5.6, 40e-12, 40e9),
                                                   ("b", 64, 3001, 0.35, 1001, np.pi / 3.1, 75e-12, 28e9)):
        al = theory.cal_symbols_qam(M) / np.sqrt(theory.cal_scaling_factor_qam(M))
        syms = al[rng.randint(0, M, (2, n))].astype(np.complex128)
        fs = 2 * fb
        shaped = np.array([resample.rrcos_resample(s, fb, fs, beta=beta, taps=taps, renormalise=True) for s in syms])
        plain = np.array([resample.rrcos_resample(s, fb, fs, beta=beta, taps=taps, renormalise=False) for s in syms])
        pmd = impairments.apply_PMD_to_field(shaped, theta, dgd, fs)
        out.update({tag + "_symbols": syms, tag + "_shaped": shaped, tag + "_plain": plain, tag + "_pmd": pmd,
                    tag + "_par": np.array([M, n, beta, taps, theta, dgd, fb])})
        print(tag, shaped.shape, pmd.shape, float(np.mean(np.abs(shaped) ** 2)), float(np.mean(np.abs(pmd) ** 2)))
    np.savez_compressed(os.path.join(HERE, "g13_synth.npz"), **out)
