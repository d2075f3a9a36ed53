"""N1 of the drop-in boundary: the reference's own ``Scripts/*_equalisation.py`` run UNCHANGED with the CUDA path's
L1 wrappers installed underneath QAMpy (``qampy_b200.patch("l1")``), and compute what the unpatched reference computes.

As in ``test_dropin_reference.py`` this container has the reference but no GPU, so the shared library -- and only the
shared library -- is replaced by the oracle-backed stand-in with the C ABI's ``*_host`` entry points; patching, the L1
wrappers, their pointer marshalling, dtype handling (the scripts run in complex128 with ``adaptive_stepsize=(True,
True)`` and ``mddma`` / ``sbd``) and everything QAMpy does around the kernels is the shipped / the reference's code.
Expected values: digests of the UNPATCHED runs (``tests/golden/make_golden_scripts.py``; the interpreted reference
needs minutes per script); ``QB_SCRIPTS_LIVE=1`` re-runs the unpatched reference in the test as well.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import scripts_runner as sr  # noqa: E402
from test_dropin_reference import OracleLib  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(sr.REF, "qampy")), reason="reference checkout not present")


def _close(got, want, rel):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    scale = max(float(np.max(np.abs(want))), 1e-30)
    assert float(np.max(np.abs(got - want))) <= rel * scale, (float(np.max(np.abs(got - want))), scale)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name", ["mrde_equaliser.py", "64_qam_equalisation.py", "32_qam_equalisation.py"])
def test_reference_script_runs_unchanged_through_the_patch(name, monkeypatch):
    root = os.path.dirname(HERE)
    for pth in (sr.REF, os.path.join(root, "oracle")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import cpu_oracle as co
    from qampy_b200 import _lib, patch
    lib = OracleLib(co, _lib.METHODS)
    monkeypatch.setattr(_lib, "load", lambda: lib)
    gold = np.load(os.path.join(HERE, "golden", "g12_scripts.npz"))
    want = {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(name + "/")}
    if os.environ.get("QB_SCRIPTS_LIVE") == "1":
        ns_ref, _ = sr.run_script(name)
        for k, v in sr.summary(name, ns_ref).items():
            assert np.array_equal(v, want[k]), "golden digest is stale: " + k
    with patch.patched("l1"):
        ns, text = sr.run_script(name)
    assert lib.calls.count("train") >= 2 and "apply" in lib.calls        # the script's calls went through the C ABI
    got = sr.summary(name, ns)
    assert set(got) == set(want)
    # the scripts run in complex128: the CUDA path's arithmetic contract (any association of the same real operations)
    # leaves 1e-12-level differences that the adaptive step size and the decision-directed stage do not amplify
    _close(got["taps"], want["taps"], 1e-8)
    _close(got["sig_head"], want["sig_head"], 1e-8)
    _close(got["sig_rms"], want["sig_rms"], 1e-9)
    _close(got["sig_tail"], want["sig_tail"], 1e-8)
    for k in ("err1_rms", "err2_rms"):
        _close(got[k], want[k], 1e-8)
    for k in ("err1_tail", "err2_tail"):
        _close(got[k], want[k], 1e-7)
    # what the scripts print / plot: EVM and GMI
    _close(got["evm"], want["evm"], 1e-8)
    if "gmi" in want:
        _close(got["gmi"], want["gmi"], 1e-8)
        _close(got["gmi_per_bit"], want["gmi_per_bit"], 1e-8)
        assert "array" in text                                           # the GMI line was printed
