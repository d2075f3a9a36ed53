"""The reference's OWN test files, unmodified, with ``qampy_b200.patch("l1")`` installed underneath QAMpy
("re-run the reference's statistical tests through the patched API"; VERDICT r01 missing item 2):

* ``test/test_phaserec.py``  -- all 65 tests: return objects, 1-D / 2-D input, dtypes, and ``TestCorrect`` (:124-145): a
  known rotation is found by ``bps`` and ``bps_twostage`` to within one test angle, zero symbol errors;
* ``test/test_equalisation.py`` -- ``TestReturnObject`` / ``TestEqualisation`` (:10-46);
* ``test/test_signal_recover_functional.py`` -- ``TestLMS`` (:162-185): sbd, mddma, dd, sbd_data, rde, mrde and the
  real-valued dd_real / dd_data_real, complex64 and complex128, ``adaptive_stepsize=True``, ``Niter=3``: at most three
  wrong symbols.  ``QB_REF_SUITE_FULL=1`` adds ``TestDualMode`` (without ``test_pmd_phase``, which fails on the unpatched reference
  itself) and ``TestCMA::test_pol_rot`` (2 minutes).

Each selection runs in a subprocess with ``tests/ref_patch_plugin.py`` (patch + repeatable randomness; in this
container the shared library is the oracle-backed stand-in, as in ``test_dropin_reference.py``) and must pass; the
plugin's report of C-ABI calls proves the reference's tests went through the drop-in entry points.
"""
import os
import re
import subprocess
import sys

import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "test")), reason="reference checkout not present")

SELECTIONS = [
    ("test_phaserec.py", None, 65, ("bps", "bps_rows", "select")),
    ("test_equalisation.py", "TestReturnObject or TestEqualisation", 6, ("train", "apply")),
    ("test_signal_recover_functional.py", "TestLMS", 16, ("train", "apply", "decide")),
]
if os.environ.get("QB_REF_SUITE_FULL") == "1":
    # (TestDualMode::test_pmd_phase -- laser phase noise in front of a phase-locking first stage -- fails on the
    # UNPATCHED reference for all eight of its parameter sets with the plugin's seed, and for two of them patched)
    SELECTIONS.append(("test_signal_recover_functional.py", "(TestDualMode and not test_pmd_phase) or test_pol_rot", 30,
                       ("train", "apply")))


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("fname,kexpr,min_passed,must_call", SELECTIONS)
def test_reference_tests_pass_through_the_patch(fname, kexpr, min_passed, must_call, tmp_path):
    cmd = [sys.executable, "-m", "pytest", os.path.join(REF, "test", fname), "-p", "ref_patch_plugin", "-q",
           "-p", "no:cacheprovider"]
    if kexpr:
        cmd += ["-k", kexpr]
    env = dict(os.environ)
    env["PYTHONPATH"] = HERE + os.pathsep + env.get("PYTHONPATH", "")
    res = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True)
    tail = (res.stdout + res.stderr)[-3000:]
    m = re.search(r"(\d+) passed", res.stdout)
    assert res.returncode == 0 and m and " failed" not in res.stdout.splitlines()[-1], tail
    assert int(m.group(1)) >= min_passed, tail
    calls = re.search(r"QB_CABI_CALLS (.*)", res.stderr)
    assert calls, tail
    seen = dict(kv.split("=") for kv in calls.group(1).split())
    for name in must_call:
        assert int(seen.get(name, 0)) > 0, (name, seen)
