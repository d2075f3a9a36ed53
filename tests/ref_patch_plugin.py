"""pytest plugin (test infrastructure): run the REFERENCE's own test files with qampy_b200.patch("l1") installed.

    python -m pytest /root/reference/test/test_phaserec.py -p ref_patch_plugin ...

Loaded by tests/test_reference_suite_under_patch.py in a subprocess.  In this container (reference checkout, no GPU) the
shared library is replaced by the oracle-backed stand-in of test_dropin_reference.py; with a GPU and QB_REF_SUITE_CUDA=1
the real library is used.  The interpreted reference kernel's ``assert p == 0 or p == L`` (pythran_dsp.py:69, compiled
away by Pythran) never runs: the patched entry points replace it."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for pth in (HERE, ROOT, os.path.join(ROOT, "oracle"), "/root/reference"):
    if pth not in sys.path:
        sys.path.insert(0, pth)

CALLS = []


def pytest_configure(config):
    import warnings
    warnings.filterwarnings("ignore")
    try:
        import matplotlib  # noqa: F401
    except ImportError:      # some reference test files import matplotlib at module level and never draw
        from scripts_runner import stub_module
        for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
            sys.modules[name] = stub_module(name)
    # the reference's tests draw their signals from unseeded generators: make the session repeatable
    from scripts_runner import _repeatable_randomness
    config._qb_rng = _repeatable_randomness(int(os.environ.get("QB_REF_SUITE_SEED", "4321")))
    config._qb_rng.__enter__()
    if os.environ.get("QB_REF_SUITE_UNPATCHED") == "1":      # the same session against the reference itself
        config._qb_lib = None
        return
    from qampy_b200 import _lib, patch
    if os.environ.get("QB_REF_SUITE_CUDA") != "1":
        import cpu_oracle as co
        from test_dropin_reference import OracleLib
        lib = OracleLib(co, _lib.METHODS)
        _lib.load = lambda: lib
        config._qb_lib = lib
    patch.patch("l1")


def pytest_unconfigure(config):
    rng = getattr(config, "_qb_rng", None)
    if rng is not None:
        rng.__exit__(None, None, None)
    if os.environ.get("QB_REF_SUITE_UNPATCHED") == "1":
        return
    from qampy_b200 import patch
    lib = getattr(config, "_qb_lib", None)
    if lib is not None:
        # proof for the caller that the reference's tests really went through the C ABI entry points
        sys.stderr.write("QB_CABI_CALLS %s\n" % " ".join("%s=%d" % (k, lib.calls.count(k)) for k in sorted(set(lib.calls))))
    patch.unpatch()
