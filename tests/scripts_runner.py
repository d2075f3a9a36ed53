"""Test infrastructure: execute a reference ``Scripts/*.py`` file UNMODIFIED (SURVEY.md section 4 / BASELINE north_star:
"Scripts/*_equalisation.py run unchanged") and hand back what it computed.

What is arranged around the script, never inside it:
* ``matplotlib`` is replaced by a stub that accepts every call (the scripts end in ``plt.show()``);
* randomness is made repeatable: ``numpy.random.seed`` is set and ``numpy.random.RandomState(None)`` -- what
  ``qampy.signals`` uses for its bit streams when no seed is given (signals.py:76) -- draws its seed from a counter,
  so the patched and the unpatched run of a script see the same signal;
* stdout is captured.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"


class _Anything:
    """Accepts any attribute access, call, indexing or iteration: enough of matplotlib for the scripts."""

    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, k):
        return self

    def __iter__(self):
        return iter(())


def stub_module(name):
    """A module object whose every public attribute is an :class:`_Anything` (dunder look-ups fail as usual)."""
    stub = _Anything()
    m = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return stub
    m.__getattr__ = _getattr
    return m


@contextlib.contextmanager
def _stubbed_matplotlib():
    saved = {k: v for k, v in sys.modules.items() if k == "matplotlib" or k.startswith("matplotlib.")}
    for k in saved:
        del sys.modules[k]
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
        sys.modules[name] = stub_module(name)
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    try:
        yield
    finally:
        for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
            sys.modules.pop(name, None)
        sys.modules.update(saved)


@contextlib.contextmanager
def _repeatable_randomness(seed):
    real = np.random.RandomState
    counter = [int(seed)]

    class Seeded(real):
        def __init__(self, s=None):
            if s is None:
                counter[0] += 1
                s = counter[0]
            super().__init__(s)

    np.random.RandomState = Seeded
    state = np.random.get_state()
    np.random.seed(int(seed))
    try:
        yield
    finally:
        np.random.RandomState = real
        np.random.set_state(state)


def run_script(name, seed=1234):
    """exec ``/root/reference/Scripts/<name>`` as ``__main__`` would run it.  Returns (namespace, stdout text)."""
    path = os.path.join(REF, "Scripts", name)
    with open(path) as fh:
        src = fh.read()
    ns = {"__name__": "__main__", "__file__": path}
    out = io.StringIO()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with _stubbed_matplotlib(), _repeatable_randomness(seed), contextlib.redirect_stdout(out):
        exec(compile(src, path, "exec"), ns)
    return ns, out.getvalue()


# what each script leaves behind that is worth comparing: (taps, equalised signal, error arrays)
SCRIPTS = {
    "mrde_equaliser.py": dict(taps="wxy_s", sig="E_s", errs=("err_s", "err_rde_s"), evm="evmE_s", gmi="gmiE"),
    "64_qam_equalisation.py": dict(taps="wxy_s", sig="E_s", errs=("err_s", "err_rde_s"), evm="evmE_s", gmi="gmiE"),
    "32_qam_equalisation.py": dict(taps="wx", sig="E", errs=("err", "err_rde"), evm="evm", gmi=None),
    "cma_equaliser.py": None,       # filled in by summary() from whatever the script defines
}


def summary(name, ns):
    """Small, comparable digest of a script run (float64 / complex128 arrays)."""
    spec = SCRIPTS[name]
    d = {}
    if spec is None:
        return d
    d["taps"] = np.asarray(ns[spec["taps"]]).astype(np.complex128)
    E = np.asarray(ns[spec["sig"]])
    d["sig_head"] = E[:, :256].astype(np.complex128)
    d["sig_rms"] = np.sqrt(np.mean(np.abs(E) ** 2, axis=-1)).astype(np.float64)
    d["sig_tail"] = E[:, -256:].astype(np.complex128)
    for k, e in zip(("err1", "err2"), spec["errs"]):
        e = np.asarray(ns[e])
        d[k + "_rms"] = np.sqrt(np.mean(np.abs(e) ** 2, axis=-1)).astype(np.float64)
        d[k + "_tail"] = e[:, -64:].astype(np.complex128)
    d["evm"] = np.asarray(ns[spec["evm"]], dtype=np.float64)
    if spec["gmi"]:
        g = ns[spec["gmi"]]
        d["gmi"] = np.asarray(g[0], dtype=np.float64)
        d["gmi_per_bit"] = np.asarray(g[1], dtype=np.float64)
    return d
