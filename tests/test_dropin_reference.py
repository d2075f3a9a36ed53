"""The drop-in boundary with the REAL QAMpy on top: ``qampy.equalisation.equalise_signal /
dual_mode_equalisation / apply_filter`` and ``qampy.phaserec.bps`` called on QAMpy signal objects, with
``qampy_b200.patch("l1")`` installed underneath, must keep their signatures, return types (signal-object subclass,
dtype, ``recreate_from_np_array`` attributes) and give the reference's own numbers.

This container has the reference checkout but no GPU, and the GPU box has no reference checkout, so the ONLY piece
replaced here is the shared library: a stand-in object with the C ABI's ``*_host`` entry points that rebuilds the
NumPy arrays from the raw pointers and calls the CPU oracle (test infrastructure).  Everything a QAMpy user touches is
the shipped code: ``qampy_b200.patch``, the L1 wrappers ``qampy_b200.pythran_equalisation`` / ``pythran_dsp`` and their
pointer marshalling.  ``tests/test_gpu_golden.py`` checks the same wrappers against the reference's vectors with the
real library on the GPU."""
import ctypes
import os
import sys
import warnings

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "qampy")), reason="reference checkout not present")

_CT = {0: (np.complex64, np.float32), 1: (np.complex128, np.float64)}


def _addr(ptr):
    return ptr.value if isinstance(ptr, ctypes.c_void_p) else ptr


def _arr(ptr, shape, dtype):
    ptr = _addr(ptr)
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class OracleLib:
    """``libqampy_b200.so`` stand-in: same entry points, same pointer conventions, computed by the CPU oracle."""

    def __init__(self, co, methods):
        self.co = co
        self.names = {v: k for k, v in methods.items()}
        self.calls = []

    def qb_last_error(self):
        return b"oracle stand-in"

    def qb_train_equaliser_host(self, code, E, nmodes, L, TrSyms, Niter, os_, mu, wx, ntaps, modes, nsel, adaptive,
                                symbols, K, method, mu_shared, err):
        ct, rt = _CT[code]
        self.calls.append("train")
        if method not in self.names:
            return self._train_real(ct, rt, E, nmodes, L, TrSyms, Niter, os_, mu, wx, ntaps, modes, nsel, adaptive,
                                    symbols, K, method, err)
        Ea = _arr(E, (nmodes, L), ct)
        wa = _arr(wx, (nmodes, nmodes, ntaps), ct)
        ma = _arr(modes, (nsel,), np.int64)
        sa = _arr(symbols, (nmodes, K), ct)
        mua = _arr(mu, (1,), rt)
        e, w, m = self.co.train_equaliser(Ea, TrSyms, Niter, os_, mua[0], wa.copy(), ma.copy(), bool(adaptive), sa,
                                          self.names[method], mu_shared=bool(mu_shared))
        wa[...] = w
        mua[0] = m
        if _addr(err):
            _arr(err, (nmodes, TrSyms * Niter), ct)[...] = e
        return 0

    def _train_real(self, ct, rt, E, nmodes, L, TrSyms, Niter, os_, mu, wx, ntaps, modes, nsel, adaptive, symbols, K,
                    method, err):
        """real-valued methods (QB_*_REAL ids): the caller hands real data widened to complex with zero imaginary
        parts (qampy_b200.pythran_equalisation.train_equaliser_realvalued); the oracle's NumPy restatement of
        pythran_equalisation.py:80-128 works on the real parts"""
        from qampy_b200 import pythran_equalisation as q_pe
        name = {v: k for k, v in q_pe._REAL_METHODS.items()}[method]
        Ea = np.ascontiguousarray(_arr(E, (nmodes, L), ct).real)
        wa = _arr(wx, (nmodes, nmodes, ntaps), ct)
        wr = np.ascontiguousarray(wa.real)
        sa = np.ascontiguousarray(_arr(symbols, (nmodes, K), ct).real)
        mua = _arr(mu, (1,), rt)
        e, w, m = self.co.train_equaliser_realvalued(Ea, TrSyms, Niter, os_, mua[0], wr, _arr(modes, (nsel,), np.int64).copy(),
                                                     bool(adaptive), sa, name)
        wa[...] = w
        mua[0] = m
        if _addr(err):
            _arr(err, (nmodes, TrSyms * Niter), ct)[...] = e
        return 0

    def qb_apply_filter_to_signal_host(self, code, E, nmodes, L, os_, wx, ntaps, modes, nsel, out):
        ct, _ = _CT[code]
        self.calls.append("apply")
        ma = _arr(modes, (nsel,), np.int64)
        N = max((L - ntaps + 1) // os_, 0)
        wa = _arr(wx, (nmodes, nmodes, ntaps), ct)
        res = self.co.apply_filter_to_signal(_arr(E, (nmodes, L), ct), os_, wa, ma.copy())
        _arr(out, (nsel, N), ct)[...] = res
        return 0

    def qb_bps_host(self, code, E, nstream, L, comp, angles, A, symbols, M, N, idx, ph, Eout):
        ct, rt = _CT[code]
        self.calls.append("bps")
        assert nstream == 1
        # straight to the oracle's C entry point with the caller's rotation table (the table is the L1 wrapper's
        # np.exp(1j*testangles), pythran_dsp.py:72 -- no round trip through the angles)
        fn = getattr(self.co.lib("strict"), "qo_bps" + ("_f32" if code == 0 else "_f64"))
        assert not _addr(ph) and not _addr(Eout)
        rc = fn(_addr(E), 1, L, L, _addr(comp), 1, A, _addr(symbols), M, N, _addr(idx))
        assert rc == 0
        return 0

    def qb_make_decision_host(self, code, E, n, symbols, M, det, dist, idx):
        ct, rt = _CT[code]
        self.calls.append("decide")
        d, ds, ix = self.co.make_decision(_arr(E, (n,), ct), _arr(symbols, (M,), ct))
        _arr(det, (n,), ct)[...] = d
        _arr(dist, (n,), rt)[...] = ds
        _arr(idx, (n,), np.int32)[...] = ix
        return 0

    def qb_estimate_snr_host(self, code, rx, tx, n, gray, M, out):
        ct, _ = _CT[code]
        self.calls.append("snr")
        _arr(out, (3,), np.float64)[...] = self.co.estimate_snr(_arr(rx, (n,), ct), _arr(tx, (n,), ct),
                                                               _arr(gray, (M,), ct))
        return 0

    def qb_soft_l_value_demapper_host(self, code, rx, n, num_bits, snr, bits_map, nb, half, minmax, out):
        ct, _ = _CT[code]
        self.calls.append("demap")
        bm = _arr(bits_map, (nb, half, 2), ct)
        o = _arr(out, (n, num_bits), np.float64)
        for a in range(0, n, 32768):       # the oracle forms (chunk, M/2) distance tables
            o[a:a + 32768] = self.co.soft_l_value_demapper(_arr(rx, (n,), ct)[a:a + 32768], num_bits, snr, bm,
                                                            minmax=bool(minmax))
        return 0

    def qb_bps_rows_host(self, code, E, nstream, L, comp, angles, A, symbols, M, N, idx, ph, Eout):
        """per-symbol angle table comp[L][A] (pythran_dsp.py:74-77): the oracle's C entry point with L table rows"""
        self.calls.append("bps_rows")
        assert nstream == 1 and not _addr(ph) and not _addr(Eout)
        fn = getattr(self.co.lib("strict"), "qo_bps" + ("_f32" if code == 0 else "_f64"))
        rc = fn(_addr(E), 1, L, L, _addr(comp), L, A, _addr(symbols), M, N, _addr(idx))
        assert rc == 0
        return 0

    def qb_select_angles_host(self, code, angles, p, A, idx, L, out):
        rt = np.float32 if code == 0 else np.float64
        self.calls.append("select")
        res = self.co.select_angles(_arr(angles, (p, A), rt), _arr(idx, (L,), np.int64))
        _arr(out, (L,), rt)[...] = res
        return 0


@pytest.fixture()
def qampy_env(monkeypatch):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for pth in (REF, os.path.join(root, "oracle")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    warnings.filterwarnings("ignore")
    import cpu_oracle as co
    from qampy_b200 import _lib, patch
    from qampy import equalisation, impairments, phaserec, signals
    import qampy.core.pythran_dsp as ref_dsp
    import qampy.core.phaserecovery as cph
    lib = OracleLib(co, _lib.METHODS)
    monkeypatch.setattr(_lib, "load", lambda: lib)
    # the interpreted reference kernel asserts p == 0 or p == L (pythran_dsp.py:69), which is false for every normal
    # call and compiled away by Pythran (-DNDEBUG): give the UNPATCHED runs below the compiled behaviour
    src = open(ref_dsp.__file__).read().replace("assert p == 0 or p == L", "pass  #")
    ns = {}
    exec(compile(src, ref_dsp.__file__, "exec"), ns)
    monkeypatch.setattr(cph, "_bps_idx_pyt", ns["bps"])

    class NS:
        pass
    e = NS()
    e.lib, e.patch, e.eq, e.imp, e.ph, e.sig, e.co = lib, patch, equalisation, impairments, phaserec, signals, co
    yield e
    patch.unpatch()


def _signal(env, M, N, dtype, seed):
    np.random.seed(seed)
    s = env.sig.SignalQAMGrayCoded(M, N, nmodes=2, fb=40e9, dtype=dtype, seed=[seed, seed + 1])
    s = s.resample(2 * s.fb, beta=0.1, renormalise=True)
    s = env.imp.change_snr(s, 25)
    return env.imp.apply_PMD(s, np.pi / 5.6, 30e-12)


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 2e-5), (np.complex128, 1e-10)])
def test_equaliser_and_bps_on_signal_objects_through_the_patch(qampy_env, dtype, tol):
    env = qampy_env
    sig = _signal(env, 16, 3000, dtype, 3)
    # the reference itself (interpreted Pythran sources)
    E_ref, w_ref, err_ref = env.eq.dual_mode_equalisation(sig, (2e-3, 2e-3), 11, methods=("mcma", "mrde"))
    Es_ref, ws_ref, es_ref = env.eq.equalise_signal(sig, 2e-3, Ntaps=11, method="mcma", apply=True)
    Ea_ref = env.eq.apply_filter(sig, ws_ref)
    Eb_ref, ph_ref = env.ph.bps(E_ref, 32, 10)
    n0 = len(env.lib.calls)
    assert n0 == 0
    with env.patch.patched("l1"):
        E, w, err = env.eq.dual_mode_equalisation(sig, (2e-3, 2e-3), 11, methods=("mcma", "mrde"))
        Es, ws, es = env.eq.equalise_signal(sig, 2e-3, Ntaps=11, method="mcma", apply=True)
        Ea = env.eq.apply_filter(sig, ws)
        Eb, ph = env.ph.bps(E_ref, 32, 10)
    assert {"train", "apply", "bps", "select"} <= set(env.lib.calls)       # the calls really went through the C ABI
    rms = lambda a: float(np.sqrt(np.mean(np.abs(np.asarray(a)) ** 2)))
    # same types: signal-object subclass with its attributes, dtype preserved (test_equalisation.py:10-28,
    # test_phaserec.py:10-39, :106-121)
    for got, ref in ((E, E_ref), (Es, Es_ref), (Ea, Ea_ref), (Eb, Eb_ref)):
        assert type(got) is type(ref) and got.dtype == ref.dtype == dtype and got.shape == ref.shape
        assert got.fb == ref.fb and got.fs == ref.fs and got.M == ref.M
        assert np.array_equal(got.symbols, ref.symbols)
    assert w.dtype == w_ref.dtype and w.shape == w_ref.shape == (2, 2, 11)
    assert ph.dtype == ph_ref.dtype and ph.shape == ph_ref.shape
    # same numbers
    assert rms(E - E_ref) < tol and rms(Es - Es_ref) < tol and rms(Ea - Ea_ref) < tol
    assert np.max(np.abs(w - w_ref)) < tol and np.max(np.abs(ws - ws_ref)) < tol
    assert rms(err[0] - err_ref[0]) < 5 * tol and rms(err[1] - err_ref[1]) < 5 * tol and rms(es - es_ref) < 5 * tol
    assert np.array_equal(np.asarray(ph), np.asarray(ph_ref))                 # BPS: bit exact
    assert rms(Eb - Eb_ref) < 1e-6
    # and the chain still demodulates (what Scripts/*_equalisation.py print)
    assert float(np.max(Eb.cal_ser())) < 0.05


def test_l1_wrappers_keep_the_in_place_tap_contract(qampy_env):
    """``train_equaliser`` updates ``wx`` in place AND returns it (pythran_equalisation.py:170, :173); the L2 driver
    relies on the return value, user code may rely on either."""
    env = qampy_env
    import qampy_b200.pythran_equalisation as q_pe
    from qampy_b200 import theory
    sig = np.asarray(_signal(env, 4, 1500, np.complex64, 9))
    wx = theory.init_taps(7, 2, np.complex64)
    keep = wx
    sy = theory.reshape_symbols(None, "cma", 4, np.complex64, 2)
    err, wx2, mu = q_pe.train_equaliser(sig, 1400, 1, 2, np.float32(1e-3), wx, np.arange(2), False, sy, "cma")
    assert wx2 is keep and not np.array_equal(wx2, theory.init_taps(7, 2, np.complex64))
    assert err.shape == (2, 1400) and err.dtype == np.complex64 and np.float32(mu) == np.float32(1e-3)
    with pytest.raises(ValueError, match="Unknown method"):
        q_pe.train_equaliser(sig, 1400, 1, 2, 1e-3, wx, np.arange(2), False, sy, "nonsense")
