"""Drop-in for the reference's L1 equaliser kernels (NumPy arrays in, NumPy arrays out).

Mirrors the call signatures of ``qampy/core/equalisation/pythran_equalisation.py``:

* ``train_equaliser(E, TrSyms, Niter, os, mu, wx, modes, adaptive, symbols, method)``  (:128-173)
* ``apply_filter_to_signal(E, os, wx, modes=None)``                                    (:33-76)

Both go through the C ABI's ``*_host`` entry points, i.e. hand-written CUDA on the current device;
there is no CPU path.  Where compiled Pythran would raise ``TypeError`` for a dtype combination
outside its export list (E / wx / symbols of different width, Python-int modes) this shim casts.
"""
import ctypes

import numpy as np

from . import _lib
from .theory import unique_alphabet


def _ctype(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.complex64:
        return _lib.QB_C64, np.float32, np.complex64
    if dtype == np.complex128:
        return _lib.QB_C128, np.float64, np.complex128
    raise TypeError("E must be complex64 or complex128, got %s" % dtype)


_REAL_TO_CPLX = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}
# error functions of train_equaliser_realvalued (pythran_equalisation.py:82-91) -> C ABI method ids
_REAL_METHODS = {"cma": 10, "sgncma": 11, "dd": 12, "dd_data": 13}


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def train_equaliser(E, TrSyms, Niter, os, mu, wx, modes, adaptive, symbols, method, mu_shared=True):
    """Train the equaliser taps.  ``wx`` is updated in place and returned, like the reference
    (:170, :173).  Returns ``(err, wx, mu)``.

    ``mu_shared`` only matters for ``adaptive=True`` with more than one mode: True carries one step
    size through the modes in list order (what the reference does when interpreted, :130/:162-172),
    False starts every mode from ``mu``.
    """
    if method not in _lib.METHODS:
        raise ValueError("Unknown method %s" % method)
    code, rt, ct = _ctype(np.asarray(E).dtype)
    E = np.ascontiguousarray(E, dtype=ct)
    if E.ndim != 2:
        raise ValueError("E must be 2-dimensional (modes, samples)")
    wx_in = wx
    wx_c = np.ascontiguousarray(wx, dtype=ct)
    if wx_c.ndim != 3:
        raise ValueError("wx needs to be three dimensional")
    nmodes, L = E.shape
    ntaps = wx_c.shape[-1]
    assert wx_c.shape[0] == nmodes and wx_c.shape[1] == nmodes, \
        "wx needs to have at least as many dimensions as the maximum mode"
    symbols = np.ascontiguousarray(symbols, dtype=ct)
    assert symbols.ndim == 2 and symbols.shape[0] == nmodes, "symbols must be at least size of modes"
    # searched alphabets without their repeats: same decisions, shorter search (theory.unique_alphabet)
    symbols = np.ascontiguousarray(unique_alphabet(symbols, method))
    modes = np.ascontiguousarray(np.atleast_1d(modes), dtype=np.int64)
    TrSyms, Niter, os = int(TrSyms), int(Niter), int(os)
    err = np.zeros((nmodes, TrSyms * Niter), dtype=ct)
    mu_io = np.array([mu], dtype=rt)
    _lib.check(_lib.load().qb_train_equaliser_host(
        code, _p(E), nmodes, L, TrSyms, Niter, os, _p(mu_io), _p(wx_c), ntaps, _p(modes), modes.size,
        int(bool(adaptive)), _p(symbols), symbols.shape[1], _lib.METHODS[method], int(bool(mu_shared)), _p(err)))
    if wx_c is not wx_in:
        if isinstance(wx_in, np.ndarray) and wx_in.dtype == ct:
            np.copyto(wx_in, wx_c)     # keep the in-place contract for non-contiguous views
            wx_c = wx_in
    return err, wx_c, mu_io[0]


def apply_filter_to_signal(E, os, wx, modes=None):
    """Static MIMO FIR + decimation: returns (len(modes), (L - ntaps + 1)//os).  Like the reference's
    export list (:33-36), real E with real wx is accepted: it runs on the complex kernel with the imaginary
    parts at exactly zero (same products, same order) and the real part is returned."""
    assert os > 0, "oversampling factor must be larger than 0"
    Ea = np.asarray(E)
    if Ea.dtype in _REAL_TO_CPLX:
        if np.iscomplexobj(wx):
            raise TypeError("real-valued E needs real-valued wx (pythran_equalisation.py:33-36)")
        ct = _REAL_TO_CPLX[Ea.dtype]
        out = apply_filter_to_signal(Ea.astype(ct), os, np.asarray(wx).astype(ct), modes)
        return np.ascontiguousarray(out.real)
    code, rt, ct = _ctype(Ea.dtype)
    if not np.iscomplexobj(wx):
        raise TypeError("complex-valued E needs complex-valued wx (pythran_equalisation.py:33-36)")
    E = np.ascontiguousarray(E, dtype=ct)
    wx = np.ascontiguousarray(wx, dtype=ct)
    nmodes_max = wx.shape[0]
    ntaps = wx.shape[-1]
    if modes is None:
        modes = np.arange(nmodes_max)
    else:
        modes = np.atleast_1d(modes)
        assert np.max(modes) < nmodes_max, "largest mode number is larger than shape of signal"
    modes = np.ascontiguousarray(modes, dtype=np.int64)
    nmodes, L = E.shape
    N = max((L - ntaps + 1) // int(os), 0)
    out = np.zeros((modes.size, N), dtype=ct)
    _lib.check(_lib.load().qb_apply_filter_to_signal_host(code, _p(E), nmodes, L, int(os), _p(wx), ntaps,
                                                         _p(modes), modes.size, _p(out)))
    return out


def train_equaliser_realvalued(E, TrSyms, Niter, os, mu, wx, modes, adaptive, symbols, method, mu_shared=True):
    """Real-valued MIMO trainer (:80-111; the 4x4 real form of a dual-polarisation signal): real ``E``
    (nmodes, L), real ``wx`` (nmodes, nmodes, ntaps), real ``symbols``; ``method`` in cma / sgncma / dd /
    dd_data.  Runs on the complex kernels with every imaginary part held at exactly zero -- the real
    recurrence ``wx += mu * err * X`` is what the complex one does to real data -- with the real-valued
    error functions and step-size rule selected by their own method ids.  Returns ``(err, wx, mu)``;
    ``wx`` is updated in place like the reference."""
    if method not in _REAL_METHODS:
        raise ValueError("Unknown method %s" % method)
    Ea = np.asarray(E)
    if Ea.dtype not in _REAL_TO_CPLX:
        raise TypeError("E must be float32 or float64, got %s" % Ea.dtype)
    ct = _REAL_TO_CPLX[Ea.dtype]
    rt = Ea.dtype.type
    if Ea.ndim != 2:
        raise ValueError("E must be 2-dimensional (modes, samples)")
    wx_c = np.ascontiguousarray(np.asarray(wx), dtype=ct)
    if wx_c.ndim != 3:
        raise ValueError("wx needs to be three dimensional")
    nmodes, L = Ea.shape
    ntaps = wx_c.shape[-1]
    symbols = np.ascontiguousarray(np.atleast_2d(symbols), dtype=ct)
    assert symbols.shape[0] == nmodes, "symbols must be at least size of modes"
    assert wx_c.shape[0] == nmodes, "wx needs to have at least as many dimensions as the maximum mode"
    modes = np.ascontiguousarray(np.atleast_1d(modes), dtype=np.int64)
    assert modes.max() < nmodes, "Maximum mode number must not be higher than number of modes"
    TrSyms, Niter, os = int(TrSyms), int(Niter), int(os)
    assert L > TrSyms * os + ntaps, "Field must be longer than the number of training symbols"
    Ec = np.ascontiguousarray(Ea, dtype=ct)
    err = np.zeros((nmodes, TrSyms * Niter), dtype=ct)
    mu_io = np.array([mu], dtype=rt)
    _lib.check(_lib.load().qb_train_equaliser_host(
        _lib.QB_C64 if ct == np.complex64 else _lib.QB_C128, _p(Ec), nmodes, L, TrSyms, Niter, os, _p(mu_io),
        _p(wx_c), ntaps, _p(modes), modes.size, int(bool(adaptive)), _p(symbols), symbols.shape[1],
        _REAL_METHODS[method], int(bool(mu_shared)), _p(err)))
    wx_r = np.ascontiguousarray(wx_c.real)
    if isinstance(wx, np.ndarray) and wx.dtype == Ea.dtype:
        np.copyto(wx, wx_r)            # in-place contract (:107)
        wx_r = wx
    return np.ascontiguousarray(err.real), wx_r, mu_io[0]


def make_decision(E, symbols):
    """Decision operator (:306-334): for every sample the nearest alphabet point.  Returns
    ``(det_symbs, dist, idx)`` -- decided symbols, their distance ``np.abs(E - s)`` and the int32 index."""
    code, rt, ct = _ctype(np.asarray(E).dtype)
    E = np.ascontiguousarray(E, dtype=ct)
    if E.ndim != 1:
        raise ValueError("E must be 1-dimensional")
    symbols = np.ascontiguousarray(symbols, dtype=ct).reshape(-1)
    det = np.zeros_like(E)
    dist = np.zeros(E.shape, dtype=rt)
    idx = np.zeros(E.shape, dtype=np.int32)
    _lib.check(_lib.load().qb_make_decision_host(code, _p(E), E.shape[0], _p(symbols), symbols.size, _p(det), _p(dist),
                                                 _p(idx)))
    return det, dist, idx
