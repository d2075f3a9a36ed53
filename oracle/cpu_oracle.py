"""ORACLE -- test infrastructure, NOT product code.

Python face of the CPU restatement (``qampy_oracle.c``) with the reference's own call
signatures, plus a NumPy restatement of the thin L2 drivers that sit between the public
API and the kernels.  Reference lines followed (paths relative to /root/reference/qampy):

* L1 kernels: ``core/equalisation/pythran_equalisation.py:37-76, 130-173``,
  ``core/pythran_dsp.py:47-85, 137-153`` (in C, see ``qo_kernels.inc``)
* L2 drivers: ``core/equalisation/equalisation.py:101-136, 138-188, 271-281, 311-373,
  400-466, 468-594`` and ``core/phaserecovery.py:141-159``
* constellations: ``theory.py:111-178``

Parity pin: ``tests/test_oracle_golden.py`` checks every function here against vectors
produced by executing the reference's Python sources on the same inputs
(``tests/golden/make_golden.py``).  The reference has no golden vectors of its own for
this path (its tests are statistical).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this module.
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as _build  # noqa: E402

METHODS = {"cma": 0, "cma2": 1, "sgncma": 2, "mcma": 3, "rde": 4, "mrde": 5, "sbd": 6,
           "sbd_data": 7, "mddma": 8, "dd": 9}
NONDECISION_BASED = ("cma", "cma2", "mcma", "rde", "mrde", "sgncma")
DATA_AIDED = ("sbd_data",)

_libs = {}
_c_long = ctypes.c_long
_c_vp = ctypes.c_void_p


def _declare(lib):
    for suf, real in (("_f32", ctypes.c_float), ("_f64", ctypes.c_double)):
        f = getattr(lib, "qo_train_equaliser" + suf)
        f.restype = ctypes.c_int
        f.argtypes = [_c_vp, _c_long, _c_long, _c_long, _c_long, _c_long, _c_long, _c_long, real,
                      _c_vp, _c_long, _c_vp, _c_long, ctypes.c_int, _c_vp, _c_long, ctypes.c_int,
                      ctypes.c_int, _c_vp, _c_vp]
        f = getattr(lib, "qo_apply_filter_to_signal" + suf)
        f.restype = ctypes.c_int
        f.argtypes = [_c_vp, _c_long, _c_long, _c_long, _c_long, _c_long, _c_long, _c_vp, _c_long,
                      _c_vp, _c_long, _c_vp]
        f = getattr(lib, "qo_bps" + suf)
        f.restype = ctypes.c_int
        f.argtypes = [_c_vp, _c_long, _c_long, _c_long, _c_vp, _c_long, _c_long, _c_vp, _c_long,
                      _c_long, _c_vp]
        f = getattr(lib, "qo_select_angles" + suf)
        f.restype = ctypes.c_int
        f.argtypes = [_c_vp, _c_long, _c_long, _c_vp, _c_long, _c_vp]
    lib.qo_max_threads.restype = ctypes.c_int
    lib.qo_set_threads.argtypes = [ctypes.c_int]
    return lib


def lib(kind="strict"):
    """kind: 'strict' (parity oracle), 'fast' (portable timed build), 'fast_native'."""
    if kind not in _libs:
        if kind == "strict":
            path = _build.build_strict()
        else:
            path = _build.build_fast(native=(kind == "fast_native"))
        _libs[kind] = _declare(ctypes.CDLL(path))
    return _libs[kind]


def _suf(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.complex64 or dtype == np.float32:
        return "_f32", np.float32, np.complex64
    if dtype == np.complex128 or dtype == np.float64:
        return "_f64", np.float64, np.complex128
    raise TypeError("oracle supports complex64/complex128 only, got %s" % dtype)


def _p(a):
    return a.ctypes.data_as(_c_vp)


# --------------------------------------------------------------------------------------
# L1: same signatures as the reference's Pythran exports
# --------------------------------------------------------------------------------------
def train_segments(E, TrSyms, Niter, os_, mu, wx, modes, adaptive, symbols, method,
                   mu_shared=True, kind="strict"):
    """Batched trainer: E (nseg, nmodes, L), wx (nseg, nmodes, nmodes, ntaps) updated in place.
    Returns err (nseg, nmodes, TrSyms*Niter), wx, mu (nseg,)."""
    if method not in METHODS:
        raise ValueError("Unknown method %s" % method)
    suf, rt, ct = _suf(E.dtype)
    E = np.ascontiguousarray(E, dtype=ct)
    assert E.ndim == 3 and wx.ndim == 4 and wx.dtype == ct and wx.flags.c_contiguous
    nseg, nmodes, L = E.shape
    ntaps = wx.shape[-1]
    symbols = np.ascontiguousarray(symbols, dtype=ct)
    assert symbols.ndim == 2 and symbols.shape[0] == nmodes
    modes = np.ascontiguousarray(np.atleast_1d(modes), dtype=np.int64)
    assert modes.max() < nmodes
    if method == "sbd_data":
        assert symbols.shape[1] >= TrSyms
    assert (TrSyms - 1) * os_ + ntaps <= L, "training would read past the end of the signal"
    err = np.zeros((nseg, nmodes, TrSyms * Niter), dtype=ct)
    mu_out = np.full(nseg, mu, dtype=rt)
    rc = getattr(lib(kind), "qo_train_equaliser" + suf)(
        _p(E), nseg, nmodes * L, L, nmodes, TrSyms, Niter, os_, rt(mu), _p(wx), ntaps, _p(modes),
        modes.size, int(bool(adaptive)), _p(symbols), symbols.shape[1], METHODS[method],
        int(bool(mu_shared)), _p(err), _p(mu_out))
    if rc:
        raise RuntimeError("oracle train_equaliser failed rc=%d" % rc)
    return err, wx, mu_out


def train_equaliser(E, TrSyms, Niter, os_, mu, wx, modes, adaptive, symbols, method,
                    mu_shared=True, kind="strict"):
    """pythran_equalisation.py:130-173.  wx is updated in place and returned."""
    assert wx.flags.c_contiguous
    err, _, mu_out = train_segments(E[None], TrSyms, Niter, os_, mu, wx[None], modes, adaptive,
                                    symbols, method, mu_shared, kind)
    return err[0], wx, mu_out[0]


def apply_segments(E, os_, wx, modes=None, kind="strict"):
    suf, rt, ct = _suf(E.dtype)
    E = np.ascontiguousarray(E, dtype=ct)
    wx = np.ascontiguousarray(wx, dtype=ct)
    nseg, nmodes, L = E.shape
    ntaps = wx.shape[-1]
    if modes is None:
        modes = np.arange(wx.shape[1])
    modes = np.ascontiguousarray(np.atleast_1d(modes), dtype=np.int64)
    N = max((L - ntaps + 1) // os_, 0)
    out = np.zeros((nseg, modes.size, N), dtype=ct)
    rc = getattr(lib(kind), "qo_apply_filter_to_signal" + suf)(
        _p(E), nseg, nmodes * L, L, nmodes, L, os_, _p(wx), ntaps, _p(modes), modes.size, _p(out))
    if rc:
        raise RuntimeError("oracle apply_filter_to_signal failed rc=%d" % rc)
    return out


def apply_filter_to_signal(E, os_, wx, modes=None, kind="strict"):
    """pythran_equalisation.py:37-76."""
    return apply_segments(E[None], os_, np.asarray(wx)[None], modes, kind)[0]


def bps_streams(E, testangles, symbols, N, kind="strict"):
    """pythran_dsp.py:47-85 for every row of E (nstream, L); returns int32 idx (nstream, L)."""
    suf, rt, ct = _suf(E.dtype)
    E = np.ascontiguousarray(E, dtype=ct)
    testangles = np.atleast_2d(np.asarray(testangles, dtype=rt))
    comp = np.ascontiguousarray(np.exp(1j * testangles))  # :72, complex of matching width (NumPy>=2)
    assert comp.dtype == ct
    symbols = np.ascontiguousarray(symbols, dtype=ct)
    nstream, L = E.shape
    idx = np.zeros((nstream, L), dtype=np.int32)
    rc = getattr(lib(kind), "qo_bps" + suf)(_p(E), nstream, L, L, _p(comp), comp.shape[0],
                                             comp.shape[1], _p(symbols), symbols.size, N, _p(idx))
    if rc:
        raise RuntimeError("oracle bps failed rc=%d" % rc)
    return idx


def bps(E, testangles, symbols, N, kind="strict"):
    return bps_streams(np.asarray(E)[None], testangles, symbols, N, kind)[0]


def select_angles(angles, idx, kind="strict"):
    """pythran_dsp.py:137-153."""
    suf, rt, ct = _suf(angles.dtype)
    angles = np.ascontiguousarray(np.atleast_2d(angles), dtype=rt)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    L = angles.shape[0] if angles.shape[0] > 1 else idx.shape[0]
    out = np.zeros(L, dtype=rt)
    getattr(lib(kind), "qo_select_angles" + suf)(_p(angles), angles.shape[0], angles.shape[1],
                                                 _p(idx), L, _p(out))
    return out


# --------------------------------------------------------------------------------------
# constellations and per-method constants
# --------------------------------------------------------------------------------------
def cal_symbols_qam(M):
    """theory.py:111-178 (square grid, real level slowest; cross constellations relocated)."""
    nb = int(round(np.log2(M)))
    if nb % 2 == 0:
        m = int(round(np.sqrt(M)))
        lv = np.linspace(-(m - 1), m - 1, m)
        return (lv[:, None] + 1j * lv[None, :]).flatten()
    n = (nb - 1) / 2
    s = 2 ** (n - 1)
    lr = np.linspace(-(2 ** (n + 1) - 1), 2 ** (n + 1) - 1, int(2 ** (n + 1)))
    li = np.linspace(-(2 ** n - 1), 2 ** n - 1, int(2 ** n))
    q = (lr[:, None] + 1j * li[None, :])
    far = abs(q.real) > 3 * s
    i1 = far & (abs(q.imag) > s)
    i2 = far & (abs(q.imag) <= s)
    q1 = np.sign(q.real) * (abs(q.real) - 2 * s) + 1j * np.sign(q.imag) * (4 * s - abs(q.imag))
    q2 = np.sign(q.real) * (4 * s - abs(q.real)) + 1j * np.sign(q.imag) * (abs(q.imag) + 2 * s)
    q = np.where(i1, q1, np.where(i2, q2, q))
    return q.flatten()


def cal_scaling_factor_qam(M):
    if int(round(np.log2(M))) % 2 == 0:
        return 2 / 3 * (M - 1)
    return (abs(cal_symbols_qam(M)) ** 2).mean()


def _norm_syms(M):
    return cal_symbols_qam(M) / np.sqrt(cal_scaling_factor_qam(M))


def generate_symbols_for_eq(method, M, dtype):
    """equalisation.py:101-136 (+ :271-281, :311-359)."""
    s = _norm_syms(M)
    if method in ("cma", "cma2", "sgncma"):
        R = np.mean(abs(s) ** 4) / np.mean(abs(s) ** 2)
        return np.atleast_2d(R + 0j).astype(dtype)
    if method == "mcma":
        R = np.mean(s.real ** 4) / np.mean(s.real ** 2) + 1j * np.mean(s.imag ** 4) / np.mean(s.imag ** 2)
        return np.atleast_2d(R).astype(dtype)
    if method == "rde":
        codes = np.unique(abs(s) ** 4 / abs(s) ** 2)
        parts = codes[:-1] + np.diff(codes) / 2
        return np.atleast_2d(np.hstack([codes, parts]) + 0j).astype(dtype)
    if method == "mrde":
        cr = np.unique(abs(s.real) ** 4 / abs(s.real) ** 2)
        ci = np.unique(abs(s.imag) ** 4 / abs(s.imag) ** 2)
        pr = cr[:-1] + np.diff(cr) / 2
        pi = ci[:-1] + np.diff(ci) / 2
        return np.atleast_2d(np.hstack([cr + 1j * ci, pr + 1j * pi])).astype(dtype)
    if method in ("sbd", "mddma", "dd"):
        return np.atleast_2d(s).astype(dtype)
    if method in DATA_AIDED:
        raise ValueError("%s is a data-aided method and needs the symbols to be passed" % method)
    raise ValueError("%s is unknown method" % method)


def _reshape_symbols(symbols, method, M, dtype, nmodes):
    """equalisation.py:568-594 (complex-valued methods only)."""
    if symbols is None or method in NONDECISION_BASED:
        symbols = generate_symbols_for_eq(method, M, dtype)
    symbols = np.asarray(symbols)
    if symbols.ndim == 1 or symbols.shape[0] == 1:
        symbols = np.tile(symbols, (nmodes, 1))
    elif symbols.shape[0] != nmodes:
        raise ValueError("Symbols array is shape {} but signal has {} modes".format(symbols.shape, nmodes))
    return np.atleast_2d(symbols.astype(dtype))


def cal_training_symbol_len(os_, ntaps, L):
    return int(L // os_ // ntaps - 1) * int(ntaps)  # equalisation.py:361-362


def init_taps(ntaps, nmodes, dtype):
    w = np.zeros((nmodes, nmodes, ntaps), dtype=dtype)  # equalisation.py:364-367
    for i in range(nmodes):
        w[i, i, ntaps // 2] = 1
    return w


# --------------------------------------------------------------------------------------
# L2 drivers
# --------------------------------------------------------------------------------------
def apply_filter(E, os_, wxy, modes=None, kind="strict"):
    """equalisation.py:138-188 (complex taps)."""
    E = np.copy(E)
    wxy = np.copy(wxy)
    modes = np.arange(wxy.shape[0]) if modes is None else np.copy(np.atleast_1d(modes))
    return apply_filter_to_signal(E, os_, wxy, modes, kind)


def equalise_signal(E, os_, mu, M, wxy=None, Ntaps=None, TrSyms=None, Niter=1, method="mcma",
                    adaptive_stepsize=False, symbols=None, modes=None, apply=False,
                    mu_shared=True, kind="strict", **kwargs):
    """equalisation.py:468-566."""
    method = method.lower()
    E = np.copy(np.asarray(E))
    mu = E.real.dtype.type(mu)
    nmodes = E.shape[0]
    if modes is None:
        modes = np.arange(nmodes)
    else:
        modes = np.atleast_1d(modes)
        assert np.max(modes) < nmodes, "largest mode number is larger than shape of signal"
    if wxy is None:
        wxy = init_taps(Ntaps, nmodes, E.dtype)
    else:
        wxy = np.ascontiguousarray(wxy, dtype=E.dtype)
        Ntaps = wxy.shape[-1]
        assert wxy.ndim == 3, "wxy needs to be three dimensional"
        assert wxy.shape[:2] == (nmodes, nmodes)
    if TrSyms is None:
        TrSyms = cal_training_symbol_len(os_, Ntaps, E.shape[-1])
    symbols = _reshape_symbols(symbols, method, M, E.dtype, nmodes)
    err, wxy, mu = train_equaliser(E, TrSyms, Niter, os_, mu, wxy, modes, adaptive_stepsize,
                                   symbols.copy(), method, mu_shared, kind)
    if apply:
        return apply_filter(E, os_, wxy, modes, kind), wxy, err
    return wxy, err


def dual_mode_equalisation(E, os_, mu, M, wxy=None, Ntaps=None, TrSyms=(None, None), Niter=(1, 1),
                           methods=("mcma", "sbd"), adaptive_stepsize=(False, False), symbols=None,
                           modes=None, apply=True, mu_shared=True, kind="strict", **kwargs):
    """equalisation.py:400-466: stage 2 restarts at sample 0 with the stage-1 taps; the output is
    the final taps applied to the whole signal."""
    symbols = np.atleast_1d(symbols)
    if symbols.ndim < 3:
        symbols = np.tile(symbols, (2, 1, 1))
    s0 = None if symbols[0].dtype == object else symbols[0]
    s1 = None if symbols[1].dtype == object else symbols[1]
    wxy, err1 = equalise_signal(E, os_, mu[0], M, wxy=wxy, Ntaps=Ntaps, TrSyms=TrSyms[0],
                                Niter=Niter[0], method=methods[0],
                                adaptive_stepsize=adaptive_stepsize[0], symbols=s0, modes=modes,
                                mu_shared=mu_shared, kind=kind)
    wxy2, err2 = equalise_signal(E, os_, mu[1], M, wxy=wxy, TrSyms=TrSyms[1], Niter=Niter[1],
                                 method=methods[1], adaptive_stepsize=adaptive_stepsize[1],
                                 symbols=s1, modes=modes, mu_shared=mu_shared, kind=kind)
    if apply:
        return apply_filter(E, os_, wxy2, modes, kind), wxy2, (err1, err2)
    return wxy2, (err1, err2)


def bps_driver(E, Mtestangles, symbols, N, kind="strict"):
    """phaserecovery.py:141-159: angle table, per-mode index search, unwrap*4/4 on [N:-N],
    rotate by exp(+1j*ph)."""
    E = np.asarray(E)
    dtype = np.float32 if E.dtype == np.dtype(np.complex64) else np.float64
    angles = np.linspace(-np.pi / 4, np.pi / 4, Mtestangles, endpoint=False, dtype=dtype).reshape(1, -1)
    Ew = np.atleast_2d(E).astype(E.dtype)
    ph = []
    for i in range(Ew.shape[0]):
        idx = bps(Ew[i], angles, symbols, N, kind)
        ph.append(select_angles(np.copy(angles), idx.astype(int), kind))
    ph = np.asarray(ph, dtype=dtype)
    ph[:, N:-N] = np.unwrap(ph[:, N:-N] * 4) / 4
    if E.ndim == 1:
        return (Ew * np.exp(1.j * ph)).flatten(), ph.flatten()
    return Ew * np.exp(1.j * ph), ph


def bps_twostage_driver(E, Mtestangles, symbols, N, B=4, kind="strict"):
    """phaserecovery.py:222-288: coarse search, per-symbol fine table (p == L form of bps), whole-array
    unwrap(4*ph, discont=pi)/4, rotate by exp(+1j*ph)."""
    E = np.asarray(E)
    rt = E.real.dtype
    angles = np.linspace(-np.pi / 4, np.pi / 4, Mtestangles, endpoint=False, dtype=rt).reshape(1, -1)
    Ew = np.atleast_2d(E)
    ph_out = []
    for i in range(Ew.shape[0]):
        idx = bps(np.copy(Ew[i]), angles, symbols, N, kind)
        ph = select_angles(np.copy(angles), idx, kind)
        b = np.linspace(-B / 2, B / 2, B)
        phn = (ph[:, np.newaxis] + b[np.newaxis, :] / (B * Mtestangles) * np.pi / 2).astype(rt)
        idx2 = bps(np.copy(Ew[i]), phn, symbols, N, kind)
        phf = select_angles(np.copy(phn), idx2, kind)
        ph_out.append(np.unwrap(phf * 4, discont=np.pi * 4 / 4) / 4)
    ph_out = np.asarray(ph_out, dtype=rt)
    En = Ew * np.exp(1.j * ph_out)
    if E.ndim == 1:
        return En.flatten(), ph_out.flatten()
    return En, ph_out


# --------------------------------------------------------------------------------------
# real-valued trainer: pure-Python/NumPy restatement, small cases only
# --------------------------------------------------------------------------------------
def _adapt_step_real(mu, err_p, err):
    """pythran_equalisation.py:18-22"""
    if err * err_p > 0:
        return mu
    return mu / (1 + mu * (err * err))


def train_equaliser_realvalued(E, TrSyms, Niter, os_, mu, wx, modes, adaptive, symbols, method):
    """pythran_equalisation.py:80-128, line by line (NumPy scalar semantics of the interpreted reference:
    ``abs(x)**2``, ``np.sign``, ``np.argmin`` of ``np.abs``).  ``wx`` is updated in place."""
    def cma(Xest, s1, i):
        return (s1[0] - abs(Xest) ** 2) * Xest                       # :113-115

    def sgncma(Xest, s1, i):
        return np.sign(s1[0] - abs(Xest) ** 2) * np.sign(Xest)       # :117-119

    def dd(Xest, symbs, i):
        dist = np.abs(Xest - symbs)                                  # det_symbol_argmin :232-235
        symbol = symbs[np.argmin(dist)]
        return (symbol - Xest) * abs(symbol)                         # :121-123

    def dd_data(Xest, symbs, i):
        symbol = symbs[i]
        return (symbol - Xest) * abs(symbol)                         # :125-128

    fct = {"cma": cma, "sgncma": sgncma, "dd": dd, "dd_data": dd_data}
    if method not in fct:
        raise ValueError("Unknown method %s" % method)
    errorfct = fct[method]
    nmodes = E.shape[0]
    ntaps = wx.shape[-1]
    err = np.zeros((nmodes, TrSyms * Niter), dtype=E.dtype)
    for mode in modes:
        for it in range(Niter):
            for i in range(TrSyms):
                X = E[:, i * os_:i * os_ + ntaps]
                Xest = E.dtype.type(0)
                for k in range(nmodes):                              # apply_filter :24-31
                    for t in range(ntaps):
                        Xest += X[k, t] * wx[mode, k, t]
                err[mode, it * TrSyms + i] = errorfct(Xest, symbols[mode], i)
                wx[mode] += mu * err[mode, it * TrSyms + i] * X
                if adaptive and i > 0:
                    mu = _adapt_step_real(mu, err[mode, it * TrSyms + i], err[mode, it * TrSyms + i - 1])
    return err, wx, mu


def equalise_signal_real(E, os_, mu, M, Ntaps, method, TrSyms=None, Niter=1, adaptive=False, symbols=None,
                         modes=None, apply=True):
    """equalisation.py:529-565 for the REAL_VALUED methods (cma_real, sgncma_real, dd_real, dd_data_real)."""
    from qampy_b200 import theory
    Er = theory.convert_sig_to_real(np.atleast_2d(E))
    mu = Er.dtype.type(mu)
    nmodes = Er.shape[0]
    if modes is None:
        modes = np.arange(nmodes)
    else:
        modes = np.atleast_1d(modes)
        modes = np.hstack([modes, modes + nmodes // 2])
    wxy = theory.init_taps(Ntaps, nmodes, Er.dtype)
    if TrSyms is None:
        TrSyms = theory.cal_training_symbol_len(os_, Ntaps, Er.shape[-1])
    symbols = theory.reshape_symbols(symbols, method, M, Er.dtype, nmodes)
    err, wxy, mu = train_equaliser_realvalued(Er, TrSyms, Niter, os_, mu, wxy, modes, adaptive, symbols.copy(),
                                              method[:-5])
    if not apply:
        return wxy, err
    ct = np.complex64 if Er.dtype == np.float32 else np.complex128
    out = apply_filter_to_signal(Er.astype(ct), os_, wxy.astype(ct), modes).real
    Im = np.complex64(1j) if Er.itemsize == 4 else np.complex128(1j)
    return theory.convert_sig_to_cmplx(out, modes.shape[0], Im), wxy, err


# --------------------------------------------------------------------------------------
# decisions and quality metrics: NumPy restatements
# --------------------------------------------------------------------------------------
def make_decision(E, symbols):
    """pythran_equalisation.py:306-334 with det_symbol_argmin (:232-235), vectorised in chunks."""
    E = np.asarray(E)
    symbols = np.asarray(symbols, dtype=E.dtype)
    det = np.zeros_like(E)
    dist = np.zeros(E.shape, dtype=E.real.dtype)
    idx = np.zeros(E.shape, dtype=np.int32)
    for a in range(0, E.shape[0], 16384):
        d = np.abs(E[a:a + 16384, None] - symbols[None, :])
        ix = np.argmin(d, axis=1)
        idx[a:a + 16384] = ix
        dist[a:a + 16384] = d[np.arange(ix.size), ix]
        det[a:a + 16384] = symbols[ix]
    return det, dist, idx


def soft_l_value_demapper(rx_symbs, num_bits, snr, bits_map, minmax=False):
    """pythran_dsp.py:95-108 (exact) and :110-131 (max-log), evaluated in the signal's precision."""
    rx = np.asarray(rx_symbs)
    rt = rx.real.dtype.type
    out = np.zeros((rx.shape[0], num_bits))
    for bit in range(num_bits):
        d1 = np.abs(bits_map[bit, :, 1][None, :] - rx[:, None]) ** 2
        d0 = np.abs(bits_map[bit, :, 0][None, :] - rx[:, None]) ** 2
        if minmax:
            out[:, bit] = rt(snr) * (np.minimum(d0.min(axis=1), rt(10000.)) - np.minimum(d1.min(axis=1), rt(10000.)))
        else:
            out[:, bit] = np.log(np.sum(np.exp(-rt(snr) * d1), axis=1)) - np.log(np.sum(np.exp(-rt(snr) * d0), axis=1))
    return out


def estimate_snr(signal_rx, symbols_tx, gray_symbols):
    """pythran_dsp.py:244-286."""
    L = signal_rx.shape[0]
    in_pow, N0 = 0., 0.
    for g in gray_symbols:
        sel = signal_rx[symbols_tx == g]
        K = sel.shape[0]
        Px = K / L
        mu = np.mean(sel)
        sigma = np.sqrt(np.sum(abs(sel - mu) ** 2) / K)
        N0 += abs(sigma) ** 2 * Px
        in_pow += abs(mu) ** 2 * Px
    return in_pow / N0, in_pow, N0


def viterbiviterbi(E, N, M):
    """qampy/core/phaserecovery.py:40-79 restated in NumPy, every step in the dtype the reference uses (phase and
    phasors in the signal's precision, window sum over a strided view).  Returns (Eout, phases of every row)."""
    E2d = np.atleast_2d(np.asarray(E))
    Eout = np.zeros_like(E2d)
    L = E2d.shape[1]
    o = (N - 1) // 2
    phases = []
    for i in range(E2d.shape[0]):
        raised = np.exp(1.j * np.angle(E2d[i])) ** M                                   # :63-64
        win = np.lib.stride_tricks.sliding_window_view(raised, N)                     # segment_axis(., N, N-1), :65
        est = (np.unwrap(np.angle(np.sum(win, axis=1))) - np.pi) / M                  # :66-68
        Eout[i, o:o + L - N + 1] = E2d[i, o:o + L - N + 1] * np.exp(-1.j * est)      # :69-72
        phases.append(est)
    return Eout, np.asarray(phases)
