"""Drop-in for the reference's L1 blind-phase-search kernels (``qampy/core/pythran_dsp.py``):

* ``bps(E, testangles, symbols, N)`` -> int32 angle indices   (:45-85 incl. select_angle_index :26-42)
* ``select_angles(angles, idx)``                              (:133-153)

NumPy in / NumPy out through the C ABI's ``*_host`` entry points (CUDA only, no CPU path).
"""
import ctypes

import numpy as np

from . import _lib
from .pythran_equalisation import _ctype, _p


def _rtype(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return _lib.QB_C64, np.float32
    if dtype == np.float64:
        return _lib.QB_C128, np.float64
    raise TypeError("angles must be float32 or float64, got %s" % dtype)


def bps(E, testangles, symbols, N):
    """Blind phase search index search for one 1-D signal.  ``testangles`` is (1, A) -- one table for
    every symbol -- or (L, A), one row of test angles per symbol (the second stage of two-stage BPS,
    ``phaserecovery.py:276-281``; ``pythran_dsp.py:74-77``)."""
    code, rt, ct = _ctype(np.asarray(E).dtype)
    E = np.ascontiguousarray(E, dtype=ct)
    if E.ndim != 1:
        raise ValueError("E must be 1-dimensional")
    testangles = np.atleast_2d(np.asarray(testangles, dtype=rt))
    p, A = testangles.shape
    if p != 1 and p != E.shape[0]:
        raise ValueError("p must be either 1 or the length of the input signal")   # pythran_dsp.py:69
    comp = np.ascontiguousarray(np.exp(1j * testangles), dtype=ct)      # pythran_dsp.py:72 (NumPy's table)
    symbols = np.ascontiguousarray(symbols, dtype=ct).reshape(-1)
    idx = np.zeros(E.shape[0], dtype=np.int32)
    fn = _lib.load().qb_bps_host if p == 1 else _lib.load().qb_bps_rows_host
    _lib.check(fn(code, _p(E), 1, E.shape[0], _p(comp), None, A, _p(symbols), symbols.size, int(N), _p(idx),
                  None, None))
    return idx


def select_angles(angles, idx):
    angles = np.atleast_2d(np.asarray(angles))
    code, rt = _rtype(angles.dtype)
    angles = np.ascontiguousarray(angles, dtype=rt)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    p, A = angles.shape
    L = p if p > 1 else idx.shape[0]
    assert idx.shape[0] >= L
    out = np.zeros(L, dtype=rt)
    _lib.check(_lib.load().qb_select_angles_host(code, _p(angles), p, A, _p(idx), L, _p(out)))
    return out


def _demap(rx_symbs, num_bits, snr, bits_map, minmax):
    code, rt, ct = _ctype(np.asarray(rx_symbs).dtype)
    rx = np.ascontiguousarray(rx_symbs, dtype=ct)
    if rx.ndim != 1:
        raise ValueError("rx_symbs must be 1-dimensional")
    bm = np.ascontiguousarray(bits_map, dtype=ct)
    assert bm.ndim == 3 and bm.shape[2] == 2 and bm.shape[0] >= num_bits
    out = np.zeros((rx.shape[0], int(num_bits)))
    _lib.check(_lib.load().qb_soft_l_value_demapper_host(code, _p(rx), rx.shape[0], int(num_bits), float(snr), _p(bm),
                                                         bm.shape[0], bm.shape[1], int(minmax), _p(out)))
    return out


def soft_l_value_demapper(rx_symbs, num_bits, snr, bits_map):
    """Exact log-likelihood ratios per bit (:95-108): ``bits_map[bit, :, b]`` are the alphabet points whose
    ``bit`` equals ``b``.  Returns float64 (N, num_bits) like the reference."""
    return _demap(rx_symbs, num_bits, snr, bits_map, False)


def soft_l_value_demapper_minmax(rx_symbs, num_bits, snr, bits_map):
    """Max-log approximation of the log-likelihood ratios (:110-131)."""
    return _demap(rx_symbs, num_bits, snr, bits_map, True)


def estimate_snr(signal_rx, symbols_tx, gray_symbols):
    """SNR from received and known transmitted symbols (:244-286).  Returns ``(snr, S0, N0)`` (linear)."""
    code, rt, ct = _ctype(np.asarray(signal_rx).dtype)
    rx = np.ascontiguousarray(signal_rx, dtype=ct)
    tx = np.ascontiguousarray(symbols_tx, dtype=ct)
    assert rx.shape[0] >= tx.shape[0]
    if rx.shape != tx.shape:
        raise ValueError("signal_rx and symbols_tx need to have the same length")
    gray = np.ascontiguousarray(gray_symbols, dtype=ct).reshape(-1)
    out = np.zeros(3)
    _lib.check(_lib.load().qb_estimate_snr_host(code, _p(rx), _p(tx), rx.shape[0], _p(gray), gray.size, _p(out)))
    return out[0], out[1], out[2]
