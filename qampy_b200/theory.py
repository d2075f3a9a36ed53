"""Constellations and per-method equaliser constants (host side, NumPy).

These are the numbers the kernels are parameterised with; they have to equal what the reference
computes in ``qampy/theory.py:111-178`` (``cal_symbols_qam``, ``cal_scaling_factor_qam``) and
``qampy/core/equalisation/equalisation.py:101-136, 271-281, 311-359`` (``generate_symbols_for_eq``
and the radius / partition tables).  ``tests/test_host_logic.py`` checks them against values dumped
from the reference (``tests/golden/g0_constants.npz``).
"""
import numpy as np

#: qampy/core/equalisation/equalisation.py:86-99
DECISION_BASED = ("sbd", "mddma", "dd", "sbd_data", "dd_real", "dd_data_real")
NONDECISION_BASED = ("cma", "cma2", "mcma", "rde", "mrde", "cma_real", "sgncma_real", "sgncma")
REAL_VALUED = ("cma_real", "dd_real", "dd_data_real", "sgncma_real")
DATA_AIDED = ("dd_data_real", "sbd_data")
TRAINING_FCTS = DECISION_BASED + NONDECISION_BASED


def _is_square(M):
    nb = np.log2(M)
    return not (nb % 2 > 0.5)


def cal_symbols_qam(M):
    """Unnormalised M-QAM grid in the reference's order (real level slowest)."""
    if _is_square(M):
        m = int(round(np.sqrt(M)))
        lv = np.arange(-(m - 1), m, 2, dtype=np.float64)
        return (lv[:, None] + 1j * lv[None, :]).reshape(-1)
    # cross constellations (32, 128, ...): a 2^(n+1) x 2^n rectangle whose outer columns are folded
    # onto the top and bottom (theory.py:161-178)
    n = (np.log2(M) - 1) / 2
    s = 2 ** (n - 1)
    lr = np.arange(-(2 ** (n + 1) - 1), 2 ** (n + 1), 2, dtype=np.float64)
    li = np.arange(-(2 ** n - 1), 2 ** n, 2, dtype=np.float64)
    re, im = np.meshgrid(lr, li, indexing="ij")
    outer = np.abs(re) > 3 * s
    hi = outer & (np.abs(im) > s)
    lo = outer & (np.abs(im) <= s)
    re2 = np.where(hi, np.sign(re) * (np.abs(re) - 2 * s), np.where(lo, np.sign(re) * (4 * s - np.abs(re)), re))
    im2 = np.where(hi, np.sign(im) * (4 * s - np.abs(im)), np.where(lo, np.sign(im) * (np.abs(im) + 2 * s), im))
    return (re2 + 1j * im2).reshape(-1)


def cal_scaling_factor_qam(M):
    if _is_square(M):
        return 2 / 3 * (M - 1)
    return (np.abs(cal_symbols_qam(M)) ** 2).mean()


def normalised_symbols(M):
    return cal_symbols_qam(M) / np.sqrt(cal_scaling_factor_qam(M))


def cal_Rconstant(M):
    s = normalised_symbols(M)
    return np.mean(np.abs(s) ** 4) / np.mean(np.abs(s) ** 2)


def cal_Rconstant_complex(M):
    s = normalised_symbols(M)
    return np.mean(s.real ** 4) / np.mean(s.real ** 2) + 1j * np.mean(s.imag ** 4) / np.mean(s.imag ** 2)


def partition_codes_radius(M):
    """RDE: ring radii (squared) and the decision boundaries half-way between them."""
    s = normalised_symbols(M)
    codes = np.unique(np.abs(s) ** 4 / np.abs(s) ** 2)
    return np.hstack([codes, codes[:-1] + np.diff(codes) / 2])


def partition_codes_complex(M):
    """MRDE: per-axis squared levels and boundaries, packed as complex (real axis, imaginary axis)."""
    s = normalised_symbols(M)
    cr = np.unique(np.abs(s.real) ** 4 / np.abs(s.real) ** 2)
    ci = np.unique(np.abs(s.imag) ** 4 / np.abs(s.imag) ** 2)
    codes = cr + 1j * ci
    parts = (cr[:-1] + np.diff(cr) / 2) + 1j * (ci[:-1] + np.diff(ci) / 2)
    return np.hstack([codes, parts])


def generate_symbols_for_eq(method, M, dtype):
    """Per-method constant table, shape (1, K) (equalisation.py:101-136)."""
    if method in ("cma", "cma2", "sgncma"):
        return np.atleast_2d(cal_Rconstant(M) + 0j).astype(dtype)
    if method == "mcma":
        return np.atleast_2d(cal_Rconstant_complex(M)).astype(dtype)
    if method == "rde":
        return np.atleast_2d(partition_codes_radius(M) + 0j).astype(dtype)
    if method == "mrde":
        return np.atleast_2d(partition_codes_complex(M)).astype(dtype)
    if method in ("sbd", "mddma", "dd"):
        return np.atleast_2d(normalised_symbols(M)).astype(dtype)
    if method in ("sgncma_real", "cma_real"):          # equalisation.py:126-129 (dtype is the REAL dtype here)
        return np.repeat([np.atleast_1d(cal_Rconstant_complex(M).real.astype(dtype))], 2, axis=0)
    if method == "dd_real":                            # :130-133
        symbols = normalised_symbols(M)
        return np.vstack([symbols.real, symbols.imag]).astype(dtype)
    if method in DATA_AIDED:
        raise ValueError("%s is a data-aided method and needs the symbols to be passed" % method)
    raise ValueError("%s is unknown method" % method)


SEARCHED_ALPHABET = ("sbd", "mddma", "dd")      # methods whose ``symbols`` are searched by det_symbol


def unique_alphabet(symbols, method):
    """For the methods that search their alphabet (``det_symbol``, pythran_equalisation.py:240-265): drop repeated
    points of every row, keeping first occurrences in order.  ``det_symbol`` returns the VALUE of the first strict
    minimum, and a repeat of an earlier point can never be a strict improvement, so the decisions -- and with them
    every output -- are unchanged; the search gets as short as the alphabet really is.  This matters where a whole
    training sequence is passed as ``symbols`` of such a method: the pilot equaliser hands ``sbd`` its 1024-symbol
    QPSK pilot sequence (pilotbased_receiver.py:530-541), i.e. 4 distinct points.  Rows keep a common length (short
    rows are padded with their own first point, which never wins a strict comparison against itself)."""
    if method not in SEARCHED_ALPHABET:
        return symbols
    symbols = np.atleast_2d(symbols)
    rows = []
    for r in symbols:
        _, first = np.unique(r, return_index=True)
        rows.append(r[np.sort(first)])
    K = max(len(r) for r in rows)
    if K == symbols.shape[1]:
        return symbols
    return np.stack([np.concatenate([r, np.repeat(r[:1], K - len(r))]) for r in rows])


def reshape_symbols(symbols, method, M, dtype, nmodes):
    """Bring user symbols to (nmodes, K) (equalisation.py:568-594).  For the real-valued methods
    ``nmodes`` counts the real rows (2 per polarisation) and ``dtype`` is the real dtype."""
    if symbols is None or method in NONDECISION_BASED:
        symbols = generate_symbols_for_eq(method, M, dtype)
    symbols = np.asarray(symbols)
    if method not in REAL_VALUED:
        if symbols.ndim == 1 or symbols.shape[0] == 1:
            symbols = np.tile(symbols, (nmodes, 1))
        elif symbols.shape[0] != nmodes:
            raise ValueError("Symbols array is shape {} but signal has {} modes, symbols must be 1d or of shape "
                             "(1, N) or ({}, N)".format(symbols.shape, nmodes, nmodes))
        return np.atleast_2d(symbols.astype(dtype))
    if np.iscomplexobj(symbols):                        # :579-586
        if symbols.ndim == 1 or symbols.shape[0] == 1:
            symbols = np.repeat([symbols.real, symbols.imag], nmodes // 2, axis=0).squeeze()
            symbols = symbols.reshape(nmodes, -1)
        elif symbols.shape[0] == nmodes // 2:
            symbols = np.vstack([symbols.real, symbols.imag])
        else:
            raise ValueError("Symbols array is  complex and has {} modes, but needs to either have one mode or "
                             "the same modes as the signal ({})".format(symbols.shape[0], nmodes // 2))
    else:                                               # :587-592
        if symbols.shape[0] == 2 and nmodes > 2:
            symbols = np.repeat([symbols[0], symbols[1]], nmodes // 2, axis=0).squeeze()
            symbols = symbols.reshape(nmodes, -1)
        elif symbols.shape[0] != nmodes:
            raise ValueError("Symbols array is shape {} but signal has {} modes, symbols must be 1d or of shape "
                             "(1, N) or ({}, N)".format(symbols.shape, nmodes, nmodes))
    return symbols.astype(dtype)


def convert_sig_to_real(E):
    """(nmodes, L) complex -> (2*nmodes, L) real: all real parts, then all imaginary parts (equalisation.py:253-257)."""
    Etmp = np.zeros((2 * E.shape[0], E.shape[1]), dtype=E.real.dtype)
    Etmp[:E.shape[0]] = E.real
    Etmp[E.shape[0]:] = E.imag
    return np.ascontiguousarray(Etmp)


def convert_sig_to_cmplx(E, modes, Im=np.complex128(1j)):
    """equalisation.py:259-260"""
    return E[:modes // 2, :] + Im * E[modes // 2:, :]


def cal_training_symbol_len(os, ntaps, L):
    return int(L // os // ntaps - 1) * int(ntaps)


def init_taps(ntaps, nmodes, dtype):
    """Centre-spike identity taps (equalisation.py:364-367)."""
    wxy = np.zeros((nmodes, nmodes, ntaps), dtype=dtype)
    for i in range(nmodes):
        wxy[i, i, ntaps // 2] = 1
    return wxy


def bps_test_angles(Mtestangles, dtype):
    """phaserecovery.py:145"""
    return np.linspace(-np.pi / 4, np.pi / 4, Mtestangles, endpoint=False, dtype=dtype).reshape(1, -1)
