"""Signal synthesis on the GPU with the reference generators' names, arguments and results:

* :func:`rrcos_resample`       ``qampy/core/resample.py:73-126`` (``fftconv=True`` branch: zero insertion, convolution
  with ``taps`` samples of the root-raised-cosine impulse response normalised to a peak of 1, "same" crop, decimation,
  optional re-normalisation to the input's power)
* :func:`apply_PMD_to_field`   ``qampy/core/impairments.py:94-131``
* :func:`add_awgn` / :func:`change_snr`          ``core/impairments.py:188-233``
* :func:`phase_noise` / :func:`apply_phase_noise``core/impairments.py:133-186``
* :func:`synth_signal` composes them into the receiver input of the BASELINE configs.

Tensors in, tensors out, everything on the current CUDA device.  cuFFT (through ``torch.fft``) does the transforms --
library work, as the task allows for plain FFTs; every other stage is one of this project's kernels
(``csrc/synth_ops.cu``): zero insertion, spectrum x tap spectrum, crop + decimate + re-centre + re-scale, the 2x2
PMD operator with its phase ramp, and a single tail pass for noise, Wiener phase walk (own three-kernel scan) and the
cast to the signal dtype.  Random deviates come from a counter-based generator inside the kernels, keyed by
``(seed, row, sample index)``: statistics match the reference's ``np.random`` calls, sample values cannot (different
generator), so parity is pinned on the deterministic stages (``tests/golden/g13_synth.npz``, <= 1e-9 rms against
the reference on fixed symbols) and the random stages are tested on their moments.
"""
import ctypes
import fractions
import math

import numpy as np
import torch

from . import _lib
from .device import _ptr, _stream
from .theory import normalised_symbols


def _resampling_factors(fold, fnew):
    r = fractions.Fraction(fnew / fold).limit_denominator()
    return r.numerator, r.denominator


def rrcos_impulse(t, beta, T):
    """Root-raised-cosine impulse response h(t) (unit-energy convention 1/T at the peak for beta = 0), with the two
    removable singularities t = 0 and |t| = T / (4 beta) replaced by their limits.  float64 NumPy."""
    t = np.asarray(t, dtype=np.float64)
    x = t / T
    with np.errstate(divide="ignore", invalid="ignore"):
        num = np.sin(np.pi * x * (1 - beta)) + 4 * beta * x * np.cos(np.pi * x * (1 + beta))
        den = np.pi * x * (1 - (4 * beta * x) ** 2)
        h = num / den / T
    tol = abs(t[1] - t[0]) / 4 if t.size > 1 else 1e-12
    h[np.abs(t) < tol] = (1 + beta * (4 / np.pi - 1)) / T
    if beta > 0:
        edge = np.abs(np.abs(t) - T / (4 * beta)) < tol
        h[edge] = beta / (T * math.sqrt(2)) * ((1 + 2 / np.pi) * math.sin(np.pi / (4 * beta)) +
                                               (1 - 2 / np.pi) * math.cos(np.pi / (4 * beta)))
    return h


def _pulse_taps(taps, fs, T, beta):
    """The tap vector of ``rrcos_pulseshaping`` (core/filter.py:201-205): ``taps`` samples centred on (taps-1)//2, peak 1."""
    k = np.arange(taps, dtype=np.float64)
    k -= k[(taps - 1) // 2]
    h = rrcos_impulse(k / fs, beta, T)
    return h / h.max()


def _check(x):
    if not (torch.is_tensor(x) and x.is_cuda):
        raise _lib.QampyB200Error("qampy_b200.synth_device works on CUDA tensors (no CPU fallback)")


def rrcos_resample(signal, fold, fnew, Ts=None, beta=None, taps=4001, renormalise=False):
    """``signal`` (nrows, n) or (n,) complex CUDA tensor sampled at ``fold`` -> pulse-shaped signal at ``fnew``
    (complex128, ``n * fnew / fold`` samples per row).  Rows are independent (the reference is called per mode)."""
    _check(signal)
    if beta is None:
        raise NotImplementedError("beta=None (polyphase resampling without pulse shaping) is not part of this path")
    assert 0 < beta <= 1, "beta needs to be in interval (0,1]"
    one_d = signal.dim() == 1
    x = (signal.unsqueeze(0) if one_d else signal).to(torch.complex128).contiguous()
    rows, n = x.shape
    if Ts is None:
        Ts = 1 / fold
    up, down = _resampling_factors(fold, fnew)
    n_up = n * up
    h = _pulse_taps(taps, up * fold, Ts, beta)
    # linear convolution by FFT: any length >= n_up + taps - 1 (5-smooth lengths keep cuFFT on its fast paths)
    from scipy.fft import next_fast_len
    nfft = int(next_fast_len(n_up + taps - 1))
    lib = _lib.load()
    X = torch.empty((rows, nfft), dtype=torch.complex128, device=x.device)
    _lib.check(lib.qb_synth_upsample_dev(_ptr(x), rows, n, up, _ptr(X), nfft, _stream()))
    hp = np.zeros(nfft, dtype=np.complex128)
    hp[:taps] = h
    H = torch.fft.fft(torch.from_numpy(hp).to(x.device))
    Xf = torch.fft.fft(X, dim=1)
    _lib.check(lib.qb_synth_specmul_dev(_ptr(Xf), rows, nfft, _ptr(H), _stream()))
    y = torch.fft.ifft(Xf, dim=1)
    n_out = (n_up + down - 1) // down
    out = torch.empty((rows, n_out), dtype=torch.complex128, device=x.device)
    power = (x.real ** 2 + x.imag ** 2).mean(dim=1).contiguous() if renormalise else None
    _lib.check(lib.qb_synth_crop_norm_dev(_ptr(y), rows, nfft, (taps - 1) // 2, down, n_out, _ptr(power),
                                          int(bool(renormalise)), _ptr(out), _stream()))
    return out[0] if one_d else out


def apply_PMD_to_field(field, theta, t_dgd, fs):
    """First-order PMD on a dual-polarisation field (2, n): principal axes at ``theta``, differential group delay
    ``t_dgd``.  Returns a tensor of the field's dtype."""
    _check(field)
    assert field.dim() == 2 and field.shape[0] == 2, "PMD acts on a dual-polarisation field"
    S = torch.fft.fft(field.to(torch.complex128), dim=1).contiguous()
    _lib.check(_lib.load().qb_synth_pmd_dev(_ptr(S), S.shape[1], float(theta), float(t_dgd), float(fs), _stream()))
    return torch.fft.ifft(S, dim=1).to(field.dtype)


def _tail(x, noise_sigma, walk_sigma, seed, dtype, row0=0, index0=0, phase0=None, want_phase=False):
    one_d = x.dim() == 1
    x2 = (x.unsqueeze(0) if one_d else x).to(torch.complex128).contiguous()
    rows, n = x2.shape
    out = torch.empty((rows, n), dtype=dtype, device=x2.device)
    ns = None
    if noise_sigma is not None:
        ns = torch.as_tensor(np.broadcast_to(np.asarray(noise_sigma, dtype=np.float64).reshape(-1), (rows,)).copy(),
                             device=x2.device)
    p0 = None if phase0 is None else torch.as_tensor(np.asarray(phase0, dtype=np.float64).reshape(rows).copy(),
                                                     device=x2.device)
    ph = torch.empty((rows, n), dtype=torch.float64, device=x2.device) if want_phase else None
    code = _lib.QB_C64 if dtype == torch.complex64 else _lib.QB_C128
    _lib.check(_lib.load().qb_synth_tail_dev(code, _ptr(x2), rows, n, _ptr(ns), float(walk_sigma),
                                             ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(row0), int(index0),
                                             _ptr(p0), _ptr(out), n, _ptr(ph), _stream()))
    if one_d:
        out, ph = out[0], (ph[0] if ph is not None else None)
    return (out, ph) if want_phase else out


def add_awgn(sig, strgth, seed=0):
    """sig + strgth (N(0,1) + i N(0,1)) / sqrt(2) per sample; ``strgth`` a scalar or one value per row."""
    _check(sig)
    return _tail(sig, strgth, 0.0, seed, sig.dtype if sig.dtype in (torch.complex64, torch.complex128) else torch.complex128)


def change_snr(sig, snr, fb, fs, seed=0):
    """Noise for a per-symbol SNR of ``snr`` dB on a noiseless signal oversampled ``fs / fb`` times: std =
    sqrt(mean power) 10^(-snr/20) sqrt(os).  The reference takes ONE mean over the whole array; so does this."""
    _check(sig)
    p = float((sig.real.double() ** 2 + sig.imag.double() ** 2).mean())
    return add_awgn(sig, math.sqrt(p) * 10 ** (-snr / 20) * math.sqrt(fs / fb), seed=seed)


def phase_noise(sz, df, fs, seed=0, device=None):
    """Wiener phase walk(s) of shape ``sz`` with step variance 2 pi df / fs (float64 tensor)."""
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    sz = (sz,) if np.isscalar(sz) else tuple(sz)
    zero = torch.zeros(sz, dtype=torch.complex128, device=dev)
    _, ph = _tail(zero, None, math.sqrt(2 * math.pi * df / fs), seed, torch.complex128, want_phase=True)
    return ph


def apply_phase_noise(signal, df, fs, seed=0):
    """signal * exp(i phase), phase a Wiener walk per row with step variance 2 pi df / fs."""
    _check(signal)
    return _tail(signal, None, math.sqrt(2 * math.pi * df / fs), seed, signal.dtype)


def synth_signal(M, nsym, nmodes=2, os=2, beta=0.1, snr_db=28.0, theta=math.pi / 5.6, dgd=40e-12, fb=40e9,
                 linewidth=None, seed=0, dtype=torch.complex64, device=None, taps=4001):
    """The receiver input of the BASELINE configs made with the generators above: random M-QAM symbols ->
    rrcos_resample(fb -> os fb, renormalise) -> apply_PMD_to_field -> change_snr + apply_phase_noise (one tail pass).
    Returns (E (nmodes, nsym*os) of ``dtype``, symbols (nmodes, nsym))."""
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    alphabet = torch.from_numpy(normalised_symbols(M)).to(dev)
    syms = alphabet[torch.randint(0, M, (nmodes, nsym), generator=gen, device=dev)]
    x = rrcos_resample(syms, fb, os * fb, beta=beta, taps=taps, renormalise=True)
    if nmodes == 2 and theta is not None:
        x = apply_PMD_to_field(x, theta, dgd, os * fb)
    p = float((x.real ** 2 + x.imag ** 2).mean())
    sigma = None if snr_db is None else math.sqrt(p) * 10 ** (-snr_db / 20) * math.sqrt(os)
    walk = 0.0 if not linewidth else math.sqrt(2 * math.pi * linewidth / (os * fb))
    return _tail(x, sigma, walk, seed, dtype), syms.to(dtype)
