"""Estimators around the pilot-based receiver: spectral frequency-offset search, sequence location by
cross-correlation, pilot phase / frequency estimates.

Same names, arguments and results as the reference's helpers (``phaserecovery.find_freq_offset`` /
``comp_freq_offset`` :385-473, ``ber_functions.find_sequence_offset[_complex]`` :33-106, ``filter.moving_average``
:215-237, ``pilotbased_receiver.pilot_based_foe`` :32-73, ``pilot_based_cpe_new`` :258-327, ``correct_shifts``
:436-443) so that the parity tests read like the reference's, but formulated for this package:

* every function takes NumPy arrays (host) or CUDA tensors (``torch``): with tensors the FFTs run in cuFFT and the
  element-wise parts in the tensor's own device -- ``pilots.pilot_receiver`` keeps the capture resident and only a few
  scalars come back;
* the four quarter-turn hypotheses of the sequence search share ONE cross-correlation (``corr(x, y i^k) =
  (-i)^k corr(x, y)``, and multiplying by a unit of the Gaussian integers is exact in floating point) instead of four;
* the per-mode loops of the reference are array expressions here; where a result must agree with the reference bit for
  bit (pilot phase trace, fitted frequency offsets: ``tests/test_pilots.py``) the floating-point operations are the
  reference's, in its order.
"""
import warnings

import numpy as np

try:                      # torch is only needed for the device forms
    import torch
except Exception:         # pragma: no cover
    torch = None


def _is_tensor(a):
    return torch is not None and torch.is_tensor(a)


def _pow2_at_least(n):
    n = int(n)
    return n if n > 0 and n & (n - 1) == 0 else 1 << max(n - 1, 1).bit_length()


# ---------------------------------------------------------------------------------------------------------
# frequency offset
# ---------------------------------------------------------------------------------------------------------
def find_freq_offset(sig, os=1, average_over_modes=True, fft_size=2 ** 16):
    """Frequency offset of an M-QAM signal from the strongest line of the spectrum of its fourth power.

    ``sig``: (nmodes, L) or (L,), NumPy array or CUDA tensor; ``fft_size`` is rounded up to a power of two; the
    spectrum is that of the first ``fft_size`` samples (zero padded when shorter).  Returns a float64 NumPy array
    (nmodes, 1), in units of the symbol rate when ``os`` is the oversampling of ``sig`` -- all rows equal to the mean
    over the modes when ``average_over_modes``."""
    n = _pow2_at_least(fft_size)
    if _is_tensor(sig):
        s = sig if sig.dim() == 2 else sig.unsqueeze(0)
        s2 = s * s
        spec = torch.fft.fft(s2 * s2, n=n, dim=-1)
        peak = torch.argmax(spec.real.double() ** 2 + spec.imag.double() ** 2, dim=-1).cpu().numpy()
    else:
        s = np.atleast_2d(sig)
        s2 = s * s
        spec = np.fft.fft(s2 * s2, n, axis=-1)
        # squared magnitude in the signal's precision, compared in double (the reference's |.|^2 cast to float64)
        peak = np.argmax((np.abs(spec) ** 2).astype(np.float64), axis=-1)
    lines = np.fft.fftfreq(n, 1 / os) / 4          # a line at 4 f of the fourth power <-> an offset f
    offs = lines[peak].reshape(-1, 1)
    if average_over_modes:
        offs = np.full(offs.shape, np.mean(offs))
    return offs


def comp_freq_offset(sig, freq_offset, os=1):
    """De-rotate a frequency offset: row r is multiplied by exp(-2 pi i t f_r / os), t = 1 .. L (one-based, like the
    reference).  NumPy in, NumPy out (the device form is ``device.freq_shift``, a CUDA kernel)."""
    one_d = np.ndim(sig) == 1
    s = np.atleast_2d(sig)
    f = np.asarray(freq_offset, dtype=np.float64).reshape(-1, 1)
    if f.shape[0] != s.shape[0]:
        f = np.broadcast_to(f[:1] if f.shape[0] == 1 else f[:s.shape[0]], (s.shape[0], 1))
    t = np.arange(1, s.shape[1] + 1, dtype=float)
    out = (s * np.exp(-1j * (2 * np.pi * t[None, :] * f / os))).astype(s.dtype, copy=False)
    return out.ravel() if one_d else out


# ---------------------------------------------------------------------------------------------------------
# sequence location
# ---------------------------------------------------------------------------------------------------------
def _xcorr_full(x, y):
    """Full cross-correlation c[k] = sum_m x[m + k - (len(y) - 1)] conj(y[m]), k = 0 .. len(x) + len(y) - 2, by FFT."""
    nx, ny = x.shape[-1], y.shape[-1]
    n = nx + ny - 1
    nfft = _pow2_at_least(n)
    if _is_tensor(x):
        xc = x.to(torch.complex128)
        yc = y.to(torch.complex128)
        c = torch.fft.ifft(torch.fft.fft(xc, n=nfft) * torch.fft.fft(torch.flip(yc.conj(), dims=(-1,)), n=nfft))[:n]
        return c
    xc = np.asarray(x, dtype=np.complex128)
    yc = np.asarray(y, dtype=np.complex128)
    return np.fft.ifft(np.fft.fft(xc, nfft) * np.fft.fft(np.conj(yc)[::-1], nfft))[:n]


def find_sequence_offset(x, y, show_cc=False):
    """Where ``y`` sits inside ``x``: lag of the largest |cross-correlation| (negative when ``y`` starts before
    ``x``).  Returns the lag, with ``show_cc`` also the full correlation."""
    c = _xcorr_full(x, y)
    real_in = not (np.iscomplexobj(x) if not _is_tensor(x) else x.is_complex()) and \
        not (np.iscomplexobj(y) if not _is_tensor(y) else y.is_complex())
    if _is_tensor(c):
        lag = int(torch.argmax(c.abs())) - (y.shape[-1] - 1)
        c = c.real if real_in else c
    else:
        lag = int(np.argmax(np.abs(c))) - (y.shape[-1] - 1)
        c = c.real if real_in else c
    return (lag, c) if show_cc else lag


def find_sequence_offset_complex(x, y):
    """As :func:`find_sequence_offset` for a ``y`` known only up to a quarter turn.  Returns ``(lag, y turned, number
    of quarter turns k, correlation peak)`` for the k in 0..3 whose correlation has the largest real part (ties: the
    smallest k; the peak must be positive, else k = 0 with peak 0 like the reference).

    One correlation serves all four hypotheses: corr(x, y i^k) = (-i)^k corr(x, y), so the real parts in question are
    Re c, Im c, -Re c, -Im c and the lag (largest modulus) is the same for every k."""
    x_c = x.is_complex() if _is_tensor(x) else np.iscomplexobj(x)
    y_c = y.is_complex() if _is_tensor(y) else np.iscomplexobj(y)
    if not x_c and not y_c:
        lag, c = find_sequence_offset(x, y, show_cc=True)
        return lag, y, 0, c
    lag, c = find_sequence_offset(x, y, show_cc=True)
    if _is_tensor(c):
        peaks = [float(c.real.max()), float(c.imag.max()), float((-c.real).max()), float((-c.imag).max())]
    else:
        peaks = [float(c.real.max()), float(c.imag.max()), float((-c.real).max()), float((-c.imag).max())]
    k, best = 0, 0.0
    for i, p in enumerate(peaks):
        if p > best:
            k, best = i, p
    if best == 0.0:
        lag = 0
    return lag, y * 1j ** k, k, best


# ---------------------------------------------------------------------------------------------------------
# small numerics
# ---------------------------------------------------------------------------------------------------------
def moving_average(sig, N=3):
    """Mean over a sliding window of N samples along the last axis (output shorter by N - 1), as the difference of a
    running sum accumulated in the signal's dtype."""
    s = np.atleast_2d(sig)
    run = np.cumsum(s, axis=-1, dtype=s.dtype)
    head = np.zeros(s.shape[:-1] + (1,), dtype=s.dtype)
    lower = np.concatenate([head, run[..., :-N]], axis=-1)
    out = (run[..., N - 1:] - lower) / N
    return out.ravel() if np.ndim(sig) == 1 else out


def correct_shifts(shift_factors, ntaps, os):
    """Frame offsets found with an ``ntaps[0]``-tap equaliser, moved to where an ``ntaps[1]``-tap one needs them (its
    window starts half the difference earlier).  The array is adjusted in place and returned."""
    grow = ntaps[1] - ntaps[0]
    if grow % os:
        raise ValueError("search and equaliser tap counts must differ by a multiple of the oversampling (%d, %d, os %d)"
                         % (ntaps[0], ntaps[1], os))
    shift_factors = np.asarray(shift_factors)
    shift_factors -= int(grow / 2)
    return shift_factors


# ---------------------------------------------------------------------------------------------------------
# pilot-based estimates
# ---------------------------------------------------------------------------------------------------------
def _pilot_phase(received, sent):
    """Unwrapped phase of the received pilots relative to the sent ones, along the last axis."""
    return np.unwrap(np.angle(np.conj(sent) * received), axis=-1)


def pilot_based_foe(rec_symbs, pilot_symbs):
    """Frequency offset per mode = slope / 2 pi of a straight line fitted to the unwrapped pilot phase.
    Returns (mean over the modes, offsets (nmodes, 1), intercepts (nmodes, 1))."""
    phase = _pilot_phase(np.atleast_2d(rec_symbs), np.atleast_2d(pilot_symbs))
    n = np.arange(phase.shape[-1])
    # one least-squares problem per mode: a joint solve shares the same design matrix but is not guaranteed to round
    # like the reference's per-mode fits, and the parity test asks for its bits
    fits = np.array([np.polyfit(n, row, 1) for row in phase]).reshape(-1, 2)
    slope = (fits[:, 0] / (2 * np.pi)).reshape(-1, 1)
    return np.mean(slope), slope, fits[:, 1].reshape(-1, 1).copy()


def pilot_based_cpe_new(signal, pilot_symbs, pilot_idx, frame_len, seq_len=None, num_average=1, use_pilot_ratio=1,
                        max_num_blocks=None, nframes=1):
    """Carrier phase from the phase pilots of ``nframes`` frames: pilot phase -> unwrap -> centred mean over
    ``num_average`` pilots (made odd) -> linear interpolation to every symbol (held flat outside the first / last
    averaged pilot) -> de-rotation.  Returns (signal with the phase removed, phase trace), ``nframes * frame_len``
    symbols each (fewer if the signal is shorter).  The trace has the pilots' dtype, like the reference's."""
    if num_average <= 1:
        raise AssertionError("the pilot phase must be averaged over at least 3 pilots")
    if num_average % 2 == 0:
        num_average += 1
        warnings.warn("num_average must be odd: using %d" % num_average)
    signal, pilot_symbs = np.atleast_2d(signal), np.atleast_2d(pilot_symbs)
    nlen = min(frame_len * nframes, signal.shape[-1])
    used = np.asarray(pilot_idx)[:max_num_blocks:use_pilot_ratio]
    where = np.add.outer(np.arange(nframes) * frame_len, used).ravel()      # pilot positions of all frames
    where = where[where < nlen]
    sent = np.tile(pilot_symbs[:, ::use_pilot_ratio], nframes)[:, :where.size]
    got = signal[:, where]
    if got.shape != sent.shape:
        raise AssertionError("%d pilots received but %d reference pilots given" % (got.shape[-1], sent.shape[-1]))
    if sent.shape[-1] < num_average:
        raise AssertionError("averaging over %d pilots but only %d pilots in the signal" % (num_average, sent.shape[-1]))
    smooth = moving_average(_pilot_phase(got, sent), num_average)
    side = (num_average - 1) // 2
    knots = where[side:where.size - side]
    assert knots.size == smooth.shape[-1]
    trace = np.zeros((sent.shape[0], nlen), dtype=sent.dtype)
    every = np.arange(nlen)
    for m, row in enumerate(smooth):
        trace[m] = np.interp(every, knots, row)
    keep = nframes * frame_len
    return (signal[:, :nlen] * np.exp(-1j * trace))[:, :keep], trace[:, :keep]
