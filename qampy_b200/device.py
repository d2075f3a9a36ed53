"""Tensor-level (device-resident) face of the C ABI's ``*_dev`` entry points.

torch is used for what it is good at here -- owning device memory and streams; every launch below is
one of this project's own CUDA kernels, enqueued on torch's current stream without synchronising.
All tensors must live on the current CUDA device.  Batched layout: ``E`` is ``(nseg, nmodes, L)``
with unit stride along the last axis; overlapping segment views of one long capture
(:func:`segment_view`) are passed by stride, never copied.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_CODE = {torch.complex64: _lib.QB_C64, torch.complex128: _lib.QB_C128}
_REAL = {torch.complex64: torch.float32, torch.complex128: torch.float64}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.QampyB200Error("qampy_b200.device: tensors must be CUDA tensors (no CPU fallback)")


def _modes_arg(modes, nmodes):
    modes = np.arange(nmodes) if modes is None else np.atleast_1d(np.asarray(modes))
    modes = np.ascontiguousarray(modes, dtype=np.int64)
    return modes, modes.ctypes.data_as(ctypes.c_void_p)


def set_option(name, value=None):
    """Kernel-selection override for tests and tuning (``qb_set_option`` in the header): ``name`` without the ``QB_``
    prefix ("TRAIN_KERNEL", "TRAIN_LPS", "TRAIN_GLA", "LA_TILE", "BPS_KERNEL", "BPS_SPLIT"); ``None`` clears it."""
    lib = _lib.load()
    _lib.check(lib.qb_set_option(name.encode(), None if value is None else str(value).encode()))


def segment_view(E, nseg, seg_out_symbols, os, ntaps, step_symbols=None):
    """Overlapping time-segment view (no copy) of a capture ``E`` (nmodes, L): segment s covers the
    input samples that produce output symbols [s*seg_out_symbols, (s+1)*seg_out_symbols), i.e.
    ``seg_out_symbols*os + ntaps - 1`` samples starting at ``s*seg_out_symbols*os`` (SURVEY.md section 5,
    from N = (L - ntaps + 1)//os).  ``step_symbols`` (default: ``seg_out_symbols``): distance between the first output
    symbols of neighbouring segments, smaller than their length when segments overlap (pipeline: phase-search halo).
    Returns a (nseg, nmodes, L_seg) strided view."""
    nmodes, L = E.shape
    L_seg = seg_out_symbols * os + ntaps - 1
    step = (seg_out_symbols if step_symbols is None else step_symbols) * os
    if (nseg - 1) * step + L_seg > L:
        raise ValueError("capture too short for %d segments of %d symbols" % (nseg, seg_out_symbols))
    return E.as_strided((nseg, nmodes, L_seg), (step, E.stride(0), 1), E.storage_offset())


_LAYOUTS = {"throughput": 0, "latency": 1}     # qb_set_train_layout (include/qampy_b200.h)


def train_equaliser(E, TrSyms, Niter, os, mu, wx, modes, adaptive, symbols, method, err=None, layout="throughput"):
    """Train ``nseg`` independent segments.  ``wx`` (nseg, nmodes, nmodes, ntaps) and ``mu``
    (nseg, nsel) are updated in place; ``err`` (nseg, nmodes, TrSyms*Niter) is optional.
    ``layout``: "throughput" (default, batched segments) or "latency" (one stream per warp: calls on ONE
    capture, whose time is the serial depth of a stream; ``qb_set_train_layout`` in the header)."""
    if method not in _lib.METHODS:
        raise ValueError("Unknown method %s" % method)
    _check_cuda(E, wx, mu, symbols, err)
    assert E.dim() == 3 and E.stride(2) == 1 and wx.is_contiguous() and symbols.is_contiguous()
    assert mu.is_contiguous() and mu.dtype == _REAL[E.dtype] and wx.dtype == E.dtype == symbols.dtype
    nseg, nmodes, L = E.shape
    ntaps = wx.shape[-1]
    assert wx.shape == (nseg, nmodes, nmodes, ntaps)
    modes, mp = _modes_arg(modes, nmodes)
    assert mu.numel() == nseg * modes.size
    if TrSyms > 0:
        assert (TrSyms - 1) * os + ntaps <= L, "Field must be longer than the number of training symbols"
    if err is not None:
        assert err.is_contiguous() and err.shape == (nseg, nmodes, TrSyms * Niter) and err.dtype == E.dtype
    lib = _lib.load()
    old = lib.qb_set_train_layout(_LAYOUTS[layout])
    try:
        _lib.check(lib.qb_train_equaliser_dev(
            _CODE[E.dtype], _ptr(E), nseg, E.stride(0), E.stride(1), nmodes, int(TrSyms), int(Niter), int(os),
            _ptr(wx), ntaps, mp, modes.size, int(bool(adaptive)), _ptr(symbols), symbols.shape[1],
            _lib.METHODS[method], _ptr(mu), _ptr(err), _stream()))
    finally:
        lib.qb_set_train_layout(old)
    return err, wx, mu


def apply_filter_to_signal(E, os, wx, modes=None, out=None):
    _check_cuda(E, wx, out)
    assert E.dim() == 3 and E.stride(2) == 1 and wx.is_contiguous() and wx.dtype == E.dtype
    nseg, nmodes, L = E.shape
    ntaps = wx.shape[-1]
    modes, mp = _modes_arg(modes, wx.shape[1])
    N = max((L - ntaps + 1) // os, 0)
    if out is None:
        out = torch.empty((nseg, modes.size, N), dtype=E.dtype, device=E.device)
    assert out.is_contiguous() and out.shape == (nseg, modes.size, N)
    _lib.check(_lib.load().qb_apply_filter_to_signal_dev(
        _CODE[E.dtype], _ptr(E), nseg, E.stride(0), E.stride(1), nmodes, L, int(os), _ptr(wx), ntaps, mp,
        modes.size, _ptr(out), _stream()))
    return out


class BpsTables:
    """Device copies of the per-call constant tables of the blind phase search."""

    def __init__(self, Mtestangles, symbols, cdtype, device):
        from .theory import bps_test_angles
        rt = np.float32 if cdtype == np.complex64 else np.float64
        self.angles_np = bps_test_angles(Mtestangles, rt)                        # phaserecovery.py:145
        comp = np.ascontiguousarray(np.exp(1j * self.angles_np)[0], dtype=cdtype)  # pythran_dsp.py:72
        symbols = np.ascontiguousarray(np.asarray(symbols).reshape(-1), dtype=cdtype)
        lre, lim = np.zeros(64, rt), np.zeros(64, rt)
        n_re, n_im = ctypes.c_int64(0), ctypes.c_int64(0)
        code = _lib.QB_C64 if cdtype == np.complex64 else _lib.QB_C128
        grid = _lib.check(_lib.load().qb_detect_grid_host(
            code, symbols.ctypes.data_as(ctypes.c_void_p), symbols.size, lre.ctypes.data_as(ctypes.c_void_p),
            ctypes.byref(n_re), lim.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_im)))
        self.n_re, self.n_im = (n_re.value, n_im.value) if grid else (0, 0)
        self.A, self.M = comp.size, symbols.size
        self.comp = torch.from_numpy(comp).to(device)
        self.angles = torch.from_numpy(np.ascontiguousarray(self.angles_np[0])).to(device)
        self.symbols = torch.from_numpy(symbols).to(device)
        self.lev_re = torch.from_numpy(lre).to(device)
        self.lev_im = torch.from_numpy(lim).to(device)


_BPS_ACCUM = {"exact": 0, "windowed": 1}      # qb_set_bps_accumulation (include/qampy_b200.h)


def bps(E, tables, N, want_idx=True, want_ph=True, want_out=True, use_slicer=True, accum="exact"):
    """Blind phase search over every row of ``E`` (nstream, L).  Returns (Eout, ph, idx).
    ``accum``: "exact" (default: the reference's sequential running sum, bit-identical indices) or "windowed" (direct
    2N-term window sums in double: numerically sound on long signals, a flagged deviation from the reference)."""
    _check_cuda(E)
    assert E.dim() == 2 and E.stride(1) == 1
    nstream, L = E.shape
    idx = torch.empty((nstream, L), dtype=torch.int32, device=E.device) if want_idx else None
    ph = torch.empty((nstream, L), dtype=_REAL[E.dtype], device=E.device) if (want_ph or want_out) else None
    out = torch.empty((nstream, L), dtype=E.dtype, device=E.device) if want_out else None
    n_re, n_im = (tables.n_re, tables.n_im) if use_slicer else (0, 0)
    lib = _lib.load()
    old = lib.qb_set_bps_accumulation(_BPS_ACCUM[accum])
    try:
        _lib.check(lib.qb_bps_dev(
            _CODE[E.dtype], _ptr(E), nstream, E.stride(0), L, _ptr(tables.comp), _ptr(tables.angles), tables.A,
            _ptr(tables.symbols), tables.M, _ptr(tables.lev_re), n_re, _ptr(tables.lev_im), n_im, int(N),
            _ptr(idx), _ptr(ph), _ptr(out), _stream()))
    finally:
        lib.qb_set_bps_accumulation(old)
    return out, ph, idx


def freq_shift(E, freq, os, t0=0, out=None):
    """``comp_freq_offset`` (phaserecovery.py:438-473) on the device: ``out[r, t] = E[r, t] * exp(-2j pi (t0+t+1)
    freq[r] / os)`` for every row of ``E`` (nrows, L); ``freq`` one value per row (float64)."""
    _check_cuda(E, out)
    assert E.dim() == 2 and E.stride(1) == 1
    nrows, L = E.shape
    f = torch.as_tensor(np.ascontiguousarray(np.broadcast_to(np.asarray(freq, dtype=np.float64).reshape(-1), (nrows,))),
                        device=E.device)
    if out is None:
        out = torch.empty((nrows, L), dtype=E.dtype, device=E.device)
    assert out.shape == E.shape and out.stride(1) == 1 and out.dtype == E.dtype
    _lib.check(_lib.load().qb_freq_shift_dev(_CODE[E.dtype], _ptr(E), nrows, E.stride(0), L, _ptr(f), int(os), int(t0),
                                             _ptr(out), out.stride(0), _stream()))
    return out


def pilot_cpe(E, pilot_idx, pilots, num_average, want_trace=False):
    """``pilot_based_cpe_new`` (pilotbased_receiver.py:258-327) for one frame per row of ``E`` (nrows, nlen):
    ``pilot_idx`` (npilots) sorted positions, ``pilots`` (nrows, npilots) reference pilots.  Returns
    (compensated rows, phase trace or None)."""
    _check_cuda(E, pilots)
    assert E.dim() == 2 and E.stride(1) == 1 and pilots.dim() == 2 and pilots.stride(1) == 1
    nrows, nlen = E.shape
    assert pilots.shape[0] == nrows and pilots.dtype == E.dtype
    if not (num_average % 2):
        num_average += 1
    idx = torch.as_tensor(np.ascontiguousarray(pilot_idx, dtype=np.int64), device=E.device)
    assert idx.numel() == pilots.shape[1] and int(idx[-1]) < nlen
    out = torch.empty((nrows, nlen), dtype=E.dtype, device=E.device)
    trace = torch.empty((nrows, nlen), dtype=_REAL[E.dtype], device=E.device) if want_trace else None
    _lib.check(_lib.load().qb_pilot_cpe_dev(_CODE[E.dtype], _ptr(E), nrows, E.stride(0), nlen, _ptr(idx), _ptr(pilots),
                                            pilots.stride(0), idx.numel(), int(num_average), _ptr(out), out.stride(0),
                                            _ptr(trace), trace.stride(0) if want_trace else 0, _stream()))
    return out, trace


def viterbiviterbi(E, N, M):
    """``viterbiviterbi`` (phaserecovery.py:40-79) for every row of ``E`` (nrows, L): returns (compensated rows
    with zeros where no full window exists, phase estimates (nrows, L - N + 1))."""
    _check_cuda(E)
    assert E.dim() == 2 and E.stride(1) == 1
    nrows, L = E.shape
    out = torch.empty((nrows, L), dtype=E.dtype, device=E.device)
    ph = torch.empty((nrows, max(L - int(N) + 1, 0)), dtype=_REAL[E.dtype], device=E.device)
    _lib.check(_lib.load().qb_viterbiviterbi_dev(_CODE[E.dtype], _ptr(E), nrows, E.stride(0), L, int(N), int(M),
                                                 _ptr(out), out.stride(0), _ptr(ph), ph.stride(0), _stream()))
    return out, ph
