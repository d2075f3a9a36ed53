"""L2 drop-in for ``qampy/core/equalisation/equalisation.py``: ``equalise_signal`` (:468-566),
``dual_mode_equalisation`` (:400-466) and ``apply_filter`` (:138-188) with the same signatures,
return tuples and error behaviour, but with the signal resident in HBM across
train -> train -> apply (one H2D copy of ``E``, no intermediate round trips).

Host work that stays in NumPy is exactly the reference's own glue: per-method constant tables,
default training length, tap initialisation (see ``qampy_b200/theory.py``).
"""
import numpy as np
import torch

from . import _lib, device, theory
from .theory import (DATA_AIDED, DECISION_BASED, NONDECISION_BASED, REAL_VALUED,  # noqa: F401 (API parity)
                     TRAINING_FCTS)


def _dev():
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _check_complex(E):
    E = np.asarray(E)
    if E.dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
        raise TypeError("qampy_b200 equalises complex64/complex128 signals, got %s" % E.dtype)
    return E


def _to_dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)


def apply_filter(E, os, wxy, method="pyt", modes=None):
    """Apply the equaliser taps to the signal (decimating by ``os``)."""
    if method not in ("pyt", "py", "cuda"):
        raise NotImplementedError("Only py and pythran methods are implemented")
    wxy = np.asarray(wxy)
    if not np.iscomplexobj(wxy) or not np.iscomplexobj(E):
        # real-valued taps (equalisation.py:177-186): filter the real/imaginary rows, then recombine
        from . import pythran_equalisation as pe
        E = np.asarray(E)
        if np.iscomplexobj(E):
            E = theory.convert_sig_to_real(np.atleast_2d(_check_complex(E)))
        modes = np.arange(wxy.shape[0]) if modes is None else np.copy(np.atleast_1d(modes))
        Etmp = pe.apply_filter_to_signal(np.copy(E), int(os), np.copy(wxy), modes)
        if E.itemsize == 8:
            return theory.convert_sig_to_cmplx(Etmp, modes.shape[0], np.complex128(1j))
        if E.itemsize == 4:
            return theory.convert_sig_to_cmplx(Etmp, modes.shape[0], np.complex64(1j))
        raise ValueError("The field has an unknown data type")
    E = _check_complex(E)
    dev = _dev()
    modes = np.arange(wxy.shape[0]) if modes is None else np.copy(np.atleast_1d(modes))
    assert np.max(modes) < wxy.shape[0], "largest mode number is larger than shape of signal"
    Ed = _to_dev(np.atleast_2d(E), dev)[None]
    wd = _to_dev(wxy.astype(E.dtype), dev)[None]
    out = device.apply_filter_to_signal(Ed, int(os), wd, modes)
    return out[0].cpu().numpy()


def _train_stage(Ed, os, mu, M, wd, ntaps, TrSyms, Niter, method, adaptive, symbols, modes, cdtype,
                 mu_shared=True):
    """One equalise_signal training stage on device tensors (nseg = 1: the reference's own call on one capture,
    one serial stream per trained mode -> the latency layout of the training kernels).  Returns err (numpy)."""
    dev = Ed.device
    nmodes, L = Ed.shape[1], Ed.shape[2]
    rt = np.float32 if cdtype == np.complex64 else np.float64
    if TrSyms is None:
        TrSyms = theory.cal_training_symbol_len(os, ntaps, L)
    symbols = theory.unique_alphabet(theory.reshape_symbols(symbols, method, M, cdtype, nmodes), method)
    sd = _to_dev(symbols, dev)
    err = torch.zeros((1, nmodes, TrSyms * Niter), dtype=Ed.dtype, device=dev)
    mu = rt(mu)
    if adaptive and mu_shared and len(modes) > 1:
        # the reference carries ONE step size through the modes in list order when interpreted
        mud = torch.full((1, 1), float(mu), dtype=device._REAL[Ed.dtype], device=dev)
        for m in modes:
            device.train_equaliser(Ed, TrSyms, Niter, os, mud, wd, [m], adaptive, sd, method, err, layout="latency")
    else:
        mud = torch.full((1, len(modes)), float(mu), dtype=device._REAL[Ed.dtype], device=dev)
        device.train_equaliser(Ed, TrSyms, Niter, os, mud, wd, modes, adaptive, sd, method, err, layout="latency")
    return err


def _prepare(E, wxy, Ntaps, modes, method):
    method = method.lower()
    if method in REAL_VALUED:
        raise NotImplementedError("real-valued equaliser methods (%s) run through equalise_signal only" % method)
    if method not in _lib.METHODS:
        raise ValueError("Unknown method %s" % method)
    E = np.atleast_2d(_check_complex(E))
    nmodes = E.shape[0]
    if modes is None:
        modes = np.arange(nmodes)
    else:
        modes = np.atleast_1d(modes)
        assert np.max(modes) < nmodes, "largest mode number is larger than shape of signal"
    wxy_user = None
    if wxy is None:
        wxy = theory.init_taps(Ntaps, nmodes, E.dtype)
    else:
        wxy_user = wxy
        wxy = np.ascontiguousarray(wxy, dtype=E.dtype)
        Ntaps = wxy.shape[-1]
        assert wxy.ndim == 3, "wxy needs to be three dimensional"
        assert wxy.shape[:2] == (nmodes, nmodes), "The first 2 dimensions of wxy need to be the same shape as E"
    return method, E, nmodes, modes, wxy, wxy_user, Ntaps


def equalise_signal(E, os, mu, M, wxy=None, Ntaps=None, TrSyms=None, Niter=1, method="mcma",
                    adaptive_stepsize=False, symbols=None, modes=None, apply=False, **kwargs):
    """Blind equalisation with one training method; see the reference docstring for the arguments.
    Returns ``(E_out,) wxy, err``.  A user-supplied ``wxy`` that is already a C-contiguous array of
    the signal's dtype is trained in place, as in the reference (:547)."""
    if method.lower() in REAL_VALUED:
        return _equalise_signal_real(E, os, mu, M, wxy, Ntaps, TrSyms, Niter, method.lower(), adaptive_stepsize,
                                     symbols, modes, apply, **kwargs)
    method, E, nmodes, modes, wxy, wxy_user, Ntaps = _prepare(E, wxy, Ntaps, modes, method)
    dev = _dev()
    Ed = _to_dev(E, dev)[None]
    wd = _to_dev(wxy, dev)[None].contiguous()
    err = _train_stage(Ed, int(os), mu, M, wd, Ntaps, TrSyms, int(Niter), method, adaptive_stepsize, symbols,
                       modes, E.dtype, kwargs.get("mu_shared", True))
    out = device.apply_filter_to_signal(Ed, int(os), wd, modes) if apply else None
    np.copyto(wxy, wd[0].cpu().numpy())
    err = err[0].cpu().numpy()
    if apply:
        return out[0].cpu().numpy(), wxy, err
    return wxy, err


SUPPORTS_RETURN_ERR = True      # equalise_windows(return_err=False): callers that only want the taps (pilots.py)


def equalise_windows(E, starts, window, os, mu, M, wxy=None, Ntaps=None, TrSyms=None, Niter=1, method="mcma",
                     adaptive_stepsize=False, symbols=None, modes=None, apply=False, return_err=True, **kwargs):
    """``equalise_signal(E[:, s:s + window], os, mu, M, wxy=..., ...)`` for every ``s`` in ``starts`` as ONE batched
    launch per stage: the windows are strided views of the capture on the device, one segment per window.

    This is how the pilot-based receiver maps onto the GPU: the frame search trains
    ~130 candidate windows of a capture (``pilotbased_receiver.py:395-400``) and the pilot equaliser trains
    every frame's pilot sequence from the same initial taps (``qampy/equalisation.py:384-389``) -- short,
    independent, serial trainings that only fill the machine side by side.

    ``E``: NumPy array or CUDA tensor (nmodes, L).  ``wxy``: None (centre spike), (nmodes, nmodes, Ntaps)
    shared by all windows, or (nwin, nmodes, nmodes, Ntaps).  ``symbols``: as for ``equalise_signal``, shared
    by all windows.  Returns ``(wxy (nwin, nmodes, nmodes, Ntaps), err (nwin, nmodes, TrSyms*Niter))`` and, with
    ``apply``, the equalised windows ``(nwin, len(modes), (window - Ntaps + 1)//os)`` first -- NumPy arrays.
    ``return_err=False``: the per-symbol error is neither stored nor copied back (``err`` is None): the pilot
    equaliser trains Niter = 30 passes per stage and only ever looks at the taps."""
    starts = np.atleast_1d(np.asarray(starts, dtype=np.int64))
    method_l = method.lower()
    if method_l in REAL_VALUED or (starts.size > 1 and np.unique(np.diff(starts)).size > 1):
        outs, taps, errs = [], [], []
        Eh = E.cpu().numpy() if torch.is_tensor(E) else np.asarray(E)
        for k, s0 in enumerate(starts):         # irregular windows / real-valued methods: plain loop
            w0 = None if wxy is None else (wxy[k] if np.ndim(wxy) == 4 else wxy)
            ret = equalise_signal(Eh[:, s0:s0 + window], os, mu, M, wxy=None if w0 is None else np.array(w0),
                                  Ntaps=Ntaps, TrSyms=TrSyms, Niter=Niter, method=method,
                                  adaptive_stepsize=adaptive_stepsize, symbols=symbols, modes=modes, apply=apply,
                                  **kwargs)
            if apply:
                outs.append(ret[0])
            taps.append(ret[-2])
            errs.append(ret[-1])
        res = (np.asarray(taps), np.asarray(errs))
        return (np.asarray(outs),) + res if apply else res
    dev = _dev()
    if torch.is_tensor(E):
        Ed = E if E.is_cuda else E.to(dev)
        cdtype = np.dtype(np.complex64 if Ed.dtype == torch.complex64 else np.complex128)
    else:
        Eh = np.atleast_2d(_check_complex(E))
        cdtype = Eh.dtype
        Ed = _to_dev(Eh, dev)
    if method_l not in _lib.METHODS:
        raise ValueError("Unknown method %s" % method)
    nmodes, L = Ed.shape
    modes = np.arange(nmodes) if modes is None else np.atleast_1d(modes)
    assert np.max(modes) < nmodes, "largest mode number is larger than shape of signal"
    nwin, step = starts.size, (int(starts[1] - starts[0]) if starts.size > 1 else 0)
    assert starts[0] >= 0 and starts[-1] + window <= L, "window beyond the end of the signal"
    if wxy is None:
        w0 = theory.init_taps(Ntaps, nmodes, cdtype)
    else:
        w0 = np.ascontiguousarray(wxy, dtype=cdtype)
        Ntaps = w0.shape[-1]
        assert w0.shape[-3:] == (nmodes, nmodes, Ntaps), "wxy must be (nmodes, nmodes, Ntaps) per window"
    wd = _to_dev(w0, dev)
    wd = (wd[None].repeat(nwin, 1, 1, 1) if wd.dim() == 3 else wd).contiguous()
    assert wd.shape[0] == nwin, "one set of initial taps per window expected"
    Ev = Ed.as_strided((nwin, nmodes, int(window)), (step, Ed.stride(0), 1), Ed.storage_offset() + int(starts[0]))
    rt = np.float32 if cdtype == np.complex64 else np.float64
    if TrSyms is None:
        TrSyms = theory.cal_training_symbol_len(int(os), Ntaps, int(window))
    sd = _to_dev(theory.unique_alphabet(theory.reshape_symbols(symbols, method_l, M, cdtype.type, nmodes), method_l), dev)
    err = torch.zeros((nwin, nmodes, TrSyms * int(Niter)), dtype=Ed.dtype, device=dev) if return_err else None
    mu = float(rt(mu))
    if adaptive_stepsize and kwargs.get("mu_shared", True) and len(modes) > 1:
        mud = torch.full((nwin, 1), mu, dtype=device._REAL[Ed.dtype], device=dev)   # one step size per window,
        for m in modes:                                                              # carried through the modes
            device.train_equaliser(Ev, TrSyms, int(Niter), int(os), mud, wd, [m], True, sd, method_l, err)
    else:
        mud = torch.full((nwin, len(modes)), mu, dtype=device._REAL[Ed.dtype], device=dev)
        device.train_equaliser(Ev, TrSyms, int(Niter), int(os), mud, wd, modes, adaptive_stepsize, sd, method_l, err)
    res = (wd.cpu().numpy(), err.cpu().numpy() if return_err else None)
    if apply:
        out = device.apply_filter_to_signal(Ev, int(os), wd, modes)
        return (out.cpu().numpy(),) + res
    return res


def apply_windows(E, starts, window, os, wxy, modes=None, as_tensor=False):
    """``apply_filter(E[:, s:s + window], os, wxy[k])`` for every window k as one launch (per-window taps
    ``wxy`` (nwin, nmodes, nmodes, Ntaps), or one shared set).  Returns (nwin, len(modes), (window-Ntaps+1)//os),
    a NumPy array or (``as_tensor``) a CUDA tensor."""
    starts = np.atleast_1d(np.asarray(starts, dtype=np.int64))
    dev = _dev()
    if torch.is_tensor(E):
        Ed = E if E.is_cuda else E.to(dev)
        cdtype = np.dtype(np.complex64 if Ed.dtype == torch.complex64 else np.complex128)
    else:
        Eh = np.atleast_2d(_check_complex(E))
        cdtype = Eh.dtype
        Ed = _to_dev(Eh, dev)
    nmodes, L = Ed.shape
    w = np.ascontiguousarray(wxy, dtype=cdtype)
    nwin = starts.size
    wd = _to_dev(w, dev)
    wd = (wd[None].repeat(nwin, 1, 1, 1) if wd.dim() == 3 else wd).contiguous()
    modes = np.arange(w.shape[-3]) if modes is None else np.atleast_1d(modes)
    assert starts[0] >= 0 and starts.max() + window <= L, "window beyond the end of the signal"
    if nwin > 1 and np.unique(np.diff(starts)).size == 1 and starts[1] > starts[0]:
        Ev = Ed.as_strided((nwin, nmodes, int(window)), (int(starts[1] - starts[0]), Ed.stride(0), 1),
                           Ed.storage_offset() + int(starts[0]))
        out = device.apply_filter_to_signal(Ev, int(os), wd, modes)
        return out if as_tensor else out.cpu().numpy()
    outs = [device.apply_filter_to_signal(Ed[None, :, int(s0):int(s0) + int(window)], int(os), wd[k:k + 1], modes)[0]
            for k, s0 in enumerate(starts)]
    out = torch.stack(outs)
    return out if as_tensor else out.cpu().numpy()


def _equalise_signal_real(E, os, mu, M, wxy, Ntaps, TrSyms, Niter, method, adaptive_stepsize, symbols, modes,
                          apply, **kwargs):
    """equalise_signal for the real-valued methods (equalisation.py:529-565): the signal becomes 2*nmodes
    real rows, taps and error are real, the equalised signal is recombined to complex."""
    from . import pythran_equalisation as pe
    E = theory.convert_sig_to_real(np.atleast_2d(_check_complex(E)))
    mu = E.dtype.type(mu)
    nmodes = E.shape[0]
    if modes is None:
        modes = np.arange(nmodes)
    else:
        modes = np.atleast_1d(modes)
        modes = np.hstack([modes, modes + nmodes // 2])
        assert np.max(modes) < nmodes, "largest mode number is larger than shape of signal"
    if wxy is None:
        wxy = theory.init_taps(Ntaps, nmodes, E.dtype)
    else:
        wxy = np.ascontiguousarray(wxy, dtype=E.dtype)
        Ntaps = wxy.shape[-1]
        assert wxy.ndim == 3, "wxy needs to be three dimensional"
        assert wxy.shape[:2] == (nmodes, nmodes), "The first 2 dimensions of wxy need to be the same shape as E"
    if TrSyms is None:
        TrSyms = theory.cal_training_symbol_len(os, Ntaps, E.shape[-1])
    symbols = theory.reshape_symbols(symbols, method, M, E.dtype, nmodes)
    err, wxy, mu = pe.train_equaliser_realvalued(E, TrSyms, int(Niter), int(os), mu, wxy, modes, adaptive_stepsize,
                                                 symbols.copy(), method[:-5], kwargs.get("mu_shared", True))
    if apply:
        return apply_filter(E, os, wxy, modes=modes), wxy, err
    return wxy, err


def dual_mode_equalisation(E, os, mu, M, wxy=None, Ntaps=None, TrSyms=(None, None), Niter=(1, 1),
                           methods=("mcma", "sbd"), adaptive_stepsize=(False, False), symbols=None,
                           modes=None, apply=True, **kwargs):
    """Two-stage blind equalisation: stage 2 restarts at sample 0 with the stage-1 taps and the
    returned signal is the final taps applied to the whole input (:460-464)."""
    symbols = np.atleast_1d(symbols)
    if symbols.ndim < 3:
        symbols = np.tile(symbols, (2, 1, 1))
    sy = [None if symbols[i].dtype == object else symbols[i] for i in range(2)]
    m0, E, nmodes, modes, wxy, wxy_user, Ntaps = _prepare(E, wxy, Ntaps, modes, methods[0])
    m1 = _prepare(E, wxy, None, modes, methods[1])[0]
    dev = _dev()
    Ed = _to_dev(E, dev)[None]
    wd = _to_dev(wxy, dev)[None].contiguous()
    shared = kwargs.get("mu_shared", True)
    err1 = _train_stage(Ed, int(os), mu[0], M, wd, Ntaps, TrSyms[0], int(Niter[0]), m0, adaptive_stepsize[0],
                        sy[0], modes, E.dtype, shared)
    err2 = _train_stage(Ed, int(os), mu[1], M, wd, Ntaps, TrSyms[1], int(Niter[1]), m1, adaptive_stepsize[1],
                        sy[1], modes, E.dtype, shared)
    out = device.apply_filter_to_signal(Ed, int(os), wd, modes) if apply else None
    np.copyto(wxy, wd[0].cpu().numpy())
    errs = (err1[0].cpu().numpy(), err2[0].cpu().numpy())
    if apply:
        return out[0].cpu().numpy(), wxy, errs
    return wxy, errs
