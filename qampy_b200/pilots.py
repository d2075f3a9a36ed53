"""Pilot-based receiver on arrays (BASELINE config C4): frame synchronisation,
pilot-sequence equaliser, pilot-based carrier phase and frequency offset estimation.

Array-level counterparts of ``qampy/core/pilotbased_receiver.py`` (``frame_sync`` :329-434,
``equalize_pilot_sequence`` :454-554, ``pilot_based_cpe_new`` :258-327, ``pilot_based_foe`` :32-73) and of
the helpers they use (``phaserecovery.find_freq_offset`` / ``comp_freq_offset`` :385-473,
``ber_functions.find_sequence_offset[_complex]`` :33-106, ``filter.moving_average`` :215-237), with the same
argument meaning and returns, plus array forms of the signal-object methods built on them
(``sync2frame`` / ``corr_foe`` ``signals.py:1709-1747``, ``pilot_equaliser[_nframes]``
``qampy/equalisation.py:266-397``).

What runs where: every adaptive-filter training and every FIR goes through the CUDA equaliser
(``qampy_b200.equalisation``); the frame search trains ALL candidate windows of a capture in one batched
launch (one segment per window, the windows being strided views of the capture), and
``pilot_equaliser_nframes`` trains the frames that share the initial taps side by side.  What is left on
the host is O(pilots) NumPy: correlations of a 4k-sample sequence, phase unwrap / moving average /
interpolation of the pilot phases.

``backend`` (tests only): a module-like object with ``equalise_signal``, ``apply_filter`` and optionally
``equalise_windows``; the default is this package's CUDA path.
"""
import warnings

import numpy as np

from . import theory

FRAME_SYNC_THRS = 120   # autocorrelation peak below which the sync is reported as failed (pilotbased_receiver.py:368)


def _backend(backend):
    if backend is not None:
        return backend
    from . import equalisation
    return equalisation


# ---------------------------------------------------------------------------------------------------------
# estimators (spectral frequency offset, sequence location, pilot phase / frequency): pilot_estimators.py
# ---------------------------------------------------------------------------------------------------------
from .pilot_estimators import (comp_freq_offset, correct_shifts, find_freq_offset, find_sequence_offset,  # noqa: E402,F401
                               find_sequence_offset_complex, moving_average, pilot_based_cpe_new, pilot_based_foe)


# ---------------------------------------------------------------------------------------------------------
# frame synchronisation
# ---------------------------------------------------------------------------------------------------------
def _train_windows(be, rx, starts, window, os, mu, M_pilot, Ntaps, eqargs):
    """Taps and error of an equaliser trained from centre-spike taps on every window rx[:, s : s + window]."""
    if hasattr(be, "equalise_windows"):
        return be.equalise_windows(rx, starts, window, os, mu, M_pilot, Ntaps=Ntaps, **eqargs)
    taps, errs = [], []
    for s in starts:
        w, e = be.equalise_signal(rx[:, s:s + window], os, mu, M_pilot, Ntaps=Ntaps, **eqargs)
        taps.append(w)
        errs.append(e)
    return np.asarray(taps), np.asarray(errs)


def frame_sync(rx_signal, ref_symbs, os, frame_len=2 ** 16, M_pilot=4, mu=1e-3, Ntaps=17, backend=None,
               rx_device=None, **eqargs):
    """Locate the pilot sequence inside a frame (pilotbased_receiver.py:329-434).

    A blind equaliser (``eqargs``: method, Niter, adaptive_stepsize ...) is trained on windows of one
    pilot-sequence length, half a window apart, over one frame; the window with the lowest error variance
    per mode holds pilots (constant-modulus QPSK) rather than payload.  Its taps equalise a three-window
    stretch, a coarse frequency offset is removed and the pilot sequence is located by cross-correlation,
    which also resolves which transmitted mode each received mode carries.

    Returns (shift_factor per mode, coarse frequency offset, mode_sync_order, taps of the last mode's
    window, sync_ok)."""
    be = _backend(backend)
    ref_symbs = np.atleast_2d(ref_symbs)
    on_device = rx_device is not None
    if on_device:
        # the capture is resident on the GPU: the stretch around the quietest window is filtered, de-rotated and
        # correlated there (own FIR / frequency-shift kernels, cuFFT); a handful of scalars come back
        import torch
        from . import device
        rx_signal = rx_device
        ref_dev = torch.from_numpy(np.ascontiguousarray(ref_symbs)).to(rx_device.device)
        is_complex = rx_device.is_complex()
    else:
        rx_signal = np.atleast_2d(rx_signal)
        is_complex = np.iscomplexobj(rx_signal)
    nmodes, seq_len = rx_signal.shape[0], ref_symbs.shape[-1]
    if rx_signal.shape[-1] < (frame_len + 2 * seq_len) * os:
        raise AssertionError("the capture must hold one frame plus two pilot-sequence lengths")
    method = eqargs.get("method")
    if method in theory.REAL_VALUED and is_complex:
        raise ValueError("frame search with the real-valued equaliser method %s is not supported" % method)
    if method in theory.DATA_AIDED:
        raise ValueError("frame search with the data-aided equaliser method %s is not supported" % method)
    window, hop, starts = _search_windows(seq_len, frame_len, os)
    taps, errs = _train_windows(be, rx_signal, starts, window, os, mu, M_pilot, Ntaps, eqargs)
    # a window full of pilots (constant-modulus QPSK) leaves the blind equaliser with the smallest error variance
    quietest = np.argmin(np.var(errs, axis=-1), axis=0)            # (nmodes,): window index per received mode
    shift_factor = np.zeros(nmodes, dtype=int)
    mode_sync_order = np.zeros(nmodes, dtype=int)
    free = np.ones(ref_symbs.shape[0], dtype=bool)                 # transmitted modes not yet claimed
    sync_ok, foe_coarse, wx1 = True, None, None
    for mode in range(nmodes):
        start, wx1 = int(starts[quietest[mode]]), taps[quietest[mode]]
        # three half-windows around the quietest one, equalised with its taps, coarse frequency offset removed
        stretch = rx_signal[:, start - window:start + window]
        if on_device:
            wd = torch.from_numpy(np.ascontiguousarray(wx1)).to(rx_device.device)
            symbs = device.apply_filter_to_signal(stretch.unsqueeze(0), os, wd.unsqueeze(0))[0]
            foe_coarse = find_freq_offset(symbs)
            row = device.freq_shift(symbs, foe_coarse[:, 0], 1)[mode]
            refs = ref_dev
        else:
            symbs = be.apply_filter(stretch, os, wx1)
            foe_coarse = find_freq_offset(symbs)
            row = comp_freq_offset(symbs, foe_coarse)[mode]
            refs = ref_symbs
        peak = np.zeros(ref_symbs.shape[0])
        lag = np.zeros(ref_symbs.shape[0], dtype=np.int64)
        for cand in np.nonzero(free)[0]:
            lag[cand], _, _, peak[cand] = find_sequence_offset_complex(refs[cand], row)
        sent = int(np.argmax(peak))
        if peak[sent] < FRAME_SYNC_THRS:
            warnings.warn("frame search: correlation peak %.0f below %d, the pilot sequence was probably not found"
                          % (peak[sent], FRAME_SYNC_THRS))
            sync_ok = False
        free[sent] = False
        mode_sync_order[mode] = sent
        shift_factor[mode] = start - window - os * lag[sent]
    return shift_factor, foe_coarse, mode_sync_order, wx1, sync_ok


def _search_windows(seq_len, frame_len, os):
    """Candidate windows of the frame search: each one pilot-sequence long (``window`` samples), half a window apart
    (``hop``), from one window into the capture up to one hop past a frame -- so that every window has a full window
    of signal on either side.  Returns (window, hop, start sample of every candidate)."""
    window = seq_len * os
    hop = window // 2
    return window, hop, hop * np.arange(2, (frame_len * os) // hop + 1)


def sync2frame(rx_signal, pilot_seq, os, frame_len, M_pilot=4, backend=None, **kwargs):
    """Array form of ``SignalWithPilots.sync2frame`` (signals.py:1709-1741): frame search with the
    reference's defaults, then the modes re-ordered to the pilots' order.
    Returns (aligned signal, shift factors, coarse frequency offset, sync taps, sync_ok)."""
    eqargs = {"adaptive_stepsize": True, "Niter": 10, "method": "cma", "Ntaps": 17, "mu": 5e-3}
    eqargs.update(kwargs)
    mu, Ntaps = eqargs.pop("mu"), eqargs.pop("Ntaps")
    shift, foe, order, wx1, ok = frame_sync(rx_signal, pilot_seq, os, frame_len=frame_len, M_pilot=M_pilot, mu=mu,
                                            Ntaps=Ntaps, backend=backend, **eqargs)
    aligned = np.atleast_2d(rx_signal)[order, :]
    shift[shift < 0] += frame_len * os
    return aligned, shift[order], foe, wx1, ok


def corr_foe(rx_signal, foe, os, additional_foe=0):
    """Array form of ``SignalWithPilots.corr_foe`` (signals.py:1744-1747)."""
    foe = np.asarray(foe)
    return comp_freq_offset(rx_signal, np.ones(foe.shape) * (np.mean(foe) + additional_foe), os)


# ---------------------------------------------------------------------------------------------------------
# pilot-sequence equaliser
# ---------------------------------------------------------------------------------------------------------
def equalize_pilot_sequence(rx_signal, ref_symbs, shift_fctrs, os, foe_comp=False, mu=(1e-4, 1e-4), M_pilot=4,
                            Ntaps=45, Niter=30, adaptive_stepsize=True, methods=('cma', 'cma'), wxinit=None,
                            backend=None):
    """Train the equaliser on the pilot sequence in two steps (pilotbased_receiver.py:454-554): ``methods[0]``
    from ``wxinit`` (optionally followed by a pilot-based frequency offset estimate), then ``methods[0]`` and
    ``methods[1]`` again with the pilot sequence as training symbols.  Returns (taps, frequency offsets)."""
    be = _backend(backend)
    rx_signal, ref_symbs = np.atleast_2d(rx_signal), np.atleast_2d(ref_symbs)
    npols = rx_signal.shape[0]
    seq_len = ref_symbs.shape[-1]
    if (methods[0] in theory.REAL_VALUED) != (methods[1] in theory.REAL_VALUED):
        raise ValueError("Using a complex and real-valued equalisation method is not supported")
    span = seq_len * os + Ntaps - 1
    per_mode = np.unique(shift_fctrs).shape[0] > 1
    groups = [(shift_fctrs[i], [i]) for i in range(npols)] if per_mode else [(shift_fctrs[0], None)]
    # step 1: blind pre-convergence (equalised pilots are only needed for the frequency offset estimate)
    wx = wxinit
    syms_out = np.zeros_like(ref_symbs)
    for start, modes in groups:
        seg = rx_signal[:, start:start + span]
        out, wx, _ = be.equalise_signal(seg, os, mu[0], M_pilot, wxy=wx if per_mode else wxinit, Ntaps=Ntaps,
                                        Niter=Niter, method=methods[0], adaptive_stepsize=adaptive_stepsize,
                                        apply=True, **({"modes": modes} if modes is not None else {}))
        if modes is None:
            syms_out = out
        else:
            syms_out[modes[0]] = out
    if foe_comp:
        foe, foe_per_mode, _ = pilot_based_foe(syms_out, ref_symbs)
        foe_all = np.ones(foe_per_mode.shape) * foe
    else:
        foe_all = np.zeros([npols, 1])
    # step 2: both methods again, now with the pilots as training symbols
    taps = wx.copy()
    for start, modes in groups:
        seg = rx_signal[:, start:start + span]
        if foe_comp:
            seg = comp_freq_offset(seg, foe_all, os=os)
        kw = {"modes": modes} if modes is not None else {}
        taps, _ = be.equalise_signal(seg, os, mu[0], M_pilot, wxy=taps, Ntaps=Ntaps, Niter=Niter, method=methods[0],
                                     adaptive_stepsize=adaptive_stepsize, symbols=ref_symbs, apply=False, **kw)
        # the per-mode branch of the reference passes M = 4 here (:542), the common branch M_pilot (:551)
        taps, _ = be.equalise_signal(seg, os, mu[1], 4 if per_mode else M_pilot, wxy=taps, Ntaps=Ntaps, Niter=Niter,
                                     method=methods[1], adaptive_stepsize=adaptive_stepsize, symbols=ref_symbs,
                                     apply=False, **kw)
    return np.array(taps), foe_all


def apply_to_frames(rx_signal, wxy, shiftfctrs, os, frame_len, frames=(0,), synctaps=None, backend=None):
    """Equalise whole frames with trained taps: array form of ``_apply_to_pilotsignal``
    (qampy/equalisation.py:42-87).  ``shiftfctrs`` are the frame offsets found with ``synctaps`` taps."""
    be = _backend(backend)
    rx_signal = np.atleast_2d(rx_signal)
    frames = list(frames)
    Ntaps = wxy.shape[-1]
    shifts = np.array(shiftfctrs, dtype=int)
    if synctaps is not None and Ntaps != synctaps:
        shifts = shifts - (Ntaps - synctaps) // 2
    if np.min(shifts) < 0:
        shifts += os * frame_len
    assert shifts.max() + os * frame_len * (max(frames) + 1) < rx_signal.shape[-1] - (Ntaps - 1), \
        "Trying to equalise frame {}, but signal is not long enough".format(max(frames))
    per_mode = np.unique(shifts).shape[0] > 1
    mode_groups = np.arange(wxy.shape[0]).reshape(-1, rx_signal.shape[0]).T   # real-valued taps: (re, im) rows of a mode
    if np.all(np.diff(frames) == 1):
        runs = [(frames[0], frames[-1] - frames[0] + 1)]      # consecutive frames: one FIR over the whole run
    else:
        runs = [(f, 1) for f in frames]
    pieces = []
    for f0, n in runs:
        span = n * frame_len * os + Ntaps - 1
        if per_mode:
            rows = []
            for mode in mode_groups:
                i0 = shifts[mode[0]] + f0 * os * frame_len
                rows.append(be.apply_filter(rx_signal[:, i0:i0 + span], os, wxy, modes=mode))
            pieces.append(np.squeeze(np.array(rows)))
        else:
            i0 = shifts[0] + f0 * os * frame_len
            pieces.append(be.apply_filter(rx_signal[:, i0:i0 + span], os, wxy))
    return np.hstack(pieces) if len(pieces) > 1 else pieces[0]


def pilot_equaliser(rx_signal, pilot_seq, shiftfctrs, os, frame_len, mu, Ntaps, synctaps=17, apply=True,
                    foe_comp=True, wxinit=None, frame=0, verbose=False, backend=None, **eqkwargs):
    """Pilot-based equalisation of one frame: array form of ``qampy.equalisation.pilot_equaliser`` (:266-334).
    ``rx_signal`` must be frame-synchronised (``sync2frame``).  Returns taps (and the equalised frame if
    ``apply``; with ``verbose`` also the frequency offsets and ``(Ntaps, synctaps)``)."""
    if shiftfctrs is None:
        raise ValueError("The signal has to be synchronised to the frame first")
    shifts = np.array(shiftfctrs, dtype=int)
    mu = np.atleast_1d(mu)
    if len(mu) == 1:
        mu = np.repeat(mu, 2)
    if wxinit is not None:
        Ntaps = wxinit.shape[-1]
    if (abs(Ntaps - synctaps) % 2) != 0:
        raise ValueError("Tap difference need to be an integer of the oversampling")
    elif Ntaps != synctaps:
        shifts = shifts - (Ntaps - synctaps) // 2 + os * frame_len * frame
    rx_signal = np.atleast_2d(rx_signal)
    assert rx_signal.shape[-1] - shifts.max() > frame_len * os, \
        "You are trying to equalise an incomplete frame which does not work"
    taps, foe_all = equalize_pilot_sequence(rx_signal, pilot_seq, shifts, os=os, mu=mu, foe_comp=foe_comp,
                                            Ntaps=Ntaps, wxinit=wxinit, backend=backend, **eqkwargs)
    sig = comp_freq_offset(rx_signal, foe_all, os) if foe_comp else rx_signal
    if not apply:
        return (taps, foe_all, (Ntaps, synctaps)) if verbose else taps
    eq = apply_to_frames(sig, taps, shiftfctrs, os, frame_len, frames=[frame], synctaps=synctaps, backend=backend)
    return (taps, eq, foe_all, (Ntaps, synctaps)) if verbose else (taps, eq)


def _pilot_frames_batched(be, rx_signal, pilot_seq, shiftfctrs, os, frame_len, mu, Ntaps, synctaps, frames, wxinit,
                          apply, M_pilot=4, Niter=30, adaptive_stepsize=True, methods=('cma', 'cma'), as_tensor=False):
    """equalize_pilot_sequence + apply for several frames that start from the SAME initial taps, without
    frequency offset estimation: every training stage is one batched launch over the frames (one segment
    per frame, strided views of the capture).  Frame f gives what ``pilot_equaliser(frame=f)`` gives."""
    ref = np.atleast_2d(pilot_seq)
    npols = rx_signal.shape[0]
    frames = np.asarray(frames, dtype=np.int64)
    shifts = np.array(shiftfctrs, dtype=int)
    train_shifts = shifts - (Ntaps - synctaps) // 2 if Ntaps != synctaps else shifts.copy()
    # the reference adds the frame offset to the training window only when Ntaps != synctaps (:316-317)
    foff = os * frame_len * frames if Ntaps != synctaps else 0 * frames
    assert rx_signal.shape[-1] - (train_shifts.max() + foff.max()) > frame_len * os, \
        "You are trying to equalise an incomplete frame which does not work"
    if (methods[0] in theory.REAL_VALUED) != (methods[1] in theory.REAL_VALUED):
        raise ValueError("Using a complex and real-valued equalisation method is not supported")
    span = ref.shape[-1] * os + Ntaps - 1
    per_mode = np.unique(train_shifts).shape[0] > 1
    groups = [(train_shifts[i], [i]) for i in range(npols)] if per_mode else [(train_shifts[0], None)]
    wx = wxinit
    # only the taps are used: skip storing / downloading Niter * TrSyms errors per window where the backend can
    noerr = {"return_err": False} if getattr(be, "SUPPORTS_RETURN_ERR", False) else {}
    for start, modes in groups:                 # step 1: blind pre-convergence from the initial taps
        wx, _ = be.equalise_windows(rx_signal, start + foff, span, os, mu[0], M_pilot,
                                    wxy=wx if per_mode else wxinit, Ntaps=Ntaps, Niter=Niter, method=methods[0],
                                    adaptive_stepsize=adaptive_stepsize, modes=modes, **noerr)
    taps = wx
    for start, modes in groups:                 # step 2: both methods with the pilot sequence as training symbols
        taps, _ = be.equalise_windows(rx_signal, start + foff, span, os, mu[0], M_pilot, wxy=taps, Ntaps=Ntaps,
                                      Niter=Niter, method=methods[0], adaptive_stepsize=adaptive_stepsize,
                                      symbols=ref, modes=modes, **noerr)
        taps, _ = be.equalise_windows(rx_signal, start + foff, span, os, mu[1], 4 if per_mode else M_pilot, wxy=taps,
                                      Ntaps=Ntaps, Niter=Niter, method=methods[1],
                                      adaptive_stepsize=adaptive_stepsize, symbols=ref, modes=modes, **noerr)
    if not apply:
        return taps, None
    ashifts = shifts - (Ntaps - synctaps) // 2 if Ntaps != synctaps else shifts.copy()
    if np.min(ashifts) < 0:
        ashifts += os * frame_len
    aspan = frame_len * os + Ntaps - 1
    assert ashifts.max() + os * frame_len * (frames.max() + 1) < rx_signal.shape[-1] - (Ntaps - 1), \
        "Trying to equalise frame {}, but signal is not long enough".format(frames.max())
    if np.unique(ashifts).shape[0] > 1:
        mode_groups = np.arange(taps.shape[-3]).reshape(-1, npols).T
        kw = {"as_tensor": True} if as_tensor else {}
        rows = [be.apply_windows(rx_signal, ashifts[mode[0]] + os * frame_len * frames, aspan, os, taps, modes=mode, **kw)
                for mode in mode_groups]
        if as_tensor:
            import torch
            eq = torch.cat(rows, dim=1)
        else:
            eq = np.concatenate(rows, axis=1)   # (nframes, nmodes, frame_len)
    else:
        kw = {"as_tensor": True} if as_tensor else {}
        eq = be.apply_windows(rx_signal, ashifts[0] + os * frame_len * frames, aspan, os, taps, **kw)
    return taps, eq


def pilot_equaliser_nframes(rx_signal, pilot_seq, shiftfctrs, os, frame_len, mu, Ntaps, synctaps=17, apply=True,
                            foe_comp=True, frames=(0,), wxinit=None, backend=None, batched=False, **eqkwargs):
    """Pilot-based equalisation over several frames: array form of ``pilot_equaliser_nframes`` (:336-397).
    Frame 0 starts from ``wxinit`` (centre spike by default); every frame after it starts from frame 0's taps
    -- those frames are independent of one another and, without frequency offset estimation, are trained and
    filtered side by side (``batched``; one launch per stage for all of them).

    ``batched=False`` (default: the drop-in behaviour) is the reference's loop as it actually behaves: its ``wxinit``
    array is trained in place by every call (``equalisation.py:547``), so frame f warm-starts from frame f-1's
    pre-convergence taps rather than from frame 0's.  ``batched=True`` (opt-in; what ``pilot_receiver`` uses) is what
    that loop reads like -- every frame from frame 0's taps -- and frame f then equals ``pilot_equaliser(frame=f,
    wxinit=<copy of frame 0's taps>)``: independent frames, one launch per stage for all of them.
    Returns (list of taps per frame, equalised frames stacked along time or None, list of frequency offsets)."""
    if shiftfctrs is None:
        raise ValueError("The signal has to be synchronised to the frame first")
    be = _backend(backend)
    rx_signal = np.atleast_2d(rx_signal)
    if frames is None:
        frames = np.arange((rx_signal.shape[-1] - np.max(shiftfctrs)) // (os * frame_len))
    frames = np.atleast_1d(frames)
    assert rx_signal.shape[-1] - (np.max(shiftfctrs) + np.max(frames) * frame_len * os) > frame_len * os, \
        "The last frame must be complete for equalisation"
    if wxinit is not None:
        Ntaps = wxinit.shape[-1]
    mu2 = np.atleast_1d(mu)
    mu2 = np.repeat(mu2, 2) if len(mu2) == 1 else mu2
    taps_all, eq_all, foe_all = [], [], []
    k = 0
    while k < len(frames):
        # frames up to and including frame 0 one at a time (frame 0 redefines the initial taps, :386-387) ...
        rest = frames[k:]
        if batched and not foe_comp and hasattr(be, "equalise_windows") and len(rest) > 1 and 0 not in rest \
                and (abs(Ntaps - synctaps) % 2) == 0:
            # ... everything after it in one batch
            taps, eq = _pilot_frames_batched(be, rx_signal, pilot_seq, shiftfctrs, os, frame_len, mu2, Ntaps, synctaps,
                                             rest, wxinit, apply, **eqkwargs)
            for j in range(len(rest)):
                taps_all.append(taps[j])
                foe_all.append(np.zeros([rx_signal.shape[0], 1]))
                if apply:
                    eq_all.append(eq[j])
            break
        f = frames[k]
        ret = pilot_equaliser(rx_signal, pilot_seq, shiftfctrs, os, frame_len, mu, Ntaps, synctaps=synctaps,
                              apply=apply, foe_comp=foe_comp, wxinit=wxinit, frame=int(f), verbose=True,
                              backend=backend, **eqkwargs)
        if f == 0:
            wxinit = ret[0]
        taps_all.append(ret[0])
        if apply:
            eq_all.append(ret[1])
            foe_all.append(ret[2])
        else:
            foe_all.append(ret[1])
        k += 1
    return taps_all, (np.hstack(eq_all) if apply else None), foe_all


def pilot_receiver(rx_signal, pilot_seq, ph_pilots, idx_pil, frame_len, os, frames, mu=(1e-3, 1e-3), Ntaps=45,
                   num_average=5, M_pilot=4, methods=("cma", "sbd"), Niter=30, adaptive_stepsize=True,
                   sync_kwargs=None, to_host=True):
    """The pilot-based receiver chain of BASELINE config C4 with the capture resident on the GPU:

        sync2frame -> corr_foe -> pilot_equaliser_nframes(foe_comp=False) -> pilot_cpe(use_seq=False) per frame

    (``signals.py:1709-1747``, ``qampy/equalisation.py:336-397``, ``qampy/phaserec.py:156-192``).  One upload of
    the capture; the frame search, the frequency shift, the pilot trainings of all frames (frame 0 first, the
    others side by side from its taps), the FIRs and the per-frame phase recovery are CUDA launches on device
    data -- including the frame search's spectral frequency-offset estimate and sequence correlation (cuFFT); the host
    sees a few scalars and the final result.

    ``ph_pilots`` (nmodes, n_ph): the phase pilots of one frame; ``idx_pil`` (frame_len,) bool: pilot positions
    in a frame (the first ``pilot_seq.shape[-1]`` are the sequence).  Returns a dict: ``out`` (nmodes,
    nframes*frame_len) phase-compensated frames, ``eq`` the same before phase recovery, ``taps`` per frame,
    ``shiftfctrs``, ``foe``, ``mode_order``, ``sync_ok``."""
    import torch
    from . import device, equalisation as be
    from ._lib import require_device
    require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    Ed = rx_signal.to(dev) if torch.is_tensor(rx_signal) else \
        torch.from_numpy(np.ascontiguousarray(np.atleast_2d(np.asarray(rx_signal)))).to(dev)
    pilot_seq = np.atleast_2d(pilot_seq)
    sl = pilot_seq.shape[-1]
    # sync2frame (signals.py:1709-1741)
    eqargs = {"adaptive_stepsize": True, "Niter": 10, "method": "cma", "Ntaps": 17, "mu": 5e-3}
    eqargs.update(sync_kwargs or {})
    mu_s, synctaps = eqargs.pop("mu"), eqargs.pop("Ntaps")
    shift, foe, order, _, ok = frame_sync(None, pilot_seq, os, frame_len=frame_len, M_pilot=M_pilot, mu=mu_s,
                                          Ntaps=synctaps, rx_device=Ed, **eqargs)
    shift[shift < 0] += frame_len * os
    shiftf = shift[order]
    Ed = Ed[torch.as_tensor(order, device=dev)].contiguous()
    # corr_foe (signals.py:1744-1747)
    foe = np.asarray(foe)
    Ed = device.freq_shift(Ed, np.ones(foe.shape[0]) * np.mean(foe), os)
    # pilot equaliser: frame 0 (or the first listed frame) from centre-spike taps, the rest from its taps, batched
    frames = np.atleast_1d(np.asarray(frames, dtype=np.int64))
    mu2 = np.atleast_1d(mu)
    mu2 = np.repeat(mu2, 2) if len(mu2) == 1 else mu2
    kw = dict(M_pilot=M_pilot, Niter=Niter, adaptive_stepsize=adaptive_stepsize, methods=methods, as_tensor=True)
    t0, e0 = _pilot_frames_batched(be, Ed, pilot_seq, shiftf, os, frame_len, mu2, Ntaps, synctaps, frames[:1], None,
                                   True, **kw)
    taps, eqs = [t0[0]], [e0]
    if len(frames) > 1:
        t1, e1 = _pilot_frames_batched(be, Ed, pilot_seq, shiftf, os, frame_len, mu2, Ntaps, synctaps, frames[1:],
                                       t0[0], True, **kw)
        taps += [t1[j] for j in range(len(frames) - 1)]
        eqs.append(e1)
    eq = torch.cat(eqs, dim=0)                                   # (nframes, nmodes, frame_len)
    nfr, nmodes = eq.shape[0], eq.shape[1]
    # pilot_cpe(use_seq=False): phase pilots only (qampy/phaserec.py:183-186)
    idx = np.nonzero(np.asarray(idx_pil))[0][sl:]
    php = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(ph_pilots)[:, :idx.size])).to(dev).to(eq.dtype)
    rows = eq.reshape(nfr * nmodes, frame_len)
    out, _ = device.pilot_cpe(rows, idx, php.repeat(nfr, 1), num_average)
    out = out.reshape(nfr, nmodes, frame_len).permute(1, 0, 2).reshape(nmodes, nfr * frame_len)
    eq2 = eq.permute(1, 0, 2).reshape(nmodes, nfr * frame_len)
    res = dict(out=out, eq=eq2, taps=taps, shiftfctrs=shiftf, foe=foe, mode_order=order, sync_ok=ok)
    if to_host:
        res["out"], res["eq"] = out.cpu().numpy(), eq2.cpu().numpy()
    return res
