"""L2 drop-in for ``qampy/core/phaserecovery.py::bps`` (:93-159): same signature and returns
``(Eout, ph)``; the whole chain (distance search, running sum, arg-min, angle gather, unwrap,
rotation) is one fused CUDA kernel per call, all modes in one launch."""
import numpy as np
import torch

from . import _lib, device


def bps(E, Mtestangles, symbols, N, method="pyt", **kwargs):
    """Blind phase search.  ``method`` is accepted for signature compatibility ("pyt"/"pyx"/"af"/"py"
    all run the CUDA kernel; anything else raises like the reference).  ``accum="windowed"`` (keyword, not in the
    reference): direct 2N-term window sums instead of the reference's running sum -- see ``device.bps``."""
    if method.lower() not in ("pyx", "af", "py", "pyt", "cuda"):
        raise ValueError("Method needs to be 'pyx', 'py' or 'af'")
    Ein = E
    E = np.asarray(E)
    if E.dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
        raise TypeError("qampy_b200 bps needs a complex64/complex128 signal, got %s" % E.dtype)
    _lib.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    Ew = np.atleast_2d(E)
    tables = device.BpsTables(int(Mtestangles), symbols, E.dtype.type, dev)
    Ed = torch.from_numpy(np.ascontiguousarray(Ew)).to(dev)
    out, ph, _ = device.bps(Ed, tables, int(N), want_idx=False, accum=kwargs.get("accum", "exact"))
    ph = ph.cpu().numpy()
    # keep the SignalObject subclass (and its attributes) of the input, like Ew*np.exp(1j*ph) does; the device result
    # is copied straight into the new array (no intermediate host copy)
    Eout = np.empty_like(np.atleast_2d(Ein), dtype=E.dtype, order='C')
    torch.from_numpy(np.asarray(Eout)).copy_(out)
    if E.ndim == 1:
        return Eout.flatten(), ph.flatten()
    return Eout, ph


def bps_twostage(E, Mtestangles, symbols, N, B=4, method="pyt", **kwargs):
    """Two-stage blind phase search, drop-in for ``qampy/core/phaserecovery.py::bps_twostage`` (:222-288,
    after Zhuge et al., OFC 2011): a coarse search over ``Mtestangles`` angles, then a search over ``B``
    angles around every symbol's coarse estimate.  Both index searches run on the GPU (the second one
    with a per-symbol angle table); angle tables, gathers, ``np.unwrap(4*ph, discont=pi)/4`` over the whole
    array and the rotation follow the reference line by line in NumPy."""
    from . import pythran_dsp
    if method.lower() not in ("pyx", "af", "py", "pyt", "cuda"):
        raise ValueError("Method needs to be 'pyx', 'py' or 'af'")
    Ein = E
    E = np.asarray(E)
    if E.dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
        raise TypeError("qampy_b200 bps_twostage needs a complex64/complex128 signal, got %s" % E.dtype)
    rt = E.real.dtype
    angles = np.linspace(-np.pi / 4, np.pi / 4, Mtestangles, endpoint=False, dtype=rt).reshape(1, -1)   # :270
    Ew = np.atleast_2d(E)
    ph_out = []
    for i in range(Ew.shape[0]):
        idx = pythran_dsp.bps(np.copy(Ew[i]), angles, symbols, N)
        ph = pythran_dsp.select_angles(np.copy(angles), idx)
        b = np.linspace(-B / 2, B / 2, B)
        phn = (ph[:, np.newaxis] + b[np.newaxis, :] / (B * Mtestangles) * np.pi / 2).astype(rt)   # :277
        idx2 = pythran_dsp.bps(np.copy(Ew[i]), phn, symbols, N)
        phf = pythran_dsp.select_angles(np.copy(phn), idx2)
        ph_out.append(np.unwrap(phf * 4, discont=np.pi * 4 / 4) / 4)                              # :280
    ph_out = np.asarray(ph_out, dtype=rt)
    En = np.atleast_2d(Ein).astype(E.dtype) * np.exp(1.j * ph_out)
    if E.ndim == 1:
        return En.flatten(), ph_out.flatten()
    return En, ph_out


def viterbiviterbi(E, N, M):
    """Viterbi-Viterbi blind phase recovery for an M-PSK signal, drop-in for
    ``qampy/core/phaserecovery.py::viterbiviterbi`` (:40-79): M-th power, sliding sum over ``N`` symbols, unwrapped
    angle, rotation; three CUDA kernels for all modes (``csrc/vv.cu``).  Returns ``(Eout, phase_est)`` like the
    reference, including its quirk that for a 2-D input ``phase_est`` is the 1-D estimate of the LAST mode only
    (:77-79); ``qampy_b200.device.viterbiviterbi`` returns the estimates of every row."""
    Ein = E
    E = np.asarray(E)
    if E.dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
        raise TypeError("qampy_b200 viterbiviterbi needs a complex64/complex128 signal, got %s" % E.dtype)
    _lib.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    Ew = np.atleast_2d(E)
    out, ph = device.viterbiviterbi(torch.from_numpy(np.ascontiguousarray(Ew)).to(dev), int(N), int(M))
    Eout = np.zeros_like(np.atleast_2d(Ein), dtype=E.dtype)           # keeps the SignalObject subclass (:60)
    Eout[...] = out.cpu().numpy()
    phase_est = ph[-1].cpu().numpy()
    if E.ndim == 1:
        return Eout.flatten(), phase_est.flatten()
    return Eout, phase_est
