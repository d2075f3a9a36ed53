"""Synthetic impaired coherent-receiver input for tests and bench.py (``"data": "synthetic"``).

Not on the hot path: plain torch ops, runs on CPU or on the GPU.  The recipe mirrors what the
reference's generators do for the BASELINE configs: Gray-agnostic random M-QAM
symbols at unit power -> root-raised-cosine pulse shaping to ``os`` samples per symbol
(``qampy/core/resample.py:73-126``) -> AWGN at a per-symbol SNR (``core/impairments.py:210-233``:
noise std = sqrt(P * os) * 10^(-snr/20)) -> first-order PMD in the frequency domain
(``core/impairments.py:94-131``) -> optional Wiener phase noise (``:133-186``).  The reference's own
generators are used for the golden fixtures (``tests/golden``); this one exists because the GPU box
has no reference checkout and config C3/C5 inputs are too large to synthesise on the host.
"""
import math

import numpy as np
import torch

from .theory import normalised_symbols


def _rrc_freq(n, os, beta, device):
    """|H(f)| of a root-raised-cosine with symbol rate 1 and sampling rate ``os`` on an n-point FFT grid."""
    f = torch.fft.fftfreq(n, d=1.0 / os, device=device, dtype=torch.float64).abs()
    H = torch.zeros(n, dtype=torch.float64, device=device)
    lo, hi = (1 - beta) / 2, (1 + beta) / 2
    H[f <= lo] = 1.0
    if beta > 0:
        band = (f > lo) & (f <= hi)
        H[band] = torch.sqrt(0.5 * (1 + torch.cos(math.pi / beta * (f[band] - lo))))
    return H


def synth_signal(M, nsym, nmodes=2, os=2, beta=0.1, snr_db=28.0, theta=math.pi / 5.6, dgd=40e-12, fb=40e9,
                 linewidth=None, seed=0, dtype=torch.complex64, device="cpu"):
    """Returns ``(E, symbols)``: E (nmodes, nsym*os) impaired signal, symbols (nmodes, nsym) sent."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    alphabet = torch.from_numpy(normalised_symbols(M)).to(device)
    idx = torch.randint(0, M, (nmodes, nsym), generator=gen, device=device)
    syms = alphabet[idx]
    return shape_and_impair(syms, os, beta, snr_db, theta, dgd, fb, linewidth, seed, gen, dtype)


def shape_and_impair(syms, os=2, beta=0.1, snr_db=28.0, theta=math.pi / 5.6, dgd=40e-12, fb=40e9, linewidth=None,
                     seed=0, gen=None, dtype=torch.complex64, freq_off=None, tx_linewidth=None):
    """Pulse shaping and channel for given symbols (nmodes, nsym): RRC to ``os`` samples per symbol, PMD, AWGN,
    optional frequency offset (Hz) and Wiener phase noise -- ``linewidth`` on the received rows (after the
    polarisation mixing), ``tx_linewidth`` on the transmitted polarisations before it, which is where the
    reference's ``simulate_transmission`` puts it (``impairments.py:159-160`` before ``apply_PMD``).
    Returns (E, syms)."""
    device = syms.device
    nmodes, nsym = syms.shape
    if gen is None:
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed) + 17)
    n = nsym * os
    x = torch.zeros((nmodes, n), dtype=torch.complex128, device=device)
    x[:, ::os] = syms
    X = torch.fft.fft(x, dim=1) * _rrc_freq(n, os, beta, device)
    if tx_linewidth:
        x = apply_phase_noise(torch.fft.ifft(X, dim=1), tx_linewidth, os * fb, seed=seed + 2)
        X = torch.fft.fft(x.to(torch.complex128), dim=1)
    if nmodes == 2 and theta is not None:
        # first-order PMD: rotate, delay the axes by +-dgd/2, rotate back
        omega = 2 * math.pi * torch.fft.fftfreq(n, d=1.0 / (os * fb), device=device, dtype=torch.float64)
        c, s = math.cos(theta), math.sin(theta)
        a = c * X[0] + s * X[1]
        b = -s * X[0] + c * X[1]
        a = a * torch.exp(-0.5j * omega * dgd)
        b = b * torch.exp(0.5j * omega * dgd)
        X = torch.stack([c * a - s * b, s * a + c * b])
    x = torch.fft.ifft(X, dim=1)
    x = x / torch.sqrt((x.abs() ** 2).mean(dim=1, keepdim=True))
    if snr_db is not None:
        sigma = math.sqrt(os) * 10 ** (-snr_db / 20)
        noise = torch.randn((nmodes, n, 2), generator=gen, device=device, dtype=torch.float64)
        x = x + sigma / math.sqrt(2) * torch.view_as_complex(noise)
    if freq_off:
        t = torch.arange(n, device=device, dtype=torch.float64)
        x = x * torch.exp(2j * math.pi * freq_off / (os * fb) * t)
    if linewidth:
        x = apply_phase_noise(x, linewidth, os * fb, seed=seed + 1)
    return x.to(dtype), syms.to(dtype)


def synth_pilot_signal(M, frame_len, pilot_seq_len, pilot_ins_rat, nframes, nmodes=2, Mpilots=4, os=2, beta=0.01,
                       snr_db=30.0, theta=math.pi / 3.731, dgd=10e-12, fb=24e9, freq_off=None, linewidth=None,
                       delay=0, seed=0, dtype=torch.complex64, device="cpu"):
    """Pilot-framed signal for the pilot-based receiver (BASELINE config C4; the layout of the reference's
    ``SignalWithPilots``): every frame starts with a ``pilot_seq_len`` QPSK pilot sequence, after it every
    ``pilot_ins_rat``-th symbol is a phase pilot, the rest is M-QAM payload; pilots are the same in every
    frame, payload is not.  ``delay`` rolls the signal by that many samples (unknown frame start).
    Returns a dict: E (nmodes, nframes*frame_len*os), symbols (nmodes, nframes*frame_len), pilot_seq
    (nmodes, pilot_seq_len), ph_pilots (nmodes, n_ph), idx_pil (frame_len,) bool."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    assert (frame_len - pilot_seq_len) % pilot_ins_rat == 0
    idx_pil = np.zeros(frame_len, dtype=bool)
    idx_pil[:pilot_seq_len] = True
    idx_pil[pilot_seq_len::pilot_ins_rat] = True
    npil = int(idx_pil.sum())
    pal = torch.from_numpy(normalised_symbols(Mpilots)).to(device)
    dal = torch.from_numpy(normalised_symbols(M)).to(device)
    pilots = pal[torch.randint(0, Mpilots, (nmodes, npil), generator=gen, device=device)]
    frames = []
    mask = torch.from_numpy(idx_pil).to(device)
    for _ in range(nframes):
        fr = torch.empty((nmodes, frame_len), dtype=torch.complex128, device=device)
        fr[:, mask] = pilots
        fr[:, ~mask] = dal[torch.randint(0, M, (nmodes, frame_len - npil), generator=gen, device=device)]
        frames.append(fr)
    syms = torch.cat(frames, dim=1)
    E, _ = shape_and_impair(syms, os, beta, snr_db, theta, dgd, fb, None, seed, gen, dtype, freq_off,
                            tx_linewidth=linewidth)
    if delay:
        E = torch.roll(E, int(delay), dims=1)
    return dict(E=E, symbols=syms.to(dtype), pilot_seq=pilots[:, :pilot_seq_len].to(dtype),
                ph_pilots=pilots[:, pilot_seq_len:].to(dtype), idx_pil=idx_pil)


def synth_capture(M, nsym, block=10 ** 7, seed=0, device="cpu", dtype=torch.complex64, **kwargs):
    """Long capture (BASELINE config C5: 1e9 samples) written block by block into one preallocated array: every block
    of ``block`` symbols is an independent :func:`synth_signal` realisation (seed + block index), so the FFT-based
    channel never needs more than one block of double-precision temporaries.  Blocks are not continuous with each
    other; segments are trained independently anyway and one in ``block / segment`` sees a boundary.
    Returns ``(E, symbols_of_block_0)``."""
    os_ = int(kwargs.get("os", 2))
    nmodes = int(kwargs.get("nmodes", 2))
    E = torch.empty((nmodes, nsym * os_), dtype=dtype, device=device)
    syms0 = None
    for b, a in enumerate(range(0, nsym, block)):
        n = min(block, nsym - a)
        Eb, sb = synth_signal(M, n, seed=seed + b, device=device, dtype=dtype, **kwargs)
        E[:, a * os_:(a + n) * os_] = Eb
        if syms0 is None:
            syms0 = sb
        del Eb, sb
    return E, syms0


def apply_phase_noise(x, linewidth, fs, seed=0):
    """Wiener phase walk with variance 2*pi*linewidth/fs per sample, independent per row."""
    gen = torch.Generator(device=x.device)
    gen.manual_seed(int(seed))
    var = 2 * math.pi * linewidth / fs
    steps = torch.randn(x.shape, generator=gen, device=x.device, dtype=torch.float64) * math.sqrt(var)
    ph = torch.cumsum(steps, dim=-1)
    return (x.to(torch.complex128) * torch.exp(1j * ph)).to(x.dtype)


def synth_numpy(*args, **kwargs):
    E, s = synth_signal(*args, **kwargs)
    return E.cpu().numpy(), s.cpu().numpy()


def decide(E, M):
    """Nearest-symbol decisions (host, NumPy) on the unit-power M-QAM alphabet."""
    alphabet = normalised_symbols(M).astype(np.complex64)
    E = np.asarray(E)
    out = np.empty(E.shape, dtype=np.complex64)
    flat, oflat = E.reshape(-1), out.reshape(-1)
    for a in range(0, flat.size, 65536):
        c = flat[a:a + 65536]
        oflat[a:a + 65536] = alphabet[np.argmin(np.abs(c[:, None] - alphabet[None, :]), axis=1)]
    return out


def ser(E, syms, M, max_delay=64, probe=2000):
    """Symbol error rate of equalised 1-sps ``E`` against the sent ``syms``, after resolving what a
    blind receiver leaves open: which sent row each output row carries, a multiple-of-pi/2 rotation
    and an integer symbol delay (found on the first ``probe`` symbols, then applied to all)."""
    dec = decide(E, M)
    syms = np.asarray(syms)
    total = []
    for r in range(dec.shape[0]):
        best = (2.0, 0, 0, 0)
        n = min(probe, dec.shape[1] - max_delay)
        for src in range(syms.shape[0]):
            for rot in range(4):
                ref = syms[src] * (1j ** rot)
                for d in range(0, max_delay):
                    e = float(np.mean(np.abs(dec[r, :n] - ref[d:d + n]) > 1e-3))
                    if e < best[0]:
                        best = (e, src, rot, d)
        _, src, rot, d = best
        ref = syms[src] * (1j ** rot)
        m = min(dec.shape[1], ref.size - d)
        total.append(float(np.mean(np.abs(dec[r, :m] - ref[d:d + m]) > 1e-3)))
    return float(np.mean(total))


def _nearest_index(x, alphabet, chunk=1 << 18):
    """Index of the nearest alphabet point for every element of the complex tensor ``x`` (any shape)."""
    flat = x.reshape(-1)
    out = torch.empty(flat.shape, dtype=torch.int16, device=x.device)
    for a in range(0, flat.numel(), chunk):
        c = flat[a:a + chunk]
        d = (c.real[:, None] - alphabet.real[None, :]) ** 2 + (c.imag[:, None] - alphabet.imag[None, :]) ** 2
        out[a:a + chunk] = torch.argmin(d, dim=1).to(torch.int16)
    return out.reshape(x.shape)


def ser_segments(out, syms, M, firsts, probe=512, max_delay=32, seg_chunk=128):
    """Symbol errors of EVERY segment of a segmented receiver run, on the tensors' own device (torch ops; this is
    the bench's sanity gate, not a product path).

    ``out`` (nseg, nmodes, S): recovered 1-sps symbols; ``syms`` (nmodes, nsym): what was sent; ``firsts`` (nseg,):
    index of each segment's first output symbol in the capture.  What a blind receiver leaves open -- which sent
    row an output row carries, a multiple-of-pi/2 rotation, the equaliser's symbol delay -- is resolved PER (segment,
    row) on its first ``probe`` symbols (candidates: every sent row x 4 rotations x delays [0, max_delay)) and then
    held for the whole segment.  Returns (errors (nseg, nmodes) int64, compared (nseg, nmodes) int64)."""
    dev = out.device
    nseg, nmodes, S = out.shape
    nsrc, nsym = syms.shape
    alphabet = torch.from_numpy(normalised_symbols(M)).to(dev).to(torch.complex64)
    sent = _nearest_index(syms.to(torch.complex64), alphabet).to(torch.int64)                 # (nsrc, nsym)
    # perm[r][k]: index of alphabet[k] * i^r
    perm = torch.stack([_nearest_index(alphabet * (1j ** r), alphabet).to(torch.int64) for r in range(4)])
    firsts = torch.as_tensor(firsts, dtype=torch.int64, device=dev)
    P = min(probe, S)
    errors = torch.zeros((nseg, nmodes), dtype=torch.int64, device=dev)
    compared = torch.zeros((nseg, nmodes), dtype=torch.int64, device=dev)
    ar_p = torch.arange(P, device=dev)
    ar_d = torch.arange(max_delay, device=dev)
    ar_s = torch.arange(S, device=dev)
    for a in range(0, nseg, seg_chunk):
        o = out[a:a + seg_chunk]
        n = o.shape[0]
        dec = _nearest_index(o.to(torch.complex64), alphabet).to(torch.int64)               # (n, nmodes, S)
        f = firsts[a:a + n]
        pos = (f[:, None, None] + ar_d[None, :, None] + ar_p[None, None, :]).clamp_(max=nsym - 1)    # (n, D, P)
        cand = perm[:, sent[:, pos]]                                                         # (4, nsrc, n, D, P)
        mism = (cand[None] != dec[:, :, :P].permute(1, 0, 2)[:, None, None, :, None, :]).sum(-1)   # (nmodes,4,nsrc,n,D)
        best = mism.permute(3, 0, 1, 2, 4).reshape(n, nmodes, -1).argmin(-1)                 # (n, nmodes)
        rot = best // (nsrc * max_delay)
        src = (best // max_delay) % nsrc
        dly = best % max_delay
        p_all = f[:, None, None] + dly[:, :, None] + ar_s[None, None, :]                    # (n, nmodes, S)
        valid = p_all < nsym
        ref = sent.reshape(-1)[(src[:, :, None] * nsym + p_all.clamp(max=nsym - 1)).reshape(-1)].reshape(n, nmodes, S)
        ref = perm.reshape(-1)[(rot[:, :, None] * perm.shape[1] + ref).reshape(-1)].reshape(n, nmodes, S)
        errors[a:a + n] = ((ref != dec) & valid).sum(-1)
        compared[a:a + n] = valid.sum(-1)
    return errors, compared
