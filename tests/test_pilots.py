"""Pilot-based receiver: the array-level functions of qampy_b200/pilots.py against
vectors produced by the reference (tests/golden/g9_pilot_rx.npz, make_golden_next.py).

CPU tests run the host-side glue on top of the CPU oracle (backend injection); the GPU tests run the same
functions on the CUDA equaliser, including the batched frame search."""
import types
import warnings

import numpy as np
import pytest

import cpu_oracle as co
from qampy_b200 import pilots


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2))) if np.size(a) else 0.0


ORACLE = types.SimpleNamespace(equalise_signal=co.equalise_signal, apply_filter=co.apply_filter)


def _cuda_backend():
    import qampy_b200.equalisation as eq
    from qampy_b200 import _lib
    _lib.require_device()
    return eq


def _run_chain(g, be, tol_taps, tol_sig):
    rx, seq, fl, osf = g["rx"], g["pilot_seq"], int(g["frame_len"]), int(g["os"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # frame search with the reference's sync2frame defaults (signals.py:1709-1741)
        sf, foe, order, wx1, ok = pilots.frame_sync(rx, seq, osf, frame_len=fl, M_pilot=4, mu=5e-3, Ntaps=17,
                                                    adaptive_stepsize=True, Niter=10, method="cma", backend=be)
    assert bool(ok) == bool(g["fs_ok"])
    assert np.array_equal(sf, g["fs_shift"]) and np.array_equal(order, g["fs_order"])
    assert np.allclose(foe, g["fs_foe"], rtol=0, atol=1e-12)
    assert np.max(np.abs(wx1 - g["fs_wx1"])) < tol_taps
    aligned, shiftf, foe2, _, _ = pilots.sync2frame(rx, seq, osf, fl, backend=be)
    assert np.array_equal(shiftf, g["shiftfctrs"])
    rx3 = pilots.corr_foe(aligned, foe2, osf)
    assert rx3.dtype == g["rx_synced_head"].dtype and rms(rx3[:, :4096] - g["rx_synced_head"]) < 1e-6
    # pilot equaliser, frame 0 (qampy/equalisation.py:266-334), blind and data-aided second stage
    for tag, methods in (("sbd", ("cma", "sbd")), ("data", ("cma", "sbd_data"))):
        taps, eq0 = pilots.pilot_equaliser(rx3, seq, shiftf, osf, fl, (1e-3, 1e-3), 45, synctaps=17, foe_comp=False,
                                           methods=methods, backend=be)
        assert taps.dtype == g["taps_" + tag].dtype
        assert np.max(np.abs(taps - g["taps_" + tag])) < tol_taps, tag
        if tag == "sbd":
            assert eq0.shape == g["eq_frame0"].shape and rms(eq0 - g["eq_frame0"]) < tol_sig
    return rx3, shiftf


def test_host_side_estimators_match_reference(golden):
    """pilot_based_cpe_new, pilot_based_foe and the small helpers on the reference's own intermediate arrays."""
    g = golden("g9_pilot_rx")
    fl, sl = int(g["frame_len"]), g["pilot_seq"].shape[-1]
    idx = np.nonzero(g["idx_pil"])[0][sl:]
    out, trace = pilots.pilot_based_cpe_new(g["eq_frame0"], g["ph_pilots"], idx, fl, num_average=5, nframes=1)
    assert out.dtype == g["cpe_out"].dtype and trace.dtype == g["cpe_trace"].dtype
    assert np.array_equal(trace, g["cpe_trace"]) and np.array_equal(out, g["cpe_out"])
    foe, per_mode, cond = pilots.pilot_based_foe(g["eq_frame0"][:, :sl], g["pilot_seq"])
    assert foe == g["pfoe"] and np.array_equal(per_mode, g["pfoe_mode"]) and np.array_equal(cond, g["pfoe_cond"])
    # decisions on the phase-corrected payload: the chain the vectors came from demodulates cleanly
    data = out[:, ~g["idx_pil"]]
    d = np.abs(data[:, :, None] - g["coded"][None, None, :])
    dec = g["coded"][np.argmin(d, axis=-1)]
    assert np.mean(dec != g["symbols_tx"][:, :dec.shape[1]]) < 1e-3
    x = np.arange(40.).reshape(2, 20)
    assert np.allclose(pilots.moving_average(x, 5), np.stack([np.convolve(r, np.ones(5) / 5, "valid") for r in x]))
    sf = pilots.correct_shifts(np.array([100, 120]), (17, 45), 2)
    assert list(sf) == [86, 106]
    with pytest.raises(ValueError):
        pilots.correct_shifts(np.array([1]), (17, 44), 2)


def test_pilot_receiver_on_oracle_backend(golden):
    """The whole array-level chain with the CPU oracle doing the equaliser work."""
    g = golden("g9_pilot_rx")
    _run_chain(g, ORACLE, 5e-5, 5e-5)


def test_frame_sync_rejects_unsupported_methods(golden):
    g = golden("g9_pilot_rx")
    for m in ("cma_real", "sbd_data"):
        with pytest.raises(ValueError):
            pilots.frame_sync(g["rx"], g["pilot_seq"], 2, frame_len=int(g["frame_len"]), method=m, backend=ORACLE)
    with pytest.raises(ValueError):
        pilots.equalize_pilot_sequence(g["rx"], g["pilot_seq"], [0, 0], 2, methods=("cma", "dd_real"), backend=ORACLE)


@pytest.mark.gpu
def test_pilot_receiver_on_cuda(golden):
    """Same chain on the CUDA equaliser (batched frame search, L2 equalise_signal / apply_filter)."""
    g = golden("g9_pilot_rx")
    be = _cuda_backend()
    rx3, shiftf = _run_chain(g, be, 1e-4, 1e-4)
    # the batched window trainer (throughput layout of the trainer) equals per-window equalise_signal calls (one
    # capture each: latency layout -- another kernel since the adaptive step size runs in the look-ahead form, so
    # equal to rounding, not bit for bit)
    starts = np.arange(2, 9) * 512
    taps, errs = be.equalise_windows(g["rx"], starts, 1024, 2, 5e-3, 4, Ntaps=17, Niter=10, method="cma",
                                     adaptive_stepsize=True)
    for k, s0 in enumerate(starts):
        w, e = be.equalise_signal(g["rx"][:, s0:s0 + 1024], 2, 5e-3, 4, Ntaps=17, Niter=10, method="cma",
                                  adaptive_stepsize=True)
        assert np.max(np.abs(w - taps[k])) < 1e-5 and rms(e - errs[k]) < 1e-5
    # several frames: frame 0 from centre-spike taps, the others from frame 0's taps
    fl, osf = int(g["frame_len"]), int(g["os"])
    taps_all, eq_all, _ = pilots.pilot_equaliser_nframes(rx3, g["pilot_seq"], shiftf, osf, fl, (1e-3, 1e-3), 45,
                                                        synctaps=17, foe_comp=False, frames=[0, 1],
                                                        methods=("cma", "sbd"), backend=be)
    assert eq_all.shape == (2, 2 * fl) and rms(eq_all[:, :fl] - g["eq_frame0"]) < 1e-4
    # both frames demodulate: phase recovery per frame with that frame's own phase pilots (ph_pilots holds
    # every pilot after the first frame's sequence, in time order)
    sl = g["pilot_seq"].shape[-1]
    idx = np.nonzero(g["idx_pil"])[0][sl:]
    nph, ndata = idx.size, int(np.sum(~g["idx_pil"]))
    for f in range(2):
        p0 = f * (sl + nph)
        out, _ = pilots.pilot_based_cpe_new(eq_all[:, f * fl:(f + 1) * fl], g["ph_pilots"][:, p0:p0 + nph], idx, fl,
                                            num_average=5, nframes=1)
        data = out[:, ~g["idx_pil"]]
        d = np.abs(data[:, :, None] - g["coded"][None, None, :])
        dec = g["coded"][np.argmin(d, axis=-1)]
        assert np.mean(dec != g["symbols_tx"][:, f * ndata:(f + 1) * ndata]) < 1e-3, f


def _payload_ser(eq_frames, d, fl, sl, M, tx_frames):
    """SER of the payload after per-frame pilot CPE; eq_frames (nmodes, nframes*fl) starts at tx frame tx_frames[0]."""
    from qampy_b200 import theory
    alphabet = theory.normalised_symbols(M).astype(np.complex64)
    idx_pil = d["idx_pil"]
    idx = np.nonzero(idx_pil)[0][sl:]
    sy = d["symbols"].cpu().numpy()
    php = d["ph_pilots"].cpu().numpy()
    errs = []
    for k, tf in enumerate(tx_frames):
        out, _ = pilots.pilot_based_cpe_new(eq_frames[:, k * fl:(k + 1) * fl], php, idx, fl, num_average=5, nframes=1)
        data = out[:, ~idx_pil]
        ref = sy[:, tf * fl:(tf + 1) * fl][:, ~idx_pil]
        dec = np.empty_like(data)
        for a in range(0, data.shape[1], 8192):
            c = data[:, a:a + 8192]
            dec[:, a:a + 8192] = alphabet[np.argmin(np.abs(c[:, :, None] - alphabet[None, None, :]), axis=-1)]
        errs.append(float(np.mean(np.abs(dec - ref) > 1e-3)))
    return errs


@pytest.mark.gpu
def test_batched_frames_equal_frame_by_frame():
    """pilot_equaliser_nframes: the frames trained side by side in one launch per stage give what
    the frame-by-frame loop gives (to rounding: the two run different layouts of the trainer), and the chain demodulates a synthetic pilot-framed signal."""
    from qampy_b200 import synth
    be = _cuda_backend()
    M, fl, sl, rat, nfr = 16, 2 ** 13, 512, 32, 6
    d = synth.synth_pilot_signal(M, fl, sl, rat, nfr, snr_db=26, freq_off=80e6, linewidth=100e3, delay=1400, seed=3)
    rx, seq = d["E"].numpy(), d["pilot_seq"].numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        al, shiftf, foe, _, ok = pilots.sync2frame(rx, seq, 2, fl, backend=be)
    assert ok and abs(int(shiftf[0]) - 1400) <= 16 and shiftf[0] == shiftf[1]
    assert abs(float(foe[0, 0]) - 80e6 / 24e9) < 3e-4      # coarse: 4th-power spectrum of a 1k-symbol stretch
    rx3 = pilots.corr_foe(al, foe, 2)
    args = (rx3, seq, shiftf, 2, fl, (1e-3, 1e-3), 45)
    kw = dict(synctaps=17, foe_comp=False, frames=[0, 1, 2, 3, 4], methods=("cma", "sbd"), backend=be)
    t_b, eq_b, _ = pilots.pilot_equaliser_nframes(*args, batched=True, **kw)
    assert len(t_b) == 5 and eq_b.shape == (2, 5 * fl)
    for f in range(1, 5):       # every later frame == pilot_equaliser on that frame alone from frame 0's taps
        t1, e1 = pilots.pilot_equaliser(*args, synctaps=17, foe_comp=False, wxinit=t_b[0].copy(), frame=f,
                                        methods=("cma", "sbd"), backend=be)
        # (alone: one capture -> latency layout of the trainer; in the batch: throughput layout -- equal to rounding)
        assert np.max(np.abs(t1 - t_b[f])) < 1e-5 and rms(e1 - eq_b[:, f * fl:(f + 1) * fl]) < 1e-5, f
    # the reference's own loop warm-starts frame f from frame f-1 (its wxinit array is trained in place,
    # equalisation.py:547); batched=False reproduces that chain, and both demodulate
    t_s, eq_s, _ = pilots.pilot_equaliser_nframes(*args, batched=False, **kw)
    assert rms(eq_s[:, :fl] - eq_b[:, :fl]) < 1e-5         # frame 0 is the same in both
    # (a 450-symbol pilot training occasionally leaves a frame unconverged -- the algorithm's, and the
    # reference's, behaviour; the chain as a whole has to demodulate)
    assert sorted(_payload_ser(eq_b, d, fl, sl, M, range(5)))[3] < 2e-3
    assert sorted(_payload_ser(eq_s, d, fl, sl, M, range(5)))[3] < 2e-3


@pytest.mark.gpu
def test_full_size_c4_pilot_receiver():
    """BASELINE config C4 at full size: dual-pol 256-QAM, 61 frames of 2**16 symbols (4e6 symbols), pilot sequence
    2**10, one phase pilot per 32, ntaps 45 (17 for the frame search).  Frame sync + pilot equaliser over 59
    frames side by side + pilot CPE: the frames demodulate, and a frame taken out of the batch is identical to rounding
    to the same frame equalised alone and matches the CPU oracle."""
    import torch
    from qampy_b200 import synth
    be = _cuda_backend()
    M, fl, sl, rat, nfr = 256, 2 ** 16, 2 ** 10, 32, 61
    d = synth.synth_pilot_signal(M, fl, sl, rat, nfr, snr_db=35, freq_off=100e6, linewidth=100e3, delay=4000, seed=11,
                                 device=torch.device("cuda", 0))
    rx, seq = d["E"].cpu().numpy(), d["pilot_seq"].cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        al, shiftf, foe, _, ok = pilots.sync2frame(rx, seq, 2, fl, backend=be)
    assert ok and abs(int(shiftf[0]) - 4000) <= 16
    rx3 = pilots.corr_foe(al, foe, 2)
    frames = list(range(59))
    taps, eq, _ = pilots.pilot_equaliser_nframes(rx3, seq, shiftf, 2, fl, (1e-3, 1e-3), 45, synctaps=17, foe_comp=False,
                                                 frames=frames, methods=("cma", "sbd"), backend=be, batched=True)
    assert eq.shape == (2, 59 * fl) and np.isfinite(eq).all()
    ser = _payload_ser(eq[:, :4 * fl], d, fl, sl, M, range(4)) + _payload_ser(eq[:, 57 * fl:59 * fl], d, fl, sl, M, (57, 58))
    assert sorted(ser)[4] < 5e-2, ser      # 256-QAM at 35 dB with pilot CPE: about 1-2 %
    f = 41
    t1, e1 = pilots.pilot_equaliser(rx3, seq, shiftf, 2, fl, (1e-3, 1e-3), 45, synctaps=17, foe_comp=False,
                                    wxinit=taps[0].copy(), frame=f, methods=("cma", "sbd"), backend=be)
    assert np.max(np.abs(t1 - taps[f])) < 1e-5 and rms(e1 - eq[:, f * fl:(f + 1) * fl]) < 1e-5
    t2, e2 = pilots.pilot_equaliser(rx3, seq, shiftf, 2, fl, (1e-3, 1e-3), 45, synctaps=17, foe_comp=False,
                                    wxinit=taps[0].copy(), frame=f, methods=("cma", "sbd"), backend=ORACLE)
    assert np.max(np.abs(t2 - taps[f])) < 1e-4 and rms(e2 - e1) < 1e-4


@pytest.mark.gpu
def test_device_resident_chain_matches_array_chain(golden):
    """pilots.pilot_receiver (capture resident on the GPU: frequency shift and pilot phase recovery as CUDA
    kernels) against the NumPy-glued chain on the reference's golden signal, and the two element-wise kernels
    against NumPy directly."""
    import torch
    from qampy_b200 import device
    g = golden("g9_pilot_rx")
    be = _cuda_backend()
    rx, seq, fl, osf = g["rx"], g["pilot_seq"], int(g["frame_len"]), int(g["os"])
    sl = seq.shape[-1]
    idx = np.nonzero(g["idx_pil"])[0][sl:]
    php = g["ph_pilots"][:, :idx.size]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = pilots.pilot_receiver(rx, seq, php, g["idx_pil"], fl, osf, frames=[0, 1])
    assert res["sync_ok"] and np.array_equal(res["shiftfctrs"], g["shiftfctrs"])
    assert rms(res["eq"][:, :fl] - g["eq_frame0"]) < 1e-4
    assert rms(res["out"][:, :fl] - g["cpe_out"]) < 1e-4
    assert np.max(np.abs(res["taps"][0] - g["taps_sbd"])) < 1e-4
    # kernels alone: frequency shift == comp_freq_offset, pilot CPE == pilot_based_cpe_new (NumPy, float64 ramp)
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(rx[:, :50000].copy()).to(dev)
    foe = np.array([[1.2345e-3], [-2.5e-4]])
    assert rms(device.freq_shift(x, foe, 2).cpu().numpy() - pilots.comp_freq_offset(rx[:, :50000], foe, 2)) < 2e-6
    eq0 = torch.from_numpy(g["eq_frame0"].copy()).to(dev)
    out, tr = device.pilot_cpe(eq0, idx, torch.from_numpy(php.copy()).to(dev), 5, want_trace=True)
    assert rms(out.cpu().numpy() - g["cpe_out"]) < 2e-6
    assert np.max(np.abs(tr.cpu().numpy() - g["cpe_trace"].real)) < 2e-6
