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
    # the batched window trainer equals per-window equalise_signal calls
    starts = np.arange(2, 9) * 512
    taps, errs = be.equalise_windows(g["rx"], starts, 1024, 2, 5e-3, 4, Ntaps=17, Niter=10, method="cma",
                                     adaptive_stepsize=True)
    for k, s0 in enumerate(starts):
        w, e = be.equalise_signal(g["rx"][:, s0:s0 + 1024], 2, 5e-3, 4, Ntaps=17, Niter=10, method="cma",
                                  adaptive_stepsize=True)
        assert np.array_equal(w, taps[k]) and np.array_equal(e, errs[k])
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
