"""Golden fixtures for the rows SURVEY.md section 8f marks "next", generated like make_golden.py by
EXECUTING THE REFERENCE (interpreted Pythran sources, asserts stripped):

    python tests/golden/make_golden_next.py

G7  two-stage blind phase search (qampy/core/phaserecovery.py:222-288): the per-symbol test-angle table
    form of pythran_dsp.bps (p == L) and select_angles, c64 and c128.
G8  real-valued equaliser methods (equalisation.py:529-565 -> pythran_equalisation.py:80-128):
    cma_real, sgncma_real, dd_real, dd_data_real through equalise_signal(apply=True), fixed and adaptive
    step size, c64 and c128.
G9  pilot-based receiver on arrays (qampy/core/pilotbased_receiver.py: frame_sync :329-434,
    equalize_pilot_sequence :454-554, pilot_based_cpe_new :258-327; phaserecovery.find_freq_offset /
    comp_freq_offset :385-473), the recipe of test/test_equalisation.py:150-164 at reduced size:
    dual-pol 16-QAM, frame 2**13, pilot sequence 512, one phase pilot per 32, 3 frames, SNR 25 dB,
    DGD 10 ps, 100 MHz offset, 100 kHz linewidth, modal delay 700 samples.
G10 decisions and quality metrics: make_decision (pythran_equalisation.py:306-334), soft_l_value_demapper and
    soft_l_value_demapper_minmax (pythran_dsp.py:95-131), estimate_snr (:244-286) on noisy 16/64-QAM, c64 and c128.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    load_reference()
    from qampy import signals, impairments
    from qampy.core import phaserecovery as cph
    from qampy.core import pythran_dsp as pd

    out = {}
    for tag, dt, M, nsym, A, N, B, seed in (("c64", np.complex64, 16, 1200, 16, 8, 4, 71),
                                            ("c128", np.complex128, 64, 900, 32, 10, 6, 72)):
        sg = signals.SignalQAMGrayCoded(M, nsym, nmodes=2, fb=40e9, dtype=dt, seed=[seed, seed + 1])
        np.random.seed(seed)
        s2 = impairments.apply_phase_noise(impairments.change_snr(sg, 22), 300e3)
        E = np.asarray(s2)
        coded = np.asarray(sg.coded_symbols)
        En, ph = cph.bps_twostage(E, A, coded, N, B=B)
        # the L1 calls of the second stage for mode 0, step by step (phaserecovery.py:271-281)
        rt = E.real.dtype
        angles = np.linspace(-np.pi / 4, np.pi / 4, A, endpoint=False, dtype=rt).reshape(1, -1)
        idx1 = pd.bps(np.copy(E[0]), angles, coded, N)
        ph1 = pd.select_angles(np.copy(angles), idx1)
        b = np.linspace(-B / 2, B / 2, B)
        phn = (ph1[:, np.newaxis] + b[np.newaxis, :] / (B * A) * np.pi / 2).astype(rt)
        idx2 = pd.bps(np.copy(E[0]), phn, coded, N)
        phf = pd.select_angles(np.copy(phn), idx2)
        out.update({"in_" + tag: E, "coded_" + tag: coded, "A_" + tag: A, "N_" + tag: N, "B_" + tag: B,
                    "out_" + tag: np.asarray(En), "ph_" + tag: ph, "idx1_" + tag: idx1, "phn_" + tag: phn,
                    "idx2_" + tag: idx2, "phf_" + tag: phf})
    np.savez_compressed(os.path.join(HERE, "g7_bps_twostage.npz"), **out)

    # ---- G8: real-valued equaliser methods ---------------------------------------------------------------
    from make_golden import make_signal
    from qampy.core.equalisation import equalisation as ceq
    out = {}
    for tag, dt, M, seed in (("c64", np.complex64, 4, 81), ("c128", np.complex128, 16, 82)):
        sig, s = make_signal(M, 700, 2, dt, seed, 22, theta=np.pi / 5.5, dgd=20e-12)
        E = np.asarray(s)
        out["E_" + tag] = E
        out["M_" + tag] = M
        out["tx_" + tag] = np.asarray(sig)
        for method, adaptive in (("cma_real", False), ("sgncma_real", False), ("dd_real", False),
                                 ("dd_data_real", False), ("cma_real", True), ("dd_real", True)):
            kw = {}
            if method == "dd_data_real":
                kw["symbols"] = np.asarray(sig)
            Eo, wxy, err = ceq.equalise_signal(E, 2, 2e-3, M, Ntaps=7, method=method, apply=True,
                                               adaptive_stepsize=adaptive, **kw)
            key = "%s_%s%s" % (tag, method, "_ad" if adaptive else "")
            out["out_" + key], out["wxy_" + key], out["err_" + key] = np.asarray(Eo), wxy, err
    np.savez_compressed(os.path.join(HERE, "g8_real_valued.npz"), **out)

    # ---- G9: pilot-based receiver, array level ---------------------------------------------------------------
    from qampy.core import pilotbased_receiver as pr
    from qampy.core.equalisation import equalisation as ceq2
    np.random.seed(91)
    fl, sl, rat, osf = 2 ** 13, 512, 32, 2
    sig = signals.SignalWithPilots(16, fl, sl, rat, nframes=3, nmodes=2, Mpilots=4, fb=24e9, dtype=np.complex64,
                                   seed=[91, 92])      # bit sources seeded: the file regenerates identically
    s2 = sig.resample(2 * sig.fb, beta=0.01)
    s3 = impairments.simulate_transmission(s2, snr=25, dgd=10e-12, freq_off=100e6, lwdth=100e3,
                                           modal_delay=[700, 700])
    rx = np.array(s3)
    pilot_seq, ph_pilots = np.array(sig.pilot_seq), np.array(sig.ph_pilots)
    idx_pil = np.array(sig._idx_pil)
    out = dict(rx=rx, pilot_seq=pilot_seq, ph_pilots=ph_pilots, idx_pil=idx_pil, frame_len=fl, os=osf,
               symbols_tx=np.array(sig.symbols), coded=np.array(sig.coded_symbols), M=16)
    # sync2frame (signals.py:1709-1741) on the array
    sf, foe, order, wx1, ok = pr.frame_sync(rx, pilot_seq, osf, frame_len=fl, M_pilot=4, mu=5e-3, Ntaps=17,
                                            adaptive_stepsize=True, Niter=10, method="cma")
    out.update(fs_shift=sf.copy(), fs_foe=foe, fs_order=order, fs_wx1=wx1, fs_ok=ok)
    rx2 = rx[order, :]
    sf[sf < 0] += fl * osf
    shiftf = sf[order]
    foe_off = np.ones(foe.shape) * np.mean(foe)
    rx3 = cph.comp_freq_offset(rx2, foe_off, osf)                       # corr_foe (signals.py:1744-1747)
    out.update(shiftfctrs=shiftf, rx_synced_head=rx3[:, :4096])
    # pilot_equaliser (qampy/equalisation.py:307-330) for frame 0, then apply to frame 0
    Ntaps = 45
    eq_shift = shiftf - (Ntaps - 17) // 2
    for tag, methods in (("sbd", ("cma", "sbd")), ("data", ("cma", "sbd_data"))):
        taps, foe_all = pr.equalize_pilot_sequence(rx3, pilot_seq, eq_shift, os=osf, mu=(1e-3, 1e-3), foe_comp=False,
                                                   Ntaps=Ntaps, methods=methods)
        out.update({"taps_" + tag: taps, "foe_all_" + tag: foe_all})
    taps = out["taps_sbd"]
    i0 = eq_shift[0]
    eq = ceq2.apply_filter(rx3[:, i0:i0 + fl * osf + Ntaps - 1], osf, taps)
    out.update(eq_frame0=np.asarray(eq))
    idx = np.nonzero(idx_pil)[0][sl:]
    cpe, trace = pr.pilot_based_cpe_new(eq, ph_pilots, idx, fl, seq_len=None, max_num_blocks=None,
                                        use_pilot_ratio=1, num_average=5, nframes=1)
    out.update(cpe_out=np.asarray(cpe), cpe_trace=trace)
    foe_p, foe_pm, cond = pr.pilot_based_foe(np.asarray(eq)[:, :sl], pilot_seq)
    out.update(pfoe=foe_p, pfoe_mode=foe_pm, pfoe_cond=cond)
    np.savez_compressed(os.path.join(HERE, "g9_pilot_rx.npz"), **out)
    # ---- G10: decisions and metrics ---------------------------------------------------------------------------
    from qampy.core.equalisation import pythran_equalisation as pe2
    out = {}
    for tag, dt, M, n, snr_db, seed in (("c64", np.complex64, 64, 3000, 11, 101), ("c128", np.complex128, 16, 2000, 12, 102)):
        sg = signals.SignalQAMGrayCoded(M, n, nmodes=1, fb=40e9, dtype=dt, seed=[seed])
        np.random.seed(seed)
        rx = np.asarray(impairments.change_snr(sg, snr_db))[0]
        coded = np.asarray(sg.coded_symbols)
        det, dist, idx = pe2.make_decision(rx, coded)
        bm = np.asarray(sg._bitmap_mtx)
        snr_lin = dt(0).real.dtype.type(10 ** (snr_db / 10))
        lv = pd.soft_l_value_demapper(rx, sg.Nbits, snr_lin, bm)
        lvm = pd.soft_l_value_demapper_minmax(rx, sg.Nbits, snr_lin, bm)
        est = pd.estimate_snr(rx, np.asarray(sg)[0], coded)
        out.update({"rx_" + tag: rx, "tx_" + tag: np.asarray(sg)[0], "coded_" + tag: coded, "det_" + tag: det,
                    "dist_" + tag: dist, "idx_" + tag: idx, "bitmap_" + tag: bm, "nbits_" + tag: sg.Nbits,
                    "snr_" + tag: snr_lin, "lv_" + tag: lv, "lvmm_" + tag: lvm, "est_" + tag: np.array(est, dtype=np.float64)})
    np.savez_compressed(os.path.join(HERE, "g10_decisions.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
