"""Host-side logic of the product (no GPU): constant tables, segmentation plan, argument
validation, error mapping, patch wiring."""
import os

import numpy as np
import pytest

from qampy_b200 import _lib, pipeline, theory


def test_theory_constants_match_reference(golden):
    g = golden("g0_constants")
    for M in (4, 16, 32, 64, 128, 256):
        assert np.array_equal(theory.cal_symbols_qam(M), g["syms_%d" % M])
        assert theory.cal_scaling_factor_qam(M) == pytest.approx(float(g["scale_%d" % M]), rel=1e-15)
        for m in ("cma", "cma2", "sgncma", "mcma", "rde", "mrde", "sbd", "mddma", "dd"):
            mine = theory.generate_symbols_for_eq(m, M, np.complex128)
            ref = g["eqsyms_%s_%d" % (m, M)]
            assert mine.shape == ref.shape
            np.testing.assert_allclose(mine, ref, rtol=1e-14, atol=1e-15)
            assert np.array_equal(mine.astype(np.complex64), ref.astype(np.complex64))
    for L, nt, os_ in ((20000, 11, 2), (2000000, 21, 2), (20000000, 45, 2), (6000, 11, 2), (5000, 7, 1)):
        assert theory.cal_training_symbol_len(os_, nt, L) == int(g["trsyms_%d_%d_%d" % (L, nt, os_)])


def test_method_constants_survey_values():
    # SURVEY.md section 8d "Method constants (probed, c64)"
    assert theory.generate_symbols_for_eq("cma", 64, np.complex64)[0, 0].real == pytest.approx(1.3809524, rel=1e-6)
    assert theory.generate_symbols_for_eq("mcma", 16, np.complex64)[0, 0] == pytest.approx(0.82 + 0.82j, rel=1e-6)
    mr = theory.generate_symbols_for_eq("mrde", 64, np.complex64)[0]
    np.testing.assert_allclose(mr[:4].real, [0.02380952, 0.21428572, 0.5952381, 1.1666666], rtol=1e-6)
    np.testing.assert_allclose(mr[4:].real, [0.11904762, 0.4047619, 0.88095236], rtol=1e-6)
    ang = theory.bps_test_angles(64, np.float32)
    assert ang.shape == (1, 64) and ang.dtype == np.float32 and ang[0, 0] == np.float32(-np.pi / 4)


def test_reshape_symbols_rules():
    # qampy/core/equalisation/equalisation.py:568-594 / reference test_equalisation.py:195-291
    s = theory.reshape_symbols(None, "mcma", 16, np.complex64, 2)
    assert s.shape == (2, 1) and s.dtype == np.complex64
    coded = theory.normalised_symbols(16)
    s = theory.reshape_symbols(coded, "sbd", 16, np.complex128, 3)
    assert s.shape == (3, 16) and np.array_equal(s[0], s[2])
    s = theory.reshape_symbols(coded, "cma", 16, np.complex64, 2)      # non-decision: user symbols ignored
    assert s.shape == (2, 1)
    with pytest.raises(ValueError):
        theory.reshape_symbols(np.zeros((3, 16), complex), "sbd", 16, np.complex64, 2)
    with pytest.raises(ValueError):
        theory.generate_symbols_for_eq("sbd_data", 16, np.complex64)
    with pytest.raises(ValueError):
        theory.generate_symbols_for_eq("nope", 16, np.complex64)


def test_plan_segments_tiles_the_capture():
    cfg = pipeline.ReceiverConfig(ntaps=45, os=2, seg_symbols=8192)
    L = 2 * 10 ** 7
    groups = pipeline.plan_segments(L, cfg)
    N = (L - 45 + 1) // 2
    assert groups[0] == (0, 8192, N // 8192, 0)
    first, nsym, nseg, drop = groups[1]
    assert nseg == 1 and nsym == 8192 and first + nsym == N and nsym - drop == N % 8192
    # every output symbol is produced exactly once after dropping the overlap
    assert groups[0][1] * groups[0][2] + (nsym - drop) == N
    # last sample read stays inside the capture
    assert (first + nsym - 1) * 2 + 45 <= L
    assert pipeline.plan_segments(L, pipeline.ReceiverConfig(ntaps=45, seg_symbols=None)) == [(0, N, 1, 0)]
    assert pipeline.plan_segments(2 * 8192 * 3 + 44, cfg) == [(0, 8192, 3, 0)]
    with pytest.raises(ValueError):
        pipeline.plan_segments(30, cfg)


def test_shard_segments_partition():
    for nseg in (1, 7, 1220, 1221):
        for world in (1, 2, 4, 8):
            spans = [pipeline.shard_segments(nseg, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nseg
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_argument_validation_before_any_cuda_work():
    import qampy_b200.pythran_dsp as dsp
    import qampy_b200.pythran_equalisation as pe
    E = np.zeros((2, 100), np.complex64)
    w = theory.init_taps(5, 2, np.complex64)
    sy = theory.reshape_symbols(None, "mcma", 4, np.complex64, 2)
    with pytest.raises(ValueError, match="Unknown method"):
        pe.train_equaliser(E, 10, 1, 2, 1e-3, w, np.arange(2), False, sy, "nope")
    with pytest.raises(AssertionError):
        pe.apply_filter_to_signal(E, 0, w)
    with pytest.raises(TypeError):
        pe.apply_filter_to_signal(E.real.copy(), 2, w)               # real E needs real wx (:33-36)
    with pytest.raises(ValueError, match="Unknown method"):
        pe.train_equaliser_realvalued(E.real.copy(), 10, 1, 2, 1e-3, w.real.copy(), np.arange(2), False,
                                      sy.real.copy(), "mcma")
    with pytest.raises(ValueError, match="p must be"):
        dsp.bps(E[0], np.zeros((7, 8), np.float32), sy[0], 4)        # angle table: one row or one per symbol
    with pytest.raises(TypeError):
        pe.apply_filter_to_signal(np.zeros((2, 10), np.int32), 2, w)


def test_status_codes_map_to_reference_exceptions():
    lib = _lib.load()
    assert _lib.check(0) == 0 and _lib.check(3) == 3
    assert lib.qb_method_from_name(b"mrde") == _lib.METHODS["mrde"] == 5
    for name, code in _lib.METHODS.items():
        assert lib.qb_method_from_name(name.encode()) == code
    with pytest.raises(ValueError, match="Unknown method"):
        _lib.check(lib.qb_method_from_name(b"nope"))
    with pytest.raises(NotImplementedError):
        _lib.check(-2)
    with pytest.raises(MemoryError):
        _lib.check(-4)
    with pytest.raises(_lib.QampyB200Error):
        _lib.check(-3)


@pytest.mark.skipif(not os.path.isdir("/root/reference/qampy"), reason="reference checkout not present")
def test_patch_wires_both_seams_and_restores():
    import sys
    sys.path.insert(0, "/root/reference")
    import warnings
    warnings.filterwarnings("ignore")
    import qampy.core.equalisation as ceq_pkg
    import qampy.core.equalisation.equalisation as ceq
    import qampy.core.phaserecovery as cph
    from qampy.core.equalisation import pythran_equalisation as ref_pe
    from qampy_b200 import patch
    import qampy_b200.pythran_equalisation as q_pe
    orig = (ceq.equalise_signal, ceq_pkg.dual_mode_equalisation, cph.bps, ref_pe.train_equaliser, cph._bps_idx_pyt)
    names = patch.patch("l2")
    assert {"equalise_signal", "dual_mode_equalisation", "apply_filter", "bps"} <= set(names)
    assert ceq.equalise_signal is not orig[0] and ceq_pkg.dual_mode_equalisation is not orig[1]
    assert cph.bps is not orig[2] and ref_pe.train_equaliser is orig[3]
    patch.patch("l1")                       # re-patching first restores
    assert ceq.equalise_signal is orig[0] and cph.bps is orig[2]
    assert ref_pe.train_equaliser is q_pe.train_equaliser and cph._bps_idx_pyt is not orig[4]
    patch.unpatch()
    assert (ceq.equalise_signal, ceq_pkg.dual_mode_equalisation, cph.bps, ref_pe.train_equaliser,
            cph._bps_idx_pyt) == orig
    with patch.patched("l2"):
        assert cph.bps is not orig[2]
    assert cph.bps is orig[2]
    with pytest.raises(ValueError):
        patch.patch("l3")


def test_synth_capture_is_blockwise_synth_signal():
    """Long captures (BASELINE config C5) are synthesised block by block: block b == synth_signal(seed + b)."""
    import torch
    from qampy_b200 import synth
    E, s0 = synth.synth_capture(16, 2500, block=1000, seed=7, snr_db=25.0)
    assert E.shape == (2, 5000) and E.dtype == torch.complex64 and s0.shape == (2, 1000)
    for b, (a, n) in enumerate(((0, 1000), (1000, 1000), (2000, 500))):
        Eb, sb = synth.synth_signal(16, n, seed=7 + b, snr_db=25.0)
        assert torch.equal(E[:, 2 * a:2 * (a + n)], Eb)
        if b == 0:
            assert torch.equal(s0, sb)


def test_unique_alphabet_keeps_first_occurrences_and_decisions():
    """A searched alphabet (sbd / mddma / dd) without its repeats gives the same det_symbol decisions
    (pythran_equalisation.py:240-265: first STRICT minimum, the value is returned); other methods are untouched."""
    from qampy_b200 import theory
    rng = np.random.default_rng(3)
    q = theory.normalised_symbols(4).astype(np.complex64)
    seq = np.stack([q[rng.integers(0, 4, 1024)], q[rng.integers(0, 2, 1024)]])      # row 1 uses 2 points only
    u = theory.unique_alphabet(seq, "sbd")
    assert u.shape == (2, 4)
    for r in range(2):
        _, first = np.unique(seq[r], return_index=True)
        k = first.size
        assert np.array_equal(u[r, :k], seq[r][np.sort(first)]) and np.all(u[r, k:] == u[r, 0])

    def det(x, sy):
        d0, s = 1000., 1 + 0j
        for v in sy:
            d = abs(x - v) ** 2
            if d < d0:
                d0, s = d, v
        return s
    for _ in range(300):
        x = complex(rng.normal(), rng.normal())
        for r in range(2):
            assert det(x, seq[r]) == det(x, u[r])
    assert theory.unique_alphabet(seq, "sbd_data") is seq and theory.unique_alphabet(seq, "mcma") is seq
    a16 = np.tile(theory.normalised_symbols(16).astype(np.complex64), (2, 1))
    assert theory.unique_alphabet(a16, "dd") is a16


def test_host_chunks_tile_the_segments_and_taper_the_last_run():
    """The overlapped host path cuts the main group into runs of whole segments (first run tapered 1/4 + 1/4 + 1/2,
    last run 1/2 + 1/4 + 1/4, the end-aligned extra segment last): every segment exactly once, in order."""
    from qampy_b200 import pipeline
    for nseg, extra in ((1182, True), (1184, False), (5, True), (1, False), (40, False)):
        groups = [(0, 100, nseg, 0)] + ([(nseg * 100 - 37, 100, 1, 63)] if extra else [])
        for nchunks in (1, 3, 6, 8, 64):
            for taper in (True, False):
                runs = list(pipeline._host_chunks(groups, nchunks, taper))
                main = runs[:-1] if extra else runs
                assert [r[4] for r in main] == list(np.cumsum([0] + [r[2] for r in main])[:-1])
                assert sum(r[2] for r in main) == nseg and all(r[2] > 0 and r[1] == 100 and r[3] == 0 for r in main)
                assert all(r[0] == r[4] * 100 for r in main)
                if extra:
                    assert runs[-1] == (nseg * 100 - 37, 100, 1, 63, nseg)
                if taper and nchunks > 1 and nseg // min(nchunks, nseg) >= 8:
                    assert len(main) == min(nchunks, nseg) + 4 and main[-1][2] <= main[3][2] // 3 + 1
                    assert main[0][2] <= main[3][2] // 3 + 1 and main[1][2] <= main[3][2] // 3 + 1
                else:
                    assert len(main) == min(nchunks, nseg)


def test_plan_segments_with_a_phase_search_halo():
    """bps_halo: the segments tile [H, N - H); every segment plus H symbols on either side stays inside the capture;
    a single segment has no halo (reference call on the whole capture)."""
    H, S, ntaps = 45, 8192, 45
    cfg = pipeline.ReceiverConfig(ntaps=ntaps, os=2, seg_symbols=S, bps_halo=H)
    for L in (2 * 10 ** 6, 2 * (3 * S + 2 * H) + ntaps - 1, 2 * 40000 + 7):
        N = (L - ntaps + 1) // 2
        groups = pipeline.plan_segments(L, cfg)
        assert groups[0][0] == H and groups[0][1] == S and groups[0][2] == (N - 2 * H) // S
        covered = groups[0][1] * groups[0][2]
        for first, nsym, nseg, drop in groups:
            assert first - H >= 0 and ((first + nseg * nsym + H) - 1) * 2 + ntaps <= L
        if len(groups) > 1:
            first, nsym, nseg, drop = groups[1]
            assert nseg == 1 and first + nsym == N - H
            covered += nsym - drop
        assert covered == N - 2 * H
    assert pipeline.plan_segments(2 * 5000, cfg) == [(0, (2 * 5000 - ntaps + 1) // 2, 1, 0)]


def test_rank_capture_ranges_with_halo_reproduce_the_single_process_plan():
    """Every rank plans ITS sample range with the same halo and produces exactly the symbols it owns; the ranks
    together produce [H, N - H) once."""
    ntaps, S, H, nsym = 21, 1024, 21, 9000
    cfg = pipeline.ReceiverConfig(ntaps=ntaps, os=2, seg_symbols=S, bps_halo=H)
    N = (nsym * 2 - ntaps + 1) // 2
    for world in (1, 2, 3):
        nxt = H
        for rank in range(world):
            a, b, s0, s1 = pipeline.rank_capture_range(nsym, ntaps, 2, S, rank, world, halo=H)
            assert s0 == nxt and a == (s0 - H) * 2 and b <= nsym * 2
            nxt = s1
            got = []
            for first, n, nseg, drop in pipeline.plan_segments(b - a, cfg):
                for s in range(nseg):
                    lo = a // 2 + first + s * n
                    got += list(range(lo + (drop if nseg == 1 else 0), lo + n))
            assert got == list(range(s0, s1))
        assert nxt == N - H


def test_ser_segments_counts_errors_and_resolves_row_rotation_delay():
    """bench.py's sanity gate: per (segment, row) the sent row, the pi/2 rotation and the delay are found on a probe."""
    import torch
    from qampy_b200 import synth
    rng = np.random.default_rng(3)
    M, nsym, S, nseg = 64, 9000, 2000, 4
    al = theory.normalised_symbols(M).astype(np.complex64)
    syms = al[rng.integers(0, M, (2, nsym))]
    firsts = 30 + np.arange(nseg) * S
    out = np.empty((nseg, 2, S), np.complex64)
    for s in range(nseg):
        d = 7 + s
        out[s, 0] = syms[s % 2, firsts[s] + d: firsts[s] + d + S] * (1j ** s)          # rows swapped, rotated, delayed
        out[s, 1] = syms[1 - s % 2, firsts[s] + d: firsts[s] + d + S] * (1j ** (s + 1))
    out += 0.01 * (rng.standard_normal(out.shape) + 1j * rng.standard_normal(out.shape)).astype(np.complex64)
    out[2, 1, 700] += al[0] - al[63]                                                     # three certain symbol errors
    out[2, 1, 701] += al[0] - al[63]
    out[3, 0, 1999] += al[63] - al[0]
    e, c = synth.ser_segments(torch.from_numpy(out), torch.from_numpy(syms), M, firsts, seg_chunk=3)
    want = np.zeros((nseg, 2), np.int64)
    want[2, 1], want[3, 0] = 2, 1
    assert np.array_equal(e.numpy(), want) and int(c.min()) == S


def test_rrc_tap_vector_reproduces_the_reference_pulse_shaping():
    """synth_device._pulse_taps (own restatement of the root-raised-cosine impulse response with its two removable
    singularities, normalised like core/filter.py:201-205) in a NumPy model of the device pipeline == the reference's
    rrcos_resample output on the golden symbols (tests/golden/g13_synth.npz)."""
    import os
    from qampy_b200.synth_device import _pulse_taps
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g13_synth.npz"))
    for tag in "ab":
        M, n, beta, taps, theta, dgd, fb = g[tag + "_par"]
        n, taps = int(n), int(taps)
        h = _pulse_taps(taps, 2 * fb, 1 / fb, beta)
        assert h.max() == 1.0 and np.allclose(h, h[::-1])
        up = np.zeros((2, 2 * n), complex)
        up[:, ::2] = g[tag + "_symbols"]
        nfft = 1 << int(np.ceil(np.log2(2 * n + taps - 1)))
        y = np.fft.ifft(np.fft.fft(up, nfft, axis=1) * np.fft.fft(h, nfft)[None], axis=1)
        y = y[:, (taps - 1) // 2:(taps - 1) // 2 + 2 * n]
        assert np.sqrt(np.mean(np.abs(y - g[tag + "_plain"]) ** 2)) < 1e-12
