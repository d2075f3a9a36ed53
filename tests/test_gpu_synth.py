"""On-device signal synthesis (qampy_b200/synth_device.py + csrc/synth_ops.cu) against the reference's
generators: deterministic stages on golden vectors made by the reference (tests/golden/make_golden_synth.py), random
stages on their statistics (the reference draws from np.random; the kernels from their own counter-based generator)."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g13_synth.npz")


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2)))


@pytest.fixture(scope="module")
def env():
    import torch
    from qampy_b200 import _lib, synth_device
    _lib.require_device()
    return torch, synth_device, torch.device("cuda", 0)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pulse_shaping_and_pmd_equal_the_reference(env, tag):
    """rrcos_resample (core/resample.py:73-126; per mode, with and without renormalise) and apply_PMD_to_field
    (core/impairments.py:94-131) on the reference's own input: <= 1e-9 rms in complex128 (the bound asked for is 1e-6)."""
    torch, sd, dev = env
    g = np.load(GOLD)
    M, n, beta, taps, theta, dgd, fb = g[tag + "_par"]
    syms = torch.from_numpy(g[tag + "_symbols"]).to(dev)
    plain = sd.rrcos_resample(syms, fb, 2 * fb, beta=float(beta), taps=int(taps), renormalise=False)
    shaped = sd.rrcos_resample(syms, fb, 2 * fb, beta=float(beta), taps=int(taps), renormalise=True)
    assert plain.shape == shaped.shape == g[tag + "_shaped"].shape and shaped.dtype == torch.complex128
    assert rms(plain.cpu().numpy() - g[tag + "_plain"]) < 1e-9
    assert rms(shaped.cpu().numpy() - g[tag + "_shaped"]) < 1e-9
    one = sd.rrcos_resample(syms[1], fb, 2 * fb, beta=float(beta), taps=int(taps), renormalise=True)     # 1-D like the reference
    assert one.dim() == 1 and rms(one.cpu().numpy() - g[tag + "_shaped"][1]) < 1e-9
    pmd = sd.apply_PMD_to_field(torch.from_numpy(g[tag + "_shaped"]).to(dev), float(theta), float(dgd), 2 * fb)
    assert rms(pmd.cpu().numpy() - g[tag + "_pmd"]) < 1e-9
    # unitary: PMD moves power between the polarisations but keeps the total
    assert abs(float((pmd.abs() ** 2).sum()) - float(np.sum(np.abs(g[tag + "_shaped"]) ** 2))) < 1e-6
    # complex64 field in -> complex64 out
    p32 = sd.apply_PMD_to_field(torch.from_numpy(g[tag + "_shaped"].astype(np.complex64)).to(dev), float(theta),
                                float(dgd), 2 * fb)
    assert p32.dtype == torch.complex64 and rms(p32.cpu().numpy() - g[tag + "_pmd"]) < 2e-6


def test_awgn_has_the_reference_statistics(env):
    """add_awgn / change_snr (core/impairments.py:188-233): sig + strgth (N + iN)/sqrt(2): per-component variance
    strgth^2/2, zero mean, Gaussian, white, independent rows; repeatable from the seed; block-wise == whole."""
    torch, sd, dev = env
    n = 1 << 21
    sig = torch.zeros((2, n), dtype=torch.complex128, device=dev)
    sig[0] += 0.5 - 0.25j
    out = sd.add_awgn(sig, 0.3, seed=7)
    noise = (out - sig).cpu().numpy()
    for r in range(2):
        for comp in (noise[r].real, noise[r].imag):
            assert abs(comp.mean()) < 5 * 0.3 / math.sqrt(2 * n)
            assert abs(comp.var() / (0.3 ** 2 / 2) - 1) < 0.01
            z = comp / comp.std()
            assert abs(np.mean(z ** 3)) < 0.02 and abs(np.mean(z ** 4) - 3) < 0.05            # Gaussian
            assert abs(np.mean(z[1:] * z[:-1])) < 0.005                                        # white
        assert abs(np.mean(noise[r].real * noise[r].imag)) / (0.3 ** 2 / 2) < 0.005            # I and Q independent
    assert abs(np.mean(noise[0].real * noise[1].real)) / (0.3 ** 2 / 2) < 0.005                # rows independent
    assert torch.equal(out, sd.add_awgn(sig, 0.3, seed=7)) and not torch.equal(out, sd.add_awgn(sig, 0.3, seed=8))
    # a capture made in two blocks is the same capture (counter-based deviates)
    h = n // 2
    a = sd._tail(sig[:, :h], 0.3, 0.0, 7, torch.complex128)
    b = sd._tail(sig[:, h:], 0.3, 0.0, 7, torch.complex128, index0=h)
    assert torch.equal(torch.cat([a, b], dim=1), out)
    # change_snr: noise power per sample = P os 10^(-snr/10) with P the mean power of the whole array
    s2 = torch.exp(1j * torch.linspace(0, 50, n, device=dev, dtype=torch.float64)).to(torch.complex64)[None].repeat(2, 1)
    o2 = sd.change_snr(s2, 20.0, 40e9, 80e9, seed=3)
    assert o2.dtype == torch.complex64
    pn = float(((o2 - s2).abs() ** 2).mean())
    assert abs(pn / (1.0 * 2 * 10 ** (-2.0)) - 1) < 0.01


def test_phase_noise_is_a_wiener_walk(env):
    """phase_noise / apply_phase_noise (core/impairments.py:133-186): steps N(0, 2 pi df / fs), cumulative sum,
    signal * exp(i phase)."""
    torch, sd, dev = env
    n, df, fs = 1 << 20, 100e3, 80e9
    ph = sd.phase_noise((2, n), df, fs, seed=5, device=dev)
    assert ph.shape == (2, n) and ph.dtype == torch.float64
    steps = torch.diff(ph, dim=1, prepend=torch.zeros((2, 1), dtype=torch.float64, device=dev)).cpu().numpy()
    var = 2 * math.pi * df / fs
    for r in range(2):
        assert abs(steps[r].var() / var - 1) < 0.01 and abs(steps[r].mean()) < 5 * math.sqrt(var / n)
        z = steps[r] / steps[r].std()
        assert abs(np.mean(z[1:] * z[:-1])) < 0.005 and abs(np.mean(z ** 4) - 3) < 0.05
    assert abs(np.mean(steps[0] * steps[1])) / var < 0.005
    sig = (torch.randn((2, n), dtype=torch.float64, device=dev) + 0.3j).to(torch.complex64)
    out = sd.apply_phase_noise(sig, df, fs, seed=5)
    want = sig.to(torch.complex128) * torch.exp(1j * ph)
    assert out.dtype == torch.complex64 and rms((out.to(torch.complex128) - want).cpu().numpy()) < 2e-7
    # scan == plain cumulative sum of the same steps (to rounding), at tile and block boundaries too
    assert float((ph - torch.cumsum(torch.from_numpy(steps).to(dev), dim=1)).abs().max()) < 1e-10
    # block-wise with the carried phase
    h = n // 2 + 2048 * 3
    a, pa = sd._tail(sig[:, :h], None, math.sqrt(var), 5, torch.complex64, want_phase=True)
    b, pb = sd._tail(sig[:, h:], None, math.sqrt(var), 5, torch.complex64, index0=h, phase0=pa[:, -1].cpu().numpy(),
                     want_phase=True)
    assert float((torch.cat([pa, pb], dim=1) - ph).abs().max()) < 1e-10


def test_synthesised_signal_is_what_the_receiver_expects(env):
    """The composed generator feeds the chain: unit power, and the equaliser + BPS recover the symbols."""
    torch, sd, dev = env
    from qampy_b200 import equalisation, phaserecovery, synth, theory
    E, syms = sd.synth_signal(16, 60000, snr_db=24.0, seed=11, device=dev)
    assert E.dtype == torch.complex64 and E.shape == (2, 120000)
    assert abs(float((E.abs() ** 2).mean()) - (1 + 2 * 10 ** (-2.4))) < 0.02
    Eo, w, _ = equalisation.dual_mode_equalisation(E.cpu().numpy(), 2, (2e-3, 2e-3), 16, Ntaps=21, methods=("mcma", "mrde"))
    Eb, ph = phaserecovery.bps(Eo, 32, theory.normalised_symbols(16).astype(np.complex64), 21)
    assert synth.ser(Eb[:, 30000:50000], syms.cpu().numpy()[:, 30000:50300], 16) < 1e-3
