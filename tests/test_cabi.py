"""The C-ABI library loads and exports every symbol include/qampy_b200.h declares; the host-only
helpers work; compute entry points fail LOUDLY (no CPU fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

from qampy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "qampy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libqampy_b200.so does not export %s" % n
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert _lib.load().qb_version() == 100


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_detect_grid_host_helper():
    from qampy_b200 import theory
    lib = _lib.load()
    for M, expect in ((4, 2), (16, 4), (64, 8), (256, 16), (32, 0), (128, 0)):
        for dt, code, rt in ((np.complex64, 0, np.float32), (np.complex128, 1, np.float64)):
            sy = np.ascontiguousarray(theory.normalised_symbols(M).astype(dt))
            rng = np.random.default_rng(M)
            sy = sy[rng.permutation(M)]                     # order must not matter (coded_symbols is Gray ordered)
            lre, lim = np.zeros(64, rt), np.zeros(64, rt)
            nre, nim = ctypes.c_int64(0), ctypes.c_int64(0)
            rc = lib.qb_detect_grid_host(code, sy.ctypes.data, M, lre.ctypes.data, ctypes.byref(nre),
                                         lim.ctypes.data, ctypes.byref(nim))
            assert rc == (1 if expect else 0)
            if expect:
                assert nre.value == nim.value == expect
                assert np.array_equal(lre[:expect], np.unique(sy.real))
                assert np.array_equal(lim[:expect], np.unique(sy.imag))
    # a grid with a missing point or a non-uniform axis is not a grid
    sy = np.ascontiguousarray(theory.normalised_symbols(16).astype(np.complex64)[:-1])
    lre, lim = np.zeros(64, np.float32), np.zeros(64, np.float32)
    nre, nim = ctypes.c_int64(0), ctypes.c_int64(0)
    assert lib.qb_detect_grid_host(0, sy.ctypes.data, 15, lre.ctypes.data, ctypes.byref(nre), lim.ctypes.data,
                                   ctypes.byref(nim)) == 0
    sy = np.array([a + 1j * b for a in (-1.0, 0.1, 1.0) for b in (-1.0, 0.0, 1.0)], np.complex64)
    assert lib.qb_detect_grid_host(0, sy.ctypes.data, 9, lre.ctypes.data, ctypes.byref(nre), lim.ctypes.data,
                                   ctypes.byref(nim)) == 0


def test_set_option_accepts_known_names_only():
    """qb_set_option: the kernel-selection overrides of the tests (the environment is read once, never per launch)."""
    lib = _lib.load()
    for name in (b"TRAIN_KERNEL", b"TRAIN_LPS", b"TRAIN_GLA", b"LA_TILE", b"BPS_KERNEL", b"BPS_SPLIT"):
        assert lib.qb_set_option(name, b"1") == 0
        assert lib.qb_set_option(name, None) == 0
    assert lib.qb_set_option(b"NO_SUCH_OPTION", b"1") == -1 and b"unknown option" in lib.qb_last_error()
    assert lib.qb_set_option(None, None) == -1


def test_invalid_arguments_are_rejected_with_messages():
    lib = _lib.load()
    z = ctypes.c_void_p(0)
    modes = np.array([0, 5], np.int64)
    buf = np.zeros(64, np.complex64)
    rc = lib.qb_apply_filter_to_signal_host(0, buf.ctypes.data, 2, 16, 2, buf.ctypes.data, 3, modes.ctypes.data, 2,
                                            buf.ctypes.data)
    assert rc == -1 and b"mode" in lib.qb_last_error()
    rc = lib.qb_apply_filter_to_signal_host(7, buf.ctypes.data, 2, 16, 2, buf.ctypes.data, 3, modes.ctypes.data, 1,
                                            buf.ctypes.data)
    assert rc == -1 and b"dtype" in lib.qb_last_error()
    rc = lib.qb_bps_host(0, buf.ctypes.data, 1, 16, buf.ctypes.data, z, 300, buf.ctypes.data, 4, 2, z, z, z)
    assert rc == -1 and b"test angles" in lib.qb_last_error()
    rc = lib.qb_train_equaliser_host(0, buf.ctypes.data, 2, 16, 2, 1, 2, buf.ctypes.data, buf.ctypes.data, 3,
                                     modes.ctypes.data, 1, 0, buf.ctypes.data, 1, 42, 0, z)
    assert rc == -1 and b"Unknown method" in lib.qb_last_error()


def test_no_cpu_fallback_without_a_device():
    """On a machine without a usable GPU every compute entry point must raise, not compute."""
    lib = _lib.load()
    if lib.qb_device_count() >= 1:
        pytest.skip("a CUDA device is present")
    import qampy_b200.pythran_dsp as dsp
    import qampy_b200.pythran_equalisation as pe
    from qampy_b200 import theory
    E = (np.ones((2, 64)) + 0j).astype(np.complex64)
    w = theory.init_taps(5, 2, np.complex64)
    with pytest.raises(_lib.QampyB200Error, match="no CPU fallback|CUDA"):
        pe.apply_filter_to_signal(E, 2, w)
    with pytest.raises(_lib.QampyB200Error):
        pe.train_equaliser(E, 10, 1, 2, 1e-3, w, np.arange(2), False,
                           theory.reshape_symbols(None, "mcma", 4, np.complex64, 2), "mcma")
    with pytest.raises(_lib.QampyB200Error):
        dsp.bps(E[0], theory.bps_test_angles(8, np.float32), theory.normalised_symbols(4), 4)
    with pytest.raises(_lib.QampyB200Error):
        _lib.require_device()
    import qampy_b200.equalisation as eq
    with pytest.raises(_lib.QampyB200Error):
        eq.equalise_signal(E, 2, 1e-3, 4, Ntaps=5)
