"""ctypes binding of the C ABI declared in ``include/qampy_b200.h``.

There is deliberately no fallback: if ``libqampy_b200.so`` is missing or the CUDA device is not
usable, importing / calling raises.  (Build with ``python -m qampy_b200.build``.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libqampy_b200.so")

QB_C64, QB_C128 = 0, 1
METHODS = {"cma": 0, "cma2": 1, "sgncma": 2, "mcma": 3, "rde": 4, "mrde": 5, "sbd": 6, "sbd_data": 7,
           "mddma": 8, "dd": 9}

_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p

# name -> argtypes, exactly the prototypes in include/qampy_b200.h
PROTOTYPES = {
    "qb_version": ([], _int),
    "qb_last_error": ([], ctypes.c_char_p),
    "qb_device_count": ([], _int),
    "qb_method_from_name": ([ctypes.c_char_p], _int),
    "qb_launch_count": ([], _i64),
    "qb_set_train_layout": ([_int], _int),
    "qb_set_bps_accumulation": ([_int], _int),
    "qb_set_option": ([ctypes.c_char_p, ctypes.c_char_p], _int),
    "qb_train_equaliser_dev": ([_int, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64,
                                _int, _vp, _i64, _int, _vp, _vp, _vp], _int),
    "qb_train_equaliser_host": ([_int, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _int,
                                 _vp, _i64, _int, _int, _vp], _int),
    "qb_apply_filter_to_signal_dev": ([_int, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64,
                                       _vp, _vp], _int),
    "qb_apply_filter_to_signal_host": ([_int, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp], _int),
    "qb_bps_dev": ([_int, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64,
                    _vp, _vp, _vp, _vp], _int),
    "qb_bps_host": ([_int, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp], _int),
    "qb_bps_rows_dev": ([_int, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64,
                         _vp, _vp, _vp, _vp], _int),
    "qb_bps_rows_host": ([_int, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp], _int),
    "qb_detect_grid_host": ([_int, _vp, _i64, _vp, _vp, _vp, _vp], _int),
    "qb_make_decision_dev": ([_int, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp], _int),
    "qb_make_decision_host": ([_int, _vp, _i64, _vp, _i64, _vp, _vp, _vp], _int),
    "qb_soft_l_value_demapper_host": ([_int, _vp, _i64, _i64, ctypes.c_double, _vp, _i64, _i64, _int, _vp], _int),
    "qb_estimate_snr_host": ([_int, _vp, _vp, _i64, _vp, _i64, _vp], _int),
    "qb_viterbiviterbi_dev": ([_int, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp], _int),
    "qb_viterbiviterbi_host": ([_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp], _int),
    "qb_synth_upsample_dev": ([_vp, _i64, _i64, _i64, _vp, _i64, _vp], _int),
    "qb_synth_specmul_dev": ([_vp, _i64, _i64, _vp, _vp], _int),
    "qb_synth_crop_norm_dev": ([_vp, _i64, _i64, _i64, _i64, _i64, _vp, _int, _vp, _vp], _int),
    "qb_synth_pmd_dev": ([_vp, _i64, ctypes.c_double, ctypes.c_double, ctypes.c_double, _vp], _int),
    "qb_synth_tail_dev": ([_int, _vp, _i64, _i64, _vp, ctypes.c_double, ctypes.c_uint64, _i64, _i64, _vp, _vp, _i64,
                           _vp, _vp], _int),
    "qb_freq_shift_dev": ([_int, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp], _int),
    "qb_pilot_cpe_dev": ([_int, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp], _int),
    "qb_select_angles_dev": ([_int, _vp, _i64, _i64, _vp, _i64, _vp, _vp], _int),
    "qb_select_angles_host": ([_int, _vp, _i64, _i64, _vp, _i64, _vp], _int),
}

_lib = None


class QampyB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "qampy_b200: %s not found -- build it with `python -m qampy_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def check(rc):
    """Map a qb_status to the exception type the reference raises for the same condition."""
    if rc >= 0:
        return rc
    msg = load().qb_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise NotImplementedError(msg)
    if rc == -4:
        raise MemoryError(msg)
    raise QampyB200Error(msg)


def require_device():
    n = check(load().qb_device_count())
    if n < 1:
        raise QampyB200Error("qampy_b200: no CUDA device visible; there is no CPU fallback")
    return n


def launch_count():
    return int(load().qb_launch_count())
