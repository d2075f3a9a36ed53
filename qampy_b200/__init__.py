"""qampy_b200 -- B200-native (sm_100a) implementation of QAMpy's coherent-receiver hot path:
adaptive MIMO FIR equaliser (train + apply) and blind-phase-search carrier recovery.

Layers (mirroring the reference):

* ``include/qampy_b200.h`` + ``qampy_b200/csrc``: C ABI and hand-written CUDA kernels
* ``qampy_b200.pythran_equalisation`` / ``qampy_b200.pythran_dsp``: L1 drop-ins (NumPy in/out)
* ``qampy_b200.equalisation`` / ``qampy_b200.phaserecovery``: L2 drop-ins (signal resident in HBM)
* ``qampy_b200.device`` / ``qampy_b200.pipeline``: tensor-level batched (time-segment) API
* ``qampy_b200.patch``: install the above underneath an importable ``qampy``

Nothing here falls back to a CPU implementation.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the library until first use)
