"""Golden digests of the reference's own Scripts/*_equalisation.py, run UNMODIFIED and UNPATCHED (the interpreted
Pythran sources of /root/reference) under tests/scripts_runner.py's repeatable randomness.  The unpatched runs take
minutes (pure-Python symbol loops: mrde_equaliser 84 s, 64_qam 34 s, 32_qam ~5 min), so the CPU suite compares the
PATCHED run of each script against the digests stored here instead of re-running the reference every time
(tests/test_scripts_acceptance.py; QB_SCRIPTS_LIVE=1 re-runs the reference as well).

    python tests/golden/make_golden_scripts.py        ->  tests/golden/g12_scripts.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import scripts_runner as sr  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name in ("mrde_equaliser.py", "64_qam_equalisation.py", "32_qam_equalisation.py"):
        t0 = time.time()
        ns, text = sr.run_script(name)
        print(name, "%.0f s" % (time.time() - t0), text.strip()[-200:], flush=True)
        for k, v in sr.summary(name, ns).items():
            out[name + "/" + k] = v
    np.savez_compressed(os.path.join(HERE, "g12_scripts.npz"), **out)
    print("wrote", len(out), "arrays")
