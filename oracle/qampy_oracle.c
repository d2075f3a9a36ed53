/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * CPU restatement (plain C + OpenMP) of the QAMpy hot path that qampy_b200
 * re-implements in CUDA.  It restates, line by line,
 *   qampy/core/equalisation/pythran_equalisation.py:4-31, 37-76, 130-265
 *   qampy/core/pythran_dsp.py:16-42, 47-85, 137-153
 * of the reference checkout (ChalmersPhotonicsLab/QAMpy @ 918723a).  The
 * reference's own kernels are Pythran sources; Pythran is not installable in
 * the build image, so this file (validated against the *interpreted* reference,
 * see tests/golden/make_golden.py and tests/test_oracle_golden.py) is what the
 * parity tests check the CUDA kernels against, and -- compiled with the
 * reference's own flags (setup.py:24-31) -- what bench.py times as the CPU
 * baseline ("port").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load this library.  Nothing under qampy_b200/ does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

enum {
    QO_CMA = 0,
    QO_CMA2 = 1,
    QO_SGNCMA = 2,
    QO_MCMA = 3,
    QO_RDE = 4,
    QO_MRDE = 5,
    QO_SBD = 6,
    QO_SBD_DATA = 7,
    QO_MDDMA = 8,
    QO_DD = 9
};

#define REAL float
#define SUF _f32
#define FABS fabsf
#include "qo_kernels.inc"
#undef REAL
#undef SUF
#undef FABS

#define REAL double
#define SUF _f64
#define FABS fabs
#include "qo_kernels.inc"
#undef REAL
#undef SUF
#undef FABS

#ifdef _OPENMP
#include <omp.h>
int qo_max_threads(void) { return omp_get_max_threads(); }
void qo_set_threads(int n) { omp_set_num_threads(n); }
#else
int qo_max_threads(void) { return 1; }
void qo_set_threads(int n) { (void)n; }
#endif
