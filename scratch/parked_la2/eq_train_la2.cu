// Instantiations of the depth-two look-ahead single-stream trainer (eq_train_la2.cuh): fixed step size here, the
// adaptive step size in eq_train_la2a.cu (build parallelism).
#include <stdlib.h>

#include "eq_train_la2.cuh"

namespace qb {

int train_la2_adapt(const TrainParams<float> &p, const FastGeom &g, size_t smem, int nq, cudaStream_t st);

// Returns 1 if launched, 0 if the shape does not fit this kernel (caller falls back to the depth-one kernel), < 0 on error.
// QB_TRAIN_LA2=0 keeps the depth-one kernel (tests and A/B timing run both).
int train_la2(TrainParams<float> p, cudaStream_t st)
{
    const char *e = getenv("QB_TRAIN_LA2");
    if (e && e[0] == '0') return 0;
    FastGeom g;
    size_t smem = 0;
    const int nq = la2_geometry(p, g, smem);
    if (nq != 2 && nq != 4) return 0;
    int rc;
    if (p.adaptive) rc = train_la2_adapt(p, g, smem, nq, st);
    else if (nq == 2) rc = launch_la2_method<2, false>(p, g, smem, st);
    else rc = launch_la2_method<4, false>(p, g, smem, st);
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
