// Instantiations of the look-ahead training kernel (eq_train_la.cuh), 8 lanes per stream.  Fixed step size here, the
// adaptive step size in eq_train_la_l8a.cu (build parallelism).
#include "eq_train_la.cuh"

namespace qb {

int train_la_l8_adapt(const TrainParams<float> &p, const FastGeom &g, size_t smem, int nq, cudaStream_t st);

// Returns 1 if launched, 0 if the shape does not fit this kernel (caller falls back), < 0 on error.
int train_la_l8(TrainParams<float> p, cudaStream_t st)
{
    FastGeom g;
    size_t smem = 0;
    const int nq = la_geometry<8>(p, g, smem);
    if (nq != 6 && nq != 12) return 0;
    int rc;
    if (p.adaptive) rc = train_la_l8_adapt(p, g, smem, nq, st);
    else if (nq == 6) rc = launch_la_method<8, 6, false>(p, g, smem, st);
    else rc = launch_la_method<8, 12, false>(p, g, smem, st);
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
