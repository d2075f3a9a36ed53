// Instantiations of the look-ahead training kernel (eq_train_la.cuh), 8 lanes per stream.
#include "eq_train_la.cuh"

namespace qb {

// Returns 1 if launched, 0 if the shape does not fit this kernel (caller falls back), < 0 on error.
int train_la_l8(TrainParams<float> p, cudaStream_t st)
{
    if (p.adaptive) return 0;                       // the step size would sit on the serial chain: direct form
    FastGeom g;
    size_t smem = 0;
    const int nq = la_geometry<8>(p, g, smem);
    if (!nq) return 0;
    int rc;
    switch (nq) {
    case 6: rc = launch_la_method<8, 6>(p, g, smem, st); break;
    case 12: rc = launch_la_method<8, 12>(p, g, smem, st); break;
    default: return 0;
    }
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
