// Instantiations of the LPS = 32 lanes-per-stream training kernel (see eq_train_fast.cuh).
#include "eq_train_fast.cuh"

namespace qb {

// Returns 1 if launched, 0 if the shape does not fit this layout, < 0 on error.
int train_fast_l32(TrainParams<float> p, cudaStream_t st)
{
    FastGeom g;
    size_t smem = 0;
    const int nq = fast_geometry<32>(p, g, smem, 4);
    if (!nq) return 0;
    int rc;
    switch (nq) {
    case 2: rc = launch_sub_method<32, 2>(p, g, smem, st); break;
    case 4: rc = launch_sub_method<32, 4>(p, g, smem, st); break;
    default: return 0;
    }
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
