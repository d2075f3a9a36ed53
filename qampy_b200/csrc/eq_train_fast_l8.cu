// Instantiations of the LPS = 8 lanes-per-stream training kernel (eq_train_fast.cuh).
#include "eq_train_fast.cuh"

namespace qb {

// Returns 1 if launched, 0 if the shape does not fit this layout, < 0 on error.
int train_fast_l8(TrainParams<float> p, cudaStream_t st)
{
    FastGeom g;
    size_t smem = 0;
    const int nq = fast_geometry<8>(p, g, smem, 16);
    if (!nq) return 0;
    int rc;
    switch (nq) {
    case 2: rc = launch_sub_method<8, 2>(p, g, smem, st); break;
    case 4: rc = launch_sub_method<8, 4>(p, g, smem, st); break;
    case 6: rc = launch_sub_method<8, 6>(p, g, smem, st); break;
    case 8: rc = launch_sub_method<8, 8>(p, g, smem, st); break;
    case 12: rc = launch_sub_method<8, 12>(p, g, smem, st); break;
    case 16: rc = launch_sub_method<8, 16>(p, g, smem, st); break;
    default: return 0;
    }
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
