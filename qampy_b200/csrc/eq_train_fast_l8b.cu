// Instantiations of the LPS = 8 lanes-per-stream training kernel (eq_train_fast.cuh), 8 / 12 taps per lane
// (one translation unit per group of shapes: the build compiles them in parallel).
#include "eq_train_fast.cuh"

namespace qb {

int train_fast_l8_nqb(int nq, const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    switch (nq) {
    case 8: return launch_sub_method<8, 8>(p, g, smem, st);
    case 12: return launch_sub_method<8, 12>(p, g, smem, st);
    default: return QB_ERR_UNSUPPORTED;
    }
}

}  // namespace qb
