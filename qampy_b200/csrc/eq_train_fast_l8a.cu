// Instantiations of the LPS = 8 lanes-per-stream training kernel (eq_train_fast.cuh), 2 / 4 / 6 taps per lane
// (one translation unit per group of shapes: the build compiles them in parallel).
#include "eq_train_fast.cuh"

namespace qb {

int train_fast_l8_nqa(int nq, const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    switch (nq) {
    case 2: return launch_sub_method<8, 2>(p, g, smem, st);
    case 4: return launch_sub_method<8, 4>(p, g, smem, st);
    case 6: return launch_sub_method<8, 6>(p, g, smem, st);
    default: return QB_ERR_UNSUPPORTED;
    }
}

}  // namespace qb
