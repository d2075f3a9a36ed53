// Instantiations of the LPS = 8 lanes-per-stream training kernel (eq_train_fast.cuh), 16 taps per lane
// (one translation unit per group of shapes: the build compiles them in parallel).
#include "eq_train_fast.cuh"

namespace qb {

int train_fast_l8_nqc(int nq, const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    switch (nq) {
    case 16: return launch_sub_method<8, 16>(p, g, smem, st);
    default: return QB_ERR_UNSUPPORTED;
    }
}

}  // namespace qb
