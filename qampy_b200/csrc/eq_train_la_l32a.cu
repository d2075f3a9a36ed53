// Look-ahead training kernel, one stream per warp, ADAPTIVE step size (see eq_train_la_l32.cu).
#include "eq_train_la.cuh"

namespace qb {

int train_la_l32_adapt(const TrainParams<float> &p, const FastGeom &g, size_t smem, int nq, cudaStream_t st)
{
    return nq == 2 ? launch_la_method<32, 2, true>(p, g, smem, st) : launch_la_method<32, 4, true>(p, g, smem, st);
}

}  // namespace qb
