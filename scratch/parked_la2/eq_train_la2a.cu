// Depth-two look-ahead single-stream trainer, ADAPTIVE step size (see eq_train_la2.cu).
#include "eq_train_la2.cuh"

namespace qb {

int train_la2_adapt(const TrainParams<float> &p, const FastGeom &g, size_t smem, int nq, cudaStream_t st)
{
    return nq == 2 ? launch_la2_method<2, true>(p, g, smem, st) : launch_la2_method<4, true>(p, g, smem, st);
}

}  // namespace qb
