// Look-ahead training kernel, 8 lanes per stream, ADAPTIVE step size (see eq_train_la_l8.cu).
#include "eq_train_la.cuh"

namespace qb {

int train_la_l8_adapt(const TrainParams<float> &p, const FastGeom &g, size_t smem, int nq, cudaStream_t st)
{
    return nq == 6 ? launch_la_method<8, 6, true>(p, g, smem, st) : launch_la_method<8, 12, true>(p, g, smem, st);
}

}  // namespace qb
