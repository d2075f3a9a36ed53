// Instantiations of the look-ahead training kernel (eq_train_la.cuh) with ONE stream per warp: the latency
// layout for calls with few streams (QB_LAYOUT_LATENCY: the reference's own call shape -- one capture, one stream
// per trained mode), where the time of a call is the serial depth of a stream and not the machine's throughput.
// Fixed step size here, the adaptive step size in eq_train_la_l32a.cu (build parallelism).
#include "eq_train_la.cuh"

namespace qb {

int train_la_l32_adapt(const TrainParams<float> &p, const FastGeom &g, size_t smem, int nq, cudaStream_t st);

// Returns 1 if launched, 0 if the shape does not fit this kernel (caller falls back), < 0 on error.
int train_la_l32(TrainParams<float> p, cudaStream_t st)
{
    FastGeom g;
    size_t smem = 0;
    const int nq = la_geometry<32>(p, g, smem);
    int rc;
    if (nq != 2 && nq != 4) return 0;
    if (p.adaptive) rc = train_la_l32_adapt(p, g, smem, nq, st);
    else if (nq == 2) rc = launch_la_method<32, 2, false>(p, g, smem, st);
    else rc = launch_la_method<32, 4, false>(p, g, smem, st);
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
