// Instantiations of the look-ahead training kernel (eq_train_la.cuh) with ONE stream per warp: the latency
// layout for calls with few streams (QB_LAYOUT_LATENCY: the reference's own call shape -- one capture, one stream
// per trained mode), where the time of a call is the serial depth of a stream and not the machine's throughput.
#include "eq_train_la.cuh"

namespace qb {

// Returns 1 if launched, 0 if the shape does not fit this kernel (caller falls back), < 0 on error.
int train_la_l32(TrainParams<float> p, cudaStream_t st)
{
    if (p.adaptive) return 0;                       // the step size would sit on the serial chain: direct form
    FastGeom g;
    size_t smem = 0;
    const int nq = la_geometry<32>(p, g, smem);
    int rc;
    switch (nq) {
    case 2: rc = launch_la_method<32, 2>(p, g, smem, st); break;
    case 4: rc = launch_la_method<32, 4>(p, g, smem, st); break;
    default: return 0;
    }
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
