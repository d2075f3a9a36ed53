// Instantiations of the LPS = 8 lanes-per-stream training kernels (eq_train_fast.cuh, eq_train_la.cuh).
#include <stdlib.h>

#include "eq_train_la.cuh"

namespace qb {

// Returns 1 if launched, 0 if the shape does not fit this layout, < 0 on error.
int train_fast_l8(TrainParams<float> p, cudaStream_t st)
{
    FastGeom g;
    size_t smem = 0;
    const int nq = fast_geometry<8>(p, g, smem, 16);
    if (!nq) return 0;
    // QB_TRAIN_LA=1 selects the look-ahead form of the recurrence (eq_train_la.cuh).  It is exact algebra
    // and parity-tested, but measured SLOWER than the direct form on B200 (394 vs 310 cycles per symbol,
    // profiles/README.md): ptxas does not overlap the extra work with the shuffle latency.  Kept opt-in.
    const char *la = getenv("QB_TRAIN_LA");
    if (la && la[0] == '1') {
        int r = 0;
        switch (nq) {
        case 6: r = try_la<8, 6>(p, g, st); break;
        case 12: r = try_la<8, 12>(p, g, st); break;
        default: break;
        }
        if (r != 0) return r;
    }
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
