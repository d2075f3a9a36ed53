// Fast-path selection for eq_train (complex64, os = 2): picks how many lanes share a stream.
//
// Fewer lanes per stream means fewer instructions per trained symbol (shorter shuffle reduction, more
// register-window reuse) but also fewer warps.  Measured on B200 (C3, 2440 streams of 8192 symbols):
// 8 lanes 2.0 ms/pass, 16 lanes 2.0 ms, 32 lanes 3.0 ms -- the dependent chain, not occupancy, sets the
// pace, so the 8-lane layout is the default and the 16-lane one is kept for shapes the 8-lane one
// cannot hold (nmodes*ntaps > 128) and for tests.
#include <stdlib.h>

#include "eq_train_common.cuh"

namespace qb {

int train_la_l8(TrainParams<float> p, cudaStream_t st);
int train_la_l32(TrainParams<float> p, cudaStream_t st);
int train_fast_l8(TrainParams<float> p, cudaStream_t st);
int train_fast_l16(TrainParams<float> p, cudaStream_t st);

// Layout of the calling thread's training launches (qb_set_train_layout): 0 = throughput (fixed 8-lane layout),
// 1 = latency (one stream per warp where that kernel is instantiated).
static thread_local int g_train_layout = 0;
int set_train_layout(int layout)
{
    const int old = g_train_layout;
    g_train_layout = layout ? 1 : 0;
    return old;
}

// Returns 1 if the fast path took the job, 0 if the shape is outside it (caller falls back), <0 on error.
int train_fast_try(TrainParams<float> p, cudaStream_t st)
{
    if (p.os != 2) return 0;
    if (!(p.nmodes == 1 || p.nmodes == 2 || p.nmodes == 4 || p.nmodes == 8)) return 0;
    p.nsym_smem = (p.method == QB_SBD_DATA) ? 0 : p.K;
    // methods that search their alphabet get scratch for the grid slicer behind the staged constants
    // (eq_train_fast.cuh, detect_grid): 32 axis levels + one byte per alphabet point
    const bool searched = p.method == QB_SBD || p.method == QB_DD || p.method == QB_MDDMA;
    const int side = grid_side(p.K);     // levels per axis of the smallest square grid that holds K points
    p.nsym_pitch = p.nsym_smem + ((searched && p.K >= 4 && p.K <= GRID_MAX_K && side <= 16) ? 16 + (side * side + 7) / 8 : 0);
    // option TRAIN_LPS = 8 | 16 (qb_set_option) forces a layout (tests, tuning)
    const int forced = option_int(OPT_TRAIN_LPS, 0);
    // Default: 8 lanes per stream (fewest instructions per trained symbol; measured fastest from 2 to
    // thousands of streams on B200).  A fixed layout also keeps a segment's result independent of how
    // many other segments share the launch.
    // Look-ahead form of the recurrence (eq_train_la.cuh) where it is instantiated: fixed step size, 8 lanes per
    // stream, 6 or 12 taps per lane.  Option TRAIN_KERNEL = direct keeps the direct form (tests run both).
    if (!forced && option_char(OPT_TRAIN_KERNEL) != 'd') {
        // The layout is the CALLER's choice, never a function of how many streams a launch holds: a stream's
        // result stays independent of what shares its launch.
        int rc = g_train_layout == 1 ? train_la_l32(p, st) : 0;
        if (rc != 0) return rc;
        rc = train_la_l8(p, st);
        if (rc != 0) return rc;
    }
    int order[2] = {8, 16};
    if (forced == 16) {
        order[0] = 16;
        order[1] = 8;
    }
    for (int k = 0; k < 2; k++) {
        int rc = 0;
        if (order[k] == 8) rc = train_fast_l8(p, st);
        if (order[k] == 16) rc = train_fast_l16(p, st);
        if (rc != 0) return rc;
    }
    return 0;
}

}  // namespace qb
