// Fast-path selection for eq_train (complex64, os = 2): picks how many lanes share a stream.
//
// With few streams the kernel is latency bound (one warp per SM sub-partition cannot hide its own
// dependent chain), with thousands it is bound by instruction issue.  Fewer lanes per stream means
// fewer instructions per trained symbol (shorter shuffle reduction, more register-window reuse) but
// also fewer warps; so the widest layout that still leaves about two warps per sub-partition wins.
#include <stdlib.h>

#include "eq_train_common.cuh"

namespace qb {

int train_fast_l8(TrainParams<float> p, cudaStream_t st);
int train_fast_l16(TrainParams<float> p, cudaStream_t st);
int train_fast_l32(TrainParams<float> p, cudaStream_t st);

static int sm_count()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// Returns 1 if the fast path took the job, 0 if the shape is outside it (caller falls back), <0 on error.
int train_fast_try(TrainParams<float> p, cudaStream_t st)
{
    if (p.os != 2) return 0;
    if (!(p.nmodes == 1 || p.nmodes == 2 || p.nmodes == 4 || p.nmodes == 8)) return 0;
    p.nsym_smem = (p.method == QB_SBD_DATA) ? 0 : p.K;
    // QB_TRAIN_LPS = 8 | 16 | 32 forces a layout (tests, tuning)
    int forced = 0;
    if (const char *e = getenv("QB_TRAIN_LPS")) forced = atoi(e);
    const long long target_warps = 2LL * 4 * sm_count();   // ~2 warps per SM sub-partition
    int order[3];
    if (forced == 8 || forced == 16 || forced == 32) {
        order[0] = forced; order[1] = forced == 8 ? 16 : 8; order[2] = forced == 32 ? 16 : 32;
    } else if (p.nstreams * 8 / 32 >= target_warps) {
        order[0] = 8; order[1] = 16; order[2] = 32;
    } else if (p.nstreams * 16 / 32 >= target_warps / 2) {
        order[0] = 16; order[1] = 8; order[2] = 32;
    } else {
        order[0] = 8; order[1] = 16; order[2] = 32;    // few streams: fewest instructions per symbol
    }
    for (int k = 0; k < 3; k++) {
        int rc = 0;
        if (order[k] == 8) rc = train_fast_l8(p, st);
        if (order[k] == 16) rc = train_fast_l16(p, st);
        if (order[k] == 32) rc = train_fast_l32(p, st);
        if (rc != 0) return rc;
    }
    return 0;
}

}  // namespace qb
