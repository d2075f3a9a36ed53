// Entry of the one-stream-per-CTA training kernel (eq_train_cta.cuh): picks the instantiation for the number of
// input polarisations.  Returns 1 if the launch(es) were issued, 0 if the shape is outside this kernel (the caller
// falls back to the warp-per-stream layouts), < 0 on error.
#include "eq_train_common.cuh"

namespace qb {

int train_cta_nm1(TrainParams<float> p, cudaStream_t st);
int train_cta_nm2(TrainParams<float> p, cudaStream_t st);
int train_cta_nm4(TrainParams<float> p, cudaStream_t st);

int train_cta_try(TrainParams<float> p, cudaStream_t st)
{
    switch (p.nmodes) {
    case 1: return train_cta_nm1(p, st);
    case 2: return train_cta_nm2(p, st);
    case 4: return train_cta_nm4(p, st);
    default: return 0;
    }
}

}  // namespace qb
