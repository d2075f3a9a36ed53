// Instantiation of the one-stream-per-CTA training kernel (eq_train_cta.cuh) for 1 input polarisation(s); one
// translation unit per value so that they compile side by side.
#include "eq_train_cta.cuh"

namespace qb {

int train_cta_nm1(TrainParams<float> p, cudaStream_t st) { return train_cta_nm<1>(p, st); }

}  // namespace qb
