// LPS = 8 lanes-per-stream training kernel (eq_train_fast.cuh): geometry and dispatch to the translation units that
// hold the instantiations (eq_train_fast_l8a/b/c.cu).
#include "eq_train_fast.cuh"

namespace qb {

int train_fast_l8_nqa(int nq, const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st);
int train_fast_l8_nqb(int nq, const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st);
int train_fast_l8_nqc(int nq, const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st);

// Returns 1 if launched, 0 if the shape does not fit this layout, < 0 on error.
int train_fast_l8(TrainParams<float> p, cudaStream_t st)
{
    FastGeom g;
    size_t smem = 0;
    const int nq = fast_geometry<8>(p, g, smem, 16);
    if (!nq) return 0;
    int rc;
    switch (nq) {
    case 2: case 4: case 6: rc = train_fast_l8_nqa(nq, p, g, smem, st); break;
    case 8: case 12: rc = train_fast_l8_nqb(nq, p, g, smem, st); break;
    case 16: rc = train_fast_l8_nqc(nq, p, g, smem, st); break;
    default: return 0;
    }
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
