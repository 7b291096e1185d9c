// gbmm_dmma.cu -- wide-band FP64 banded x banded product on DMMA tensor-core tiles.
// (placeholder until the DMMA kernel lands: reports "not handled" so lbm_dgbmm uses the
// diagonal-wise kernel.)
#include "lbm_common.cuh"

int lbm_gbmm_dmma_f64(lbm_handle_t, int64_t, int64_t, int64_t, double, int64_t, int64_t, const double *, int64_t,
                      int64_t, int64_t, const double *, int64_t, double, int64_t, int64_t, double *, int64_t,
                      bool *handled) {
    *handled = false;
    return 0;
}
