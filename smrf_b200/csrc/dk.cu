// dk.cu -- placeholder while the kriging builder is written (replaced in the next commit).
#include "common.cuh"
using namespace smrfb;
extern "C" {
int smrfb_dk_create(int, const double*, const double*, const double*, int, int, int, const double*, const double*,
                    const double*, int, int, smrfb_handle_t* out) { if (out) *out = nullptr; set_error("DK: not built yet"); return SMRFB_ERR_LIMIT; }
int smrfb_dk_distribute(smrfb_handle_t, const double*, int, double, double, const double*, double*, double*) { set_error("DK: not built yet"); return SMRFB_ERR_LIMIT; }
int smrfb_dk_distribute_dev(smrfb_handle_t, const double*, const double*, int, double, double, const double*, double*, double*, void*) { set_error("DK: not built yet"); return SMRFB_ERR_LIMIT; }
int smrfb_dk_weights(smrfb_handle_t, const uint8_t*, double*) { set_error("DK: not built yet"); return SMRFB_ERR_LIMIT; }
int smrfb_dk_stats(smrfb_handle_t, int64_t*, int64_t*, int*) { set_error("DK: not built yet"); return SMRFB_ERR_LIMIT; }
int smrfb_krige_grid(int, int, const double*, const double*, const double*, int, double*) { set_error("DK: not built yet"); return SMRFB_ERR_LIMIT; }
}
