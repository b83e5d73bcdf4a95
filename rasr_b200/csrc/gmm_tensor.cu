// gmm_tensor.cu -- tensor-core formulation of the pooled-covariance GMM scorer (RB_GMM_BATCH_TENSOR).
// Placeholder translation unit until the split-precision tcgen05 path lands: creation reports
// RB_ERR_UNSUPPORTED so callers fail loudly instead of silently getting another code path.
#include "common.cuh"

struct rb_gmm_tensor {};

int rb_gmm_tensor_create(const rb_mixture_set*, const rb::DeviceInfo&, cudaStream_t, rb_gmm_tensor** out) {
    *out = nullptr;
    rb::set_error("RB_GMM_BATCH_TENSOR is not available in this build");
    return RB_ERR_UNSUPPORTED;
}
void rb_gmm_tensor_destroy(rb_gmm_tensor* t) {
    delete t;
}
int rb_gmm_tensor_score(rb_gmm_tensor*, const float*, long, float*, cudaStream_t) {
    rb::set_error("RB_GMM_BATCH_TENSOR is not available in this build");
    return RB_ERR_UNSUPPORTED;
}
