// K3 (PreResNet): placeholder until the sample-batched conv path lands.
#include "common.cuh"

extern "C" size_t ursa_bma_preresnet_workspace(int, int64_t, int, int, int) { return 0; }

extern "C" int ursa_bma_preresnet_forward(const float *, int64_t, const float *, int64_t, int, const float *, int64_t,
                                          int, int, float *, float *, float *, double, void *, size_t, int, void *) {
    ursa::set_error("ursa_bma_preresnet_forward: not built in this revision");
    return URSA_ERR_UNSUPPORTED;
}
