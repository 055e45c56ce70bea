// K3 (MLP) on tcgen05: placeholder until the 3xTF32 UMMA path lands.
#include "common.cuh"

namespace ursa {

size_t mlp_workspace_tcgen05(int, int64_t, int, int, int) { return 0; }

int mlp_forward_tcgen05(const float *, int64_t, int, const float *, int64_t, int, int, int, float *, float *, float *,
                        double, void *, size_t, cudaStream_t) {
    set_error("ursa_bma_mlp_forward: URSA_ALGO_TCGEN05 is not built in this revision");
    return URSA_ERR_UNSUPPORTED;
}

}  // namespace ursa
