// Softmax-average + smoothed-entropy epilogue of the BMA forwards (reference tasks/prediction.py:60-63, util.py:126-144),
// shared by ursa_bma_accumulate (logits in global memory) and by the head kernels that fuse it (logits in registers).
// One warp owns one test row; class c lives in lane c % 32, register c / 32.
#pragma once
#include "common.cuh"

namespace ursa {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// softmax-average + smoothed-entropy of one row held as per-lane registers (shared with the fused forwards)
template <int PER_LANE>
__device__ __forceinline__ void softmax_accumulate_row(const float (&x)[PER_LANE], int C, int lane, float one_minus_gamma,
                                                       float gamma_over_c, float (&P)[PER_LANE], float &E) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j)
        if (lane + 32 * j < C) m = fmaxf(m, x[j]);
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j)
        if (lane + 32 * j < C) sum += expf(x[j] - m);
    const float lse = logf(warp_sum(sum));
    float h = 0.f;
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j)
        if (lane + 32 * j < C) {
            const float p = expf((x[j] - m) - lse);                                   // log_softmax().exp_()  (:60)
            P[j] = __fadd_rn(P[j], p);
            const float q = __fadd_rn(__fmul_rn(one_minus_gamma, p), gamma_over_c);   // util.py:134
            h = fmaf(q, logf(q), h);                                                  // util.py:144
        }
    E = __fadd_rn(E, -warp_sum(h));
}

}  // namespace ursa
