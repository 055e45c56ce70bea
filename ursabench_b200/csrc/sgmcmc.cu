// K1: fused SG-MCMC update over a flat fp32 buffer (SGLD / SGHMC / cSGLD / cSGHMC).
// One launch replaces optimSGHMC.step's per-tensor ATen loop (reference
// inference/optim_sghmc.py:30-68).  HBM-bound: 20 B/param (momentum) or 12 B/param (SGLD).
//
// Layout: p, g, v (and snapshot / external noise) are flat, 16-byte aligned; every thread moves
// 128-bit vectors, UNROLL of them per stream per trip so that >= 96 B/thread are in flight before the
// first use; the grid is persistent (SMs x resident CTAs) and grid-strides.
#include "common.cuh"

namespace ursa {

enum { NOISE_NONE = 0, NOISE_EXTERNAL = 1, NOISE_PHILOX = 2 };

struct StepArgs {
    float *p, *g, *v, *snap;
    const float *noise;
    int64_t n;
    float lr, momentum, wd_over_n, noise_mul, noise_div, noise_scale;
    uint32_t flags;
    uint2 key;
    uint64_t step, elem_offset;
    const float *dyn;   // optional device-resident [lr, momentum, wd_over_n, noise_scale, step_lo, step_hi] (CUDA-graph replay)
};

// The arithmetic of one element, in the reference's operation order.  __fmul_rn / __fadd_rn pin the
// roundings so nvcc cannot contract them into different FMAs than torch's CPU kernels use
// (add(alpha) is an FMA there; mul_ and the noise scaling are separate roundings).
template <bool HAS_MOM, int NOISE>
__device__ __forceinline__ void update_one(float &p, float g, float &v, float z, const StepArgs &a, bool first) {
    float d = g;
    if (a.wd_over_n != 0.f) d = fmaf(a.wd_over_n, p, g);                 // :47-48
    float u;
    if (HAS_MOM) {
        const float buf = first ? d : v;                                 // :51-52
        u = fmaf(-a.lr, d, __fmul_rn(buf, a.momentum));                  // :53/:56
    } else {
        u = __fmul_rn(d, -a.lr);                                         // :62
    }
    if (NOISE == NOISE_EXTERNAL) u = __fadd_rn(u, __fdiv_rn(__fmul_rn(z, a.noise_mul), a.noise_div));  // :63-64
    if (NOISE == NOISE_PHILOX) u = fmaf(z, a.noise_scale, u);
    p = __fadd_rn(p, u);                                                 // :65
    if (HAS_MOM) v = u;                                                  // :66-67
}

template <bool HAS_MOM, int NOISE>
__device__ __forceinline__ void update_vec(float4 &p, const float4 &g, float4 &v, const float4 &z,
                                           const StepArgs &a, bool first) {
    update_one<HAS_MOM, NOISE>(p.x, g.x, v.x, z.x, a, first);
    update_one<HAS_MOM, NOISE>(p.y, g.y, v.y, z.y, a, first);
    update_one<HAS_MOM, NOISE>(p.z, g.z, v.z, z.z, a, first);
    update_one<HAS_MOM, NOISE>(p.w, g.w, v.w, z.w, a, first);
}

constexpr int kThreads = 256;
constexpr int kUnroll = 2;

template <bool HAS_MOM, int NOISE>
__global__ void __launch_bounds__(kThreads, 4) sgmcmc_step_kernel(const StepArgs a_in) {
    StepArgs a = a_in;
    if (a.dyn) {        // scalars that change between replays of a captured graph live in device memory
        a.lr = __ldg(a.dyn + 0);
        a.momentum = __ldg(a.dyn + 1);
        a.wd_over_n = __ldg(a.dyn + 2);
        a.noise_scale = __ldg(a.dyn + 3);
        a.noise_mul = a.noise_scale;
        a.noise_div = 1.f;
        a.step = ((uint64_t)__float_as_uint(__ldg(a.dyn + 5)) << 32) | (uint64_t)__float_as_uint(__ldg(a.dyn + 4));
    }
    const bool first = (a.flags & URSA_STEP_FIRST) != 0;
    const bool zero_g = (a.flags & URSA_STEP_ZERO_GRAD) != 0;
    const int64_t nvec = a.n >> 2;
    float4 *__restrict__ p4 = reinterpret_cast<float4 *>(a.p);
    float4 *__restrict__ g4 = reinterpret_cast<float4 *>(a.g);
    float4 *__restrict__ v4 = reinterpret_cast<float4 *>(a.v);
    float4 *__restrict__ s4 = reinterpret_cast<float4 *>(a.snap);
    const float4 *__restrict__ z4 = reinterpret_cast<const float4 *>(a.noise);
    const uint64_t blk0 = a.elem_offset >> 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    // main trips: kUnroll independent 128-bit loads per stream before any use
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
        float4 p[kUnroll], g[kUnroll], v[kUnroll], z[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int64_t k = i + j * stride;
            p[j] = p4[k];
            g[j] = g4[k];
            if (HAS_MOM && !first) v[j] = v4[k]; else v[j] = zero4;
            if (NOISE == NOISE_EXTERNAL) z[j] = __ldg(z4 + k); else z[j] = zero4;
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int64_t k = i + j * stride;
            if (NOISE == NOISE_PHILOX) z[j] = philox_normal4(blk0 + (uint64_t)k, a.step, a.key);
            update_vec<HAS_MOM, NOISE>(p[j], g[j], v[j], z[j], a, first);
            p4[k] = p[j];
            if (HAS_MOM) v4[k] = v[j];
            if (s4) s4[k] = p[j];
            if (zero_g) g4[k] = zero4;
        }
    }
    for (; i < nvec; i += stride) {
        float4 p = p4[i], g = g4[i], v = zero4, z = zero4;
        if (HAS_MOM && !first) v = v4[i];
        if (NOISE == NOISE_EXTERNAL) z = __ldg(z4 + i);
        if (NOISE == NOISE_PHILOX) z = philox_normal4(blk0 + (uint64_t)i, a.step, a.key);
        update_vec<HAS_MOM, NOISE>(p, g, v, z, a, first);
        p4[i] = p;
        if (HAS_MOM) v4[i] = v;
        if (s4) s4[i] = p;
        if (zero_g) g4[i] = zero4;
    }
    // scalar tail (n % 4 elements), same Philox block / lanes as the vector path would use
    const int tail = (int)(a.n & 3);
    if (tail && blockIdx.x == 0 && threadIdx.x < tail) {
        const int64_t e = (nvec << 2) + threadIdx.x;
        float p = a.p[e], g = a.g[e], v = 0.f, z = 0.f;
        if (HAS_MOM && !first) v = a.v[e];
        if (NOISE == NOISE_EXTERNAL) z = a.noise[e];
        if (NOISE == NOISE_PHILOX) z = f4_get(philox_normal4(blk0 + (uint64_t)nvec, a.step, a.key), threadIdx.x);
        update_one<HAS_MOM, NOISE>(p, g, v, z, a, first);
        a.p[e] = p;
        if (HAS_MOM) a.v[e] = v;
        if (a.snap) a.snap[e] = p;
        if (zero_g) a.g[e] = 0.f;
    }
}

__global__ void __launch_bounds__(kThreads) philox_normal_kernel(float *out, int64_t n, uint2 key, uint64_t step,
                                                                 uint64_t elem_offset) {
    const int64_t nblk = (n + 3) >> 2;
    const uint64_t blk0 = elem_offset >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * kThreads) {
        const float4 z = philox_normal4(blk0 + (uint64_t)i, step, key);
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (4 * i + l < n) out[4 * i + l] = f4_get(z, l);
    }
}

__global__ void set_dyn_kernel(float *dyn, float lr, float momentum, float wd_over_n, float noise_scale, uint64_t step) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        dyn[0] = lr;
        dyn[1] = momentum;
        dyn[2] = wd_over_n;
        dyn[3] = noise_scale;
        dyn[4] = __uint_as_float((uint32_t)step);
        dyn[5] = __uint_as_float((uint32_t)(step >> 32));
    }
}

static int grid_for(int64_t nvec, int ctas_per_sm) {
    const int64_t want = (nvec + (int64_t)kThreads * kUnroll - 1) / ((int64_t)kThreads * kUnroll);
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace ursa

using namespace ursa;

extern "C" int ursa_sgmcmc_step(float *p, float *g, float *v, float *snapshot, const float *noise, int64_t n,
                                float lr, float momentum, float wd_over_n, float noise_mul, float noise_div,
                                uint32_t flags, uint64_t seed, uint64_t step, uint64_t elem_offset,
                                void *stream) {
    URSA_REQUIRE(n >= 0, "ursa_sgmcmc_step: n < 0");
    URSA_REQUIRE(n == 0 || (p && g), "ursa_sgmcmc_step: p and g must be non-null");
    URSA_REQUIRE(momentum >= 0.f, "ursa_sgmcmc_step: Invalid momentum value: %g", (double)momentum);
    URSA_REQUIRE(momentum == 0.f || v || n == 0, "ursa_sgmcmc_step: v is required when momentum != 0");
    URSA_REQUIRE(aligned16(p) && aligned16(g) && aligned16(v) && aligned16(snapshot) && aligned16(noise),
                 "ursa_sgmcmc_step: buffers must be 16-byte aligned");
    URSA_REQUIRE((elem_offset & 3u) == 0, "ursa_sgmcmc_step: elem_offset must be a multiple of 4");
    const bool add_noise = (flags & URSA_STEP_NOISE) != 0;
    URSA_REQUIRE(!add_noise || noise_div != 0.f, "ursa_sgmcmc_step: noise_div == 0");
    if (n == 0) return URSA_OK;
    StepArgs a;
    a.p = p; a.g = g; a.v = v; a.snap = snapshot; a.noise = noise; a.n = n;
    a.lr = lr; a.momentum = momentum; a.wd_over_n = wd_over_n;
    a.noise_mul = noise_mul; a.noise_div = noise_div;
    a.noise_scale = add_noise ? (float)((double)noise_mul / (double)noise_div) : 0.f;
    a.flags = flags;
    a.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    a.step = step; a.elem_offset = elem_offset; a.dyn = nullptr;
    const int mode = !add_noise ? NOISE_NONE : (noise ? NOISE_EXTERNAL : NOISE_PHILOX);
    const bool mom = momentum != 0.f;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n >> 2, 8);
#define URSA_K1(M, Z) sgmcmc_step_kernel<M, Z><<<grid, kThreads, 0, st>>>(a)
    if (mom) {
        if (mode == NOISE_NONE) URSA_K1(true, NOISE_NONE);
        else if (mode == NOISE_EXTERNAL) URSA_K1(true, NOISE_EXTERNAL);
        else URSA_K1(true, NOISE_PHILOX);
    } else {
        if (mode == NOISE_NONE) URSA_K1(false, NOISE_NONE);
        else if (mode == NOISE_EXTERNAL) URSA_K1(false, NOISE_EXTERNAL);
        else URSA_K1(false, NOISE_PHILOX);
    }
#undef URSA_K1
    URSA_LAUNCH_CHECK("sgmcmc_step_kernel");
    return URSA_OK;
}

extern "C" int ursa_philox_normal(float *out, int64_t n, uint64_t seed, uint64_t step, uint64_t elem_offset,
                                  void *stream) {
    URSA_REQUIRE(out && n >= 0, "ursa_philox_normal: bad arguments");
    URSA_REQUIRE((elem_offset & 3u) == 0, "ursa_philox_normal: elem_offset must be a multiple of 4");
    if (n == 0) return URSA_OK;
    const int grid = grid_for((n + 3) >> 2, 8);
    philox_normal_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        out, n, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), step, elem_offset);
    URSA_LAUNCH_CHECK("philox_normal_kernel");
    return URSA_OK;
}

extern "C" int ursa_sgmcmc_step_dyn(float *p, float *g, float *v, float *snapshot, const float *noise, int64_t n,
                                    const float *dyn, uint32_t flags, uint64_t seed, uint64_t elem_offset,
                                    void *stream) {
    URSA_REQUIRE(n >= 0 && dyn, "ursa_sgmcmc_step_dyn: bad arguments");
    URSA_REQUIRE(n == 0 || (p && g), "ursa_sgmcmc_step_dyn: p and g must be non-null");
    URSA_REQUIRE(aligned16(p) && aligned16(g) && aligned16(v) && aligned16(snapshot) && aligned16(noise),
                 "ursa_sgmcmc_step_dyn: buffers must be 16-byte aligned");
    URSA_REQUIRE((elem_offset & 3u) == 0, "ursa_sgmcmc_step_dyn: elem_offset must be a multiple of 4");
    if (n == 0) return URSA_OK;
    StepArgs a;
    a.p = p; a.g = g; a.v = v; a.snap = snapshot; a.noise = noise; a.n = n;
    a.lr = a.momentum = a.wd_over_n = a.noise_mul = a.noise_scale = 0.f; a.noise_div = 1.f;
    a.flags = flags;
    a.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    a.step = 0; a.elem_offset = elem_offset; a.dyn = dyn;
    const bool add_noise = (flags & URSA_STEP_NOISE) != 0;
    const int mode = !add_noise ? NOISE_NONE : (noise ? NOISE_EXTERNAL : NOISE_PHILOX);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n >> 2, 8);
#define URSA_K1(M, Z) sgmcmc_step_kernel<M, Z><<<grid, kThreads, 0, st>>>(a)
    if (v) {                       // momentum buffer present <=> momentum != 0 (the value itself is dynamic)
        if (mode == NOISE_NONE) URSA_K1(true, NOISE_NONE);
        else if (mode == NOISE_EXTERNAL) URSA_K1(true, NOISE_EXTERNAL);
        else URSA_K1(true, NOISE_PHILOX);
    } else {
        if (mode == NOISE_NONE) URSA_K1(false, NOISE_NONE);
        else if (mode == NOISE_EXTERNAL) URSA_K1(false, NOISE_EXTERNAL);
        else URSA_K1(false, NOISE_PHILOX);
    }
#undef URSA_K1
    URSA_LAUNCH_CHECK("sgmcmc_step_kernel(dyn)");
    return URSA_OK;
}

extern "C" int ursa_sgmcmc_set_dyn(float *dyn, float lr, float momentum, float wd_over_n, float noise_scale,
                                   uint64_t step, void *stream) {
    URSA_REQUIRE(dyn, "ursa_sgmcmc_set_dyn: dyn is null");
    set_dyn_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dyn, lr, momentum, wd_over_n, noise_scale, step);
    URSA_LAUNCH_CHECK("set_dyn_kernel");
    return URSA_OK;
}
