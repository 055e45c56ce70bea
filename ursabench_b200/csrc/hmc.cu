// K5: chain-batched Hamiltonian Monte Carlo over flat weights (HMC.sample, reference inference/hmc.py:62-85, whose
// arithmetic is hamiltorch.sample_model -- a third-party dependency that is absent from /root/reference, so the
// spec is the restatement in oracle/restate.py::hmc_*; parity is UNPINNED).
//
// State: theta, r (momentum) and the saved position are [C, ld] fp32, C independent chains, rows 16-byte aligned.
// All elementwise kernels treat the batch as C*ld flat elements (the step size, mass and prior precision are shared
// by the chains, as in the reference's call site hmc.py:64-75); only the energy reduction and the accept / restore
// step are per chain.  The gradient of the data term comes from PyTorch autograd (vmap over chains) as g_nll.
//
//   momentum   r = sqrt(mass) * z                                          hamiltorch gibbs: Normal(0, mass**0.5)
//   leapfrog   r += kick * grad_logp ; theta += drift * r                  kick = eps/2 | eps, drift = eps*inv_mass | 0
//              grad_logp = -(tau_out * g_nll + tau * theta)                 20 B/param (read theta, r, g; write theta, r)
//   energy     sum theta^2, sum r^2 per chain in fp64, fixed order          8 B/param
//   accept     accept[c] = log u[c] <= min(0, H_old[c] - H_new[c]) ; accepted chains commit theta -> saved,
//              rejected chains restore saved -> theta                      8 B/param
#include "common.cuh"

namespace ursa {

constexpr int kHmcThreads = 256;

static int hmc_grid(int64_t nvec, int ctas_per_sm) {
    const int64_t want = (nvec + kHmcThreads - 1) / kHmcThreads;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <bool PHILOX>
__global__ void __launch_bounds__(kHmcThreads) hmc_momentum_kernel(float *__restrict__ r, const float *__restrict__ noise,
                                                                    int64_t n, float sqrt_mass, uint2 key, uint64_t step,
                                                                    uint64_t blk0) {
    const int64_t nvec = n >> 2;
    float4 *r4 = reinterpret_cast<float4 *>(r);
    const float4 *z4 = reinterpret_cast<const float4 *>(noise);
    const int64_t stride = (int64_t)gridDim.x * kHmcThreads;
    for (int64_t i = (int64_t)blockIdx.x * kHmcThreads + threadIdx.x; i < nvec; i += stride) {
        float4 z;
        if (PHILOX) z = philox_normal4(blk0 + (uint64_t)i, step, key); else z = __ldg(z4 + i);
        r4[i] = make_float4(__fmul_rn(z.x, sqrt_mass), __fmul_rn(z.y, sqrt_mass), __fmul_rn(z.z, sqrt_mass),
                            __fmul_rn(z.w, sqrt_mass));
    }
}

struct LeapArgs {
    float *theta, *r, *snap;
    const float *g;
    int64_t n;
    float kick, drift, tau, tau_out;
};

__device__ __forceinline__ void leap_one(float &th, float &r, float g, const LeapArgs &a, bool drift) {
    const float glp = -fmaf(a.tau, th, __fmul_rn(a.tau_out, g));        // grad log p = -(tau_out*g_nll + tau*theta)
    r = __fadd_rn(r, __fmul_rn(a.kick, glp));                           // momentum += kick * grad
    if (drift) th = __fadd_rn(th, __fmul_rn(a.drift, r));               // params += (eps*inv_mass) * momentum
}

template <bool DRIFT, bool SNAP>
__global__ void __launch_bounds__(kHmcThreads, 4) hmc_leapfrog_kernel(const LeapArgs a) {
    const int64_t nvec = a.n >> 2;
    float4 *__restrict__ t4 = reinterpret_cast<float4 *>(a.theta);
    float4 *__restrict__ r4 = reinterpret_cast<float4 *>(a.r);
    float4 *__restrict__ s4 = reinterpret_cast<float4 *>(a.snap);
    const float4 *__restrict__ g4 = reinterpret_cast<const float4 *>(a.g);
    const int64_t stride = (int64_t)gridDim.x * kHmcThreads;
    int64_t i = (int64_t)blockIdx.x * kHmcThreads + threadIdx.x;
    for (; i + stride < nvec; i += 2 * stride) {             // two independent 128-bit loads per stream before use
        float4 t[2], r[2], g[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            t[j] = t4[i + j * stride];
            r[j] = r4[i + j * stride];
            g[j] = __ldg(g4 + i + j * stride);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            leap_one(t[j].x, r[j].x, g[j].x, a, DRIFT);
            leap_one(t[j].y, r[j].y, g[j].y, a, DRIFT);
            leap_one(t[j].z, r[j].z, g[j].z, a, DRIFT);
            leap_one(t[j].w, r[j].w, g[j].w, a, DRIFT);
            r4[i + j * stride] = r[j];
            if (DRIFT) t4[i + j * stride] = t[j];
            if (SNAP) s4[i + j * stride] = t[j];
        }
    }
    for (; i < nvec; i += stride) {
        float4 t = t4[i], r = r4[i];
        const float4 g = __ldg(g4 + i);
        leap_one(t.x, r.x, g.x, a, DRIFT);
        leap_one(t.y, r.y, g.y, a, DRIFT);
        leap_one(t.z, r.z, g.z, a, DRIFT);
        leap_one(t.w, r.w, g.w, a, DRIFT);
        r4[i] = r;
        if (DRIFT) t4[i] = t;
        if (SNAP) s4[i] = t;
    }
}

// ---- per-chain energy terms: partial sums per (slice, chain), then an ordered reduction --------------------------
constexpr int kEnergySlice = kHmcThreads * 4 * 8;      // 8192 elements per CTA

__global__ void __launch_bounds__(kHmcThreads) hmc_energy_partial_kernel(const float *__restrict__ theta,
                                                                          const float *__restrict__ r, int64_t D,
                                                                          int64_t ld, double *__restrict__ partial) {
    __shared__ double sh[2][kHmcThreads / 32];
    const int c = blockIdx.y, b = blockIdx.x;
    const float *tp = theta + (int64_t)c * ld, *rp = r + (int64_t)c * ld;
    const int64_t lo = (int64_t)b * kEnergySlice;
    const int64_t hi = lo + kEnergySlice < D ? lo + kEnergySlice : D;
    // fp32 products are exact in fp64; per-thread fp64 accumulation in index order, then a fixed tree
    double st = 0.0, sr = 0.0;
    const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
    for (int64_t e = lo + 4 * threadIdx.x; e < hi4; e += 4 * kHmcThreads) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(tp + e));
        const float4 q = __ldg(reinterpret_cast<const float4 *>(rp + e));
        st += (double)t.x * t.x + (double)t.y * t.y + (double)t.z * t.z + (double)t.w * t.w;
        sr += (double)q.x * q.x + (double)q.y * q.y + (double)q.z * q.z + (double)q.w * q.w;
    }
    if (threadIdx.x < hi - hi4) {
        const float t = tp[hi4 + threadIdx.x], q = rp[hi4 + threadIdx.x];
        st += (double)t * t;
        sr += (double)q * q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        st += __shfl_down_sync(0xffffffffu, st, o);
        sr += __shfl_down_sync(0xffffffffu, sr, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = st; sh[1][warp] = sr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int w = 0; w < kHmcThreads / 32; ++w) { a0 += sh[0][w]; a1 += sh[1][w]; }
        partial[((int64_t)c * gridDim.x + b) * 2 + 0] = a0;
        partial[((int64_t)c * gridDim.x + b) * 2 + 1] = a1;
    }
}

__global__ void hmc_energy_final_kernel(const double *__restrict__ partial, int nslices, int64_t C,
                                        double *__restrict__ sum_theta2, double *__restrict__ sum_r2) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a0 = 0.0, a1 = 0.0;
    for (int b = 0; b < nslices; ++b) {
        a0 += partial[(c * nslices + b) * 2 + 0];
        a1 += partial[(c * nslices + b) * 2 + 1];
    }
    sum_theta2[c] = a0;
    sum_r2[c] = a1;
}

// ---- accept / reject ----------------------------------------------------------------------------------------------
// One uniform per chain: Philox block (chain_offset + c) at `step`, first word -> u in (0, 1].
__global__ void hmc_accept_flags_kernel(const double *__restrict__ h_old, const double *__restrict__ h_new,
                                        const float *__restrict__ logu, int64_t C, uint2 key, uint64_t step,
                                        uint64_t chain_offset, int32_t *__restrict__ accept) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float lu;
    if (logu) {
        lu = logu[c];
    } else {
        const uint64_t blk = chain_offset + (uint64_t)c;
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
        lu = logf(fmaf(__uint2float_rn(rnd.x), 2.3283064365386963e-10f, 1.1641532182693481e-10f));
    }
    const double hn = h_new[c], ho = h_old[c];
    const double rho = fmin(0.0, ho - hn);                       // hamiltorch: rho = min(0., -new_ham + ham)
    // a non-finite energy is a rejection (hamiltorch raises LogProbError and re-appends the previous state)
    accept[c] = (isfinite(hn) && isfinite(ho) && rho >= (double)lu) ? 1 : 0;
}

__global__ void __launch_bounds__(kHmcThreads) hmc_commit_kernel(float *__restrict__ theta, float *__restrict__ saved,
                                                                  float *__restrict__ keep_dst, const float *__restrict__ keep_src,
                                                                  float *__restrict__ out, int64_t ld_out, int64_t ld,
                                                                  const int32_t *__restrict__ accept) {
    const int c = blockIdx.y;
    const bool acc = accept[c] != 0;
    const int64_t nvec = ld >> 2;
    float4 *t4 = reinterpret_cast<float4 *>(theta + (int64_t)c * ld);
    float4 *s4 = reinterpret_cast<float4 *>(saved + (int64_t)c * ld);
    float4 *kd = keep_dst ? reinterpret_cast<float4 *>(keep_dst + (int64_t)c * ld) : nullptr;
    const float4 *ks = keep_src ? reinterpret_cast<const float4 *>(keep_src + (int64_t)c * ld) : nullptr;
    float4 *o4 = out ? reinterpret_cast<float4 *>(out + (int64_t)c * ld_out) : nullptr;
    for (int64_t i = (int64_t)blockIdx.x * kHmcThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kHmcThreads) {
        float4 cur;
        if (acc) { cur = t4[i]; s4[i] = cur; } else { cur = s4[i]; t4[i] = cur; }
        if (kd) {
            if (acc) { cur = __ldg(ks + i); kd[i] = cur; } else { cur = kd[i]; }
        }
        if (o4) o4[i] = cur;
    }
}

}  // namespace ursa

using namespace ursa;

extern "C" int ursa_hmc_momentum(float *r, const float *noise, int64_t n, float sqrt_mass, uint64_t seed, uint64_t step,
                                 uint64_t elem_offset, void *stream) {
    URSA_REQUIRE(n >= 0 && (n & 3) == 0, "ursa_hmc_momentum: n must be a non-negative multiple of 4 (padded rows)");
    URSA_REQUIRE(n == 0 || r, "ursa_hmc_momentum: r is null");
    URSA_REQUIRE(aligned16(r) && aligned16(noise), "ursa_hmc_momentum: buffers must be 16-byte aligned");
    URSA_REQUIRE((elem_offset & 3u) == 0, "ursa_hmc_momentum: elem_offset must be a multiple of 4");
    URSA_REQUIRE(sqrt_mass > 0.f, "ursa_hmc_momentum: sqrt_mass must be positive");
    if (n == 0) return URSA_OK;
    const int grid = hmc_grid(n >> 2, 8);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    if (noise) hmc_momentum_kernel<false><<<grid, kHmcThreads, 0, (cudaStream_t)stream>>>(r, noise, n, sqrt_mass, key, step, elem_offset >> 2);
    else hmc_momentum_kernel<true><<<grid, kHmcThreads, 0, (cudaStream_t)stream>>>(r, noise, n, sqrt_mass, key, step, elem_offset >> 2);
    URSA_LAUNCH_CHECK("hmc_momentum_kernel");
    return URSA_OK;
}

extern "C" int ursa_hmc_leapfrog(float *theta, float *r, const float *g_nll, float *snapshot, int64_t n, float kick,
                                 float drift, float tau, float tau_out, void *stream) {
    URSA_REQUIRE(n >= 0 && (n & 3) == 0, "ursa_hmc_leapfrog: n must be a non-negative multiple of 4 (padded rows)");
    URSA_REQUIRE(n == 0 || (theta && r && g_nll), "ursa_hmc_leapfrog: theta, r and g_nll must be non-null");
    URSA_REQUIRE(aligned16(theta) && aligned16(r) && aligned16(g_nll) && aligned16(snapshot),
                 "ursa_hmc_leapfrog: buffers must be 16-byte aligned");
    URSA_REQUIRE(!(snapshot && drift == 0.f), "ursa_hmc_leapfrog: a snapshot needs a drift (position) update");
    if (n == 0) return URSA_OK;
    LeapArgs a;
    a.theta = theta; a.r = r; a.g = g_nll; a.snap = snapshot; a.n = n;
    a.kick = kick; a.drift = drift; a.tau = tau; a.tau_out = tau_out;
    const int grid = hmc_grid(((n >> 2) + 1) / 2, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (drift == 0.f) hmc_leapfrog_kernel<false, false><<<grid, kHmcThreads, 0, st>>>(a);
    else if (snapshot) hmc_leapfrog_kernel<true, true><<<grid, kHmcThreads, 0, st>>>(a);
    else hmc_leapfrog_kernel<true, false><<<grid, kHmcThreads, 0, st>>>(a);
    URSA_LAUNCH_CHECK("hmc_leapfrog_kernel");
    return URSA_OK;
}

static int energy_slices(int64_t D) { return (int)((D + kEnergySlice - 1) / kEnergySlice); }

extern "C" size_t ursa_hmc_energy_workspace(int64_t C, int64_t D) {
    if (C <= 0 || D <= 0) return 0;
    return (size_t)C * energy_slices(D) * 2 * sizeof(double);
}

extern "C" int ursa_hmc_energy(const float *theta, const float *r, int64_t C, int64_t D, int64_t ld, double *sum_theta2,
                               double *sum_r2, void *workspace, size_t workspace_bytes, void *stream) {
    URSA_REQUIRE(C >= 0 && D >= 0 && ld >= D && (ld & 3) == 0, "ursa_hmc_energy: bad sizes (ld >= D, ld %% 4 == 0)");
    if (C == 0) return URSA_OK;
    URSA_REQUIRE(theta && r && sum_theta2 && sum_r2, "ursa_hmc_energy: null pointer");
    URSA_REQUIRE(aligned16(theta) && aligned16(r), "ursa_hmc_energy: buffers must be 16-byte aligned");
    URSA_REQUIRE(C <= 65535, "ursa_hmc_energy: at most 65535 chains per call");
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 0) {
        URSA_CUDA(cudaMemsetAsync(sum_theta2, 0, C * sizeof(double), st));
        URSA_CUDA(cudaMemsetAsync(sum_r2, 0, C * sizeof(double), st));
        return URSA_OK;
    }
    const int ns = energy_slices(D);
    URSA_REQUIRE(workspace && workspace_bytes >= ursa_hmc_energy_workspace(C, D) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "ursa_hmc_energy: workspace too small or misaligned");
    double *partial = reinterpret_cast<double *>(workspace);
    hmc_energy_partial_kernel<<<dim3(ns, (unsigned)C), kHmcThreads, 0, st>>>(theta, r, D, ld, partial);
    URSA_LAUNCH_CHECK("hmc_energy_partial_kernel");
    hmc_energy_final_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(partial, ns, C, sum_theta2, sum_r2);
    URSA_LAUNCH_CHECK("hmc_energy_final_kernel");
    return URSA_OK;
}

extern "C" int ursa_hmc_accept(float *theta, float *saved, float *keep_dst, const float *keep_src, float *out,
                               int64_t ld_out, int64_t C, int64_t ld, const double *h_old, const double *h_new,
                               const float *logu, int32_t *accept, uint64_t seed, uint64_t step, uint64_t chain_offset,
                               void *stream) {
    URSA_REQUIRE(C >= 0 && ld >= 0 && (ld & 3) == 0, "ursa_hmc_accept: bad sizes (ld %% 4 == 0)");
    if (C == 0) return URSA_OK;
    URSA_REQUIRE(theta && saved && h_old && h_new && accept, "ursa_hmc_accept: null pointer");
    URSA_REQUIRE((keep_dst == nullptr) == (keep_src == nullptr), "ursa_hmc_accept: keep_dst and keep_src go together");
    URSA_REQUIRE(!out || (ld_out >= ld && (ld_out & 3) == 0), "ursa_hmc_accept: bad ld_out");
    URSA_REQUIRE(aligned16(theta) && aligned16(saved) && aligned16(keep_dst) && aligned16(keep_src) && aligned16(out),
                 "ursa_hmc_accept: buffers must be 16-byte aligned");
    URSA_REQUIRE(C <= 65535, "ursa_hmc_accept: at most 65535 chains per call");
    cudaStream_t st = (cudaStream_t)stream;
    hmc_accept_flags_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(
        h_old, h_new, logu, C, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), step, chain_offset, accept);
    URSA_LAUNCH_CHECK("hmc_accept_flags_kernel");
    if (ld == 0) return URSA_OK;
    int bx = (int)(((ld >> 2) + kHmcThreads - 1) / kHmcThreads);
    const int cap = (sm_count() * 8 + (int)C - 1) / (int)C;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    hmc_commit_kernel<<<dim3(bx, (unsigned)C), kHmcThreads, 0, st>>>(theta, saved, keep_dst, keep_src, out, ld_out, ld, accept);
    URSA_LAUNCH_CHECK("hmc_commit_kernel");
    return URSA_OK;
}
