// Shared device/host helpers for the ursa_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ursa_b200.h"

namespace ursa {

// ---- error plumbing (thread-local message; see include/ursa_b200.h) -------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
int sm_count();
void count_launch();                       // process-wide count of kernel launches issued by this library (ursa_launch_count)

#define URSA_REQUIRE(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            ::ursa::set_error(__VA_ARGS__);    \
            return URSA_ERR_INVALID;           \
        }                                      \
    } while (0)

#define URSA_CUDA(call)                                          \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return ::ursa::cuda_fail(e__, #call); \
    } while (0)

#define URSA_LAUNCH_CHECK(name)                                  \
    do {                                                         \
        ::ursa::count_launch();                                  \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return ::ursa::cuda_fail(e__, name); \
    } while (0)

// ---- optional per-kernel CUDA-event timing (ursa_profile_begin / ursa_profile_end): a diagnostic for benchmarks ----
// ProfScope records an event pair on `st` around the launches in its scope when profiling is on; otherwise it is two
// relaxed loads.  Only the PreResNet BMA forward is instrumented (the kernels bench.py's roofline names).
bool prof_on();
void prof_mark(int kind, cudaStream_t st, bool begin);
struct ProfScope {
    int kind;
    cudaStream_t st;
    bool on;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s), on(prof_on()) { if (on) prof_mark(kind, st, true); }
    ~ProfScope() { if (on) prof_mark(kind, st, false); }
};

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based ---------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1,
                       (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// lg2.approx.ftz: MUFU.LG2 alone -- __log2f adds a denormal rescue (compare, scale, fix-up) around it; the arguments here are
// >= 2^-33, where both give the same bits
__device__ __forceinline__ float fast_log2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Box-Muller on one Philox block: (r.x, r.y) -> (z0, z1), (r.z, r.w) -> (z2, z3).
// u1 = r*2^-32 + 2^-33 in (0, 1]; theta = 2*pi*u2.  MUFU lg2 / sqrt / sin / cos.
__device__ __forceinline__ float4 box_muller4(uint4 r) {
    const float k2m32 = 2.3283064365386963e-10f, k2m33 = 1.1641532182693481e-10f;
    const float kTwoPi2m32 = 1.4629180792671596e-9f;          // 2*pi * 2^-32
    const float u1a = fmaf(__uint2float_rn(r.x), k2m32, k2m33);
    const float u1b = fmaf(__uint2float_rn(r.z), k2m32, k2m33);
    const float ra = fast_sqrt(-1.3862943611198906f * fast_log2(u1a));   // sqrt(-2 ln u)
    const float rb = fast_sqrt(-1.3862943611198906f * fast_log2(u1b));
    const float ta = fmaf(__uint2float_rn(r.y), kTwoPi2m32, 0.5f * kTwoPi2m32);
    const float tb = fmaf(__uint2float_rn(r.w), kTwoPi2m32, 0.5f * kTwoPi2m32);
    float sa, ca, sb, cb;
    __sincosf(ta, &sa, &ca);
    __sincosf(tb, &sb, &cb);
    return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

// The 4 normals of Philox block `blk` (= element index / 4) at stream position `step`.
__device__ __forceinline__ float4 philox_normal4(uint64_t blk, uint64_t step, uint2 key) {
    const uint4 ctr = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)step, (uint32_t)(step >> 32));
    return box_muller4(philox4x32_10(ctr, key));
}

// The SWAG draw's stream (K2b): SIX normals per Philox block.  The 32-bit generator work (20 IMAD.WIDE + 20 LOP3, four and two
// issue cycles each on this SM) is what bounds that kernel, and a float carries 24 bits: the block's 128 bits are cut, from the
// top of x to the bottom of w, into three (24-bit radius uniform, 18-bit angle) pairs (2 bits unused):
//   a0 = x[31:8]              t0 = x[7:0] : y[31:22]
//   a1 = y[21:0] : z[31:30]   t1 = z[29:12]
//   a2 = z[11:0] : w[31:20]   t2 = w[19:2]
//   u = (a + 1/2) 2^-24 (rounded to fp32, in (0, 1]),  theta = 2 pi (t + 1/2) 2^-18,  z[2j] = r cos(theta), z[2j+1] = r sin(theta),
//   r = sqrt(-2 ln u) <= 5.89 (torch's CPU float normal has the same 24-bit radius resolution).
__device__ __forceinline__ void box_muller6(uint4 r, float (&z)[6]) {
    const uint32_t a[3] = {r.x >> 8, __funnelshift_l(r.z, r.y, 2) & 0xFFFFFFu, __funnelshift_l(r.w, r.z, 12) & 0xFFFFFFu};
    const uint32_t t[3] = {__funnelshift_l(r.y, r.x, 10) & 0x3FFFFu, (r.z >> 12) & 0x3FFFFu, (r.w >> 2) & 0x3FFFFu};
    const float k2m24 = 5.9604644775390625e-08f, k2m25 = 2.98023223876953125e-08f;
    const float kTwoPi2m18 = 2.3968449810247229e-05f;         // 2*pi * 2^-18
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float u = fmaf(__uint2float_rn(a[j]), k2m24, k2m25);
        const float rad = fast_sqrt(-1.3862943611198906f * fast_log2(u));
        const float th = fmaf(__uint2float_rn(t[j]), kTwoPi2m18, 0.5f * kTwoPi2m18);
        float sn, cs;
        __sincosf(th, &sn, &cs);
        z[2 * j] = rad * cs;
        z[2 * j + 1] = rad * sn;
    }
}

__device__ __forceinline__ void philox_normal6(uint64_t blk, uint64_t step, uint2 key, float (&z)[6]) {
    const uint4 ctr = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)step, (uint32_t)(step >> 32));
    box_muller6(philox4x32_10(ctr, key), z);
}

__device__ __forceinline__ float f4_get(const float4 &v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

}  // namespace ursa
