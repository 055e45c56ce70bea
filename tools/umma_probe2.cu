// Probe 2 (development aid): (a) cadence of NO-SWIZZLE K-major MMAs when the A window is not 128-byte aligned,
// (b) row-shifted windows in SWIZZLE_128B / SWIZZLE_64B K-major layouts (start address + s * row pitch, with and without
// the descriptor's base-offset field), correctness and cadence.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "tc_common.cuh"

namespace ursa { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return -2; } int sm_count() { return 148; } }
using namespace ursa;

constexpr int ROWS = 640;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// swizzled K-major: row pitch = swizzle span (128 or 64 B), SBO = 8 rows
__device__ __forceinline__ uint64_t desc_swz(uint32_t addr, int row_bytes, uint32_t base_off) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_off & 7) << 49) | ((uint64_t)(row_bytes == 128 ? 2 : 4) << 61);
}

// mode 0: no swizzle planes (K=8: 2 planes); mode 1: SW128 (32 floats per row); mode 2: SW64 (16 floats per row)
__global__ void __launch_bounds__(128) probe(const float *a, const float *b, float *d, int n, int shift, int mode, int bo_mode,
                                             int kstep, int reps, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *g = smem_raw + (base - smem_u32(smem_raw));
    const int rowf = mode == 1 ? 32 : (mode == 2 ? 16 : 8);      // floats per logical row in the source arrays (a: [ROWS][32])
    const int rb = rowf * 4;
    const uint32_t a_bytes = mode == 0 ? 2 * ROWS * 16 : ROWS * rb;
    const uint32_t b_off = (a_bytes + 1023u) & ~1023u;
    // a is [ROWS][32] row-major, b is [64][32]; K window used = floats [8*kstep, 8*kstep+8)
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) {
        const int r = i / 32, k = i % 32;
        if (mode == 0) { if (k < 8) *reinterpret_cast<float *>(g + (k / 4) * (ROWS * 16) + r * 16 + (k % 4) * 4) = a[i]; }
        else if (k < rowf) {
            uint32_t off = r * rb + k * 4;
            const uint32_t abs = base + off;
            const uint32_t sw = mode == 1 ? ((abs >> 7) & 7) << 4 : ((abs >> 7) & 3) << 4;
            *reinterpret_cast<float *>(g + ((abs ^ sw) - base)) = a[i];
        }
    }
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int r = i / 32, k = i % 32;
        if (mode == 0) { if (k < 8) *reinterpret_cast<float *>(g + b_off + (k / 4) * (64 * 16) + r * 16 + (k % 4) * 4) = b[i]; }
        else if (k < rowf) {
            const uint32_t abs = base + b_off + r * rb + k * 4;
            const uint32_t sw = mode == 1 ? ((abs >> 7) & 7) << 4 : ((abs >> 7) & 3) << 4;
            *reinterpret_cast<float *>(g + ((abs ^ sw) - base)) = b[i];
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0 && elect_one()) {
        uint64_t da, db;
        if (mode == 0) {
            da = desc_noswz(base + shift * 16, ROWS * 16, 128);
            db = desc_noswz(base + b_off, 64 * 16, 128);
        } else {
            const uint32_t astart = base + shift * rb + kstep * 32;
            const uint32_t bo = bo_mode == 0 ? 0 : (mode == 1 ? (astart >> 7) & 7 : (astart >> 7) & 3);
            da = desc_swz(astart, rb, bo);
            db = desc_swz(base + b_off + kstep * 32, rb, 0);
        }
        const uint32_t idesc = make_tf32_idesc(128, n);
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) umma_tf32(tmem, da, db, idesc, i > 0);
        umma_commit(smem_u32(&bar));
        mbar_wait_a(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (cycles) *cycles = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < n; c0 += 16) {
        uint32_t rr[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, rr);
        for (int i = 0; i < 16; ++i) d[(size_t)threadIdx.x * n + c0 + i] = __uint_as_float(rr[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
    std::vector<float> a(ROWS * 32), b(64 * 32);
    srand(1);
    for (auto &v : a) v = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : b) v = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    float *da, *db, *dd; long long *dc;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, 128 * 64 * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 100 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto run = [&](int n, int shift, int mode, int bo, int kstep, int reps, bool check) {
        cudaMemset(dd, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(da, db, dd, n, shift, mode, bo, kstep, reps, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d shift %d bo %d: CUDA error %s\n", mode, shift, bo, cudaGetErrorString(e)); exit(1); }
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        double maxerr = -1;
        if (check) {
            std::vector<float> d(128 * n);
            cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
            maxerr = 0;
            for (int r = 0; r < 128; ++r)
                for (int cc = 0; cc < n; ++cc) {
                    double ref = 0;
                    for (int k = 0; k < 8; ++k) ref += (double)a[(r + shift) * 32 + 8 * kstep + k] * b[cc * 32 + 8 * kstep + k];
                    maxerr = fmax(maxerr, fabs(ref * reps - d[r * n + cc]));
                }
        }
        printf("mode %d (%s) n=%d shift=%d base_off=%s kstep=%d reps=%d: %.2f cyc/MMA  %s\n", mode,
               mode == 0 ? "no-swizzle" : (mode == 1 ? "SW128" : "SW64"), n, shift, bo ? "formula" : "0", kstep, reps, (double)c / reps,
               check ? (maxerr < 1e-4 ? "MATCH" : "mismatch") : "");
    };
    for (int shift : {0, 1, 2, 4, 8, 33}) run(16, shift, 0, 0, 0, 512, false);     // (a) misaligned no-swizzle cadence
    for (int shift : {0, 8, 1, 3, 33, 37})
        for (int bo : {1, 0})
            for (int kstep : {0, 1, 3}) run(32, shift, 1, bo, kstep, 1, true);       // (b) SW128 shifted windows
    for (int shift : {0, 8, 1, 3, 33})
        for (int bo : {1, 0})
            for (int kstep : {0, 1}) run(16, shift, 2, bo, kstep, 1, true);          // SW64
    for (int mode : {1, 2})
        for (int shift : {0, 1, 3, 33}) run(16, shift, mode, 1, 0, 512, false);      // cadence of shifted swizzled windows
    for (int n : {32, 64, 128}) run(n, 3, 1, 1, 0, 512, false);
    return 0;
}
