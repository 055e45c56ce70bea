// Probe 3 (development aid): kind::f16 MMAs on NO-SWIZZLE K-major operands (16-byte chunks of 8 halves, K = 16 = 2 chunks,
// LBO = plane stride, SBO = 128 B): correctness of row-shifted A windows and the issue cadence at N = 16..128, plus the
// [N = 2C ; N = C] pair the fused PreResNet stage kernel issues per K step.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace ursa { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return -2; } int sm_count() { return 148; } }
using namespace ursa;

constexpr int ROWS = 640, NB = 128;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_f16_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

// a: [ROWS][16] halves, b: [NB][16] halves (row = output column n).  pair != 0: alternate an N = n MMA with an N = n/2 MMA.
__global__ void __launch_bounds__(128) probe(const __half *a, const __half *b, float *d, int n, int shift, int pair, int reps,
                                             long long *cycles, int noise = 0) {
    __shared__ volatile int stop_flag;
    __shared__ float sink[128];
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3[8];
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *g = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_off = 4 * ROWS * 16;      // two A plane sets (second = copy of the first) then B
    for (int i = threadIdx.x; i < ROWS * 16; i += blockDim.x) {
        const int r = i / 16, k = i % 16;
        *reinterpret_cast<__half *>(g + (k / 8) * (ROWS * 16) + r * 16 + (k % 8) * 2) = a[i];
        *reinterpret_cast<__half *>(g + 2 * ROWS * 16 + (k / 8) * (ROWS * 16) + r * 16 + (k % 8) * 2) = a[i];
    }
    for (int i = threadIdx.x; i < NB * 16; i += blockDim.x) {
        const int r = i / 16, k = i % 16;
        for (int sl = 0; sl < 9; ++sl)
            *reinterpret_cast<__half *>(g + b_off + sl * (NB * 16 * 2) + (k / 8) * (NB * 16) + r * 16 + (k % 8) * 2) = b[i];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 96); for (int i = 0; i < 8; ++i) mbar_init(&bar3[i], 1); fence_barrier_init(); stop_flag = 0; }
    fence_proxy_async();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0 && elect_one()) {
        const uint64_t da = desc_noswz(base + shift * 16, ROWS * 16, 128);
        const uint64_t db = desc_noswz(base + b_off, NB * 16, 128);
        // pair: 1 = same A, disjoint D; 2 = second A plane set, D overlapping the first MMA's upper half (the stage kernel);
        //       3 = second A plane set, disjoint D; 4 = same A, overlapping D; 5 = as 2 with 9 rotating window shifts
        const uint64_t da2 = (pair == 2 || pair == 3 || pair >= 5) ? desc_noswz(base + 2 * ROWS * 16 + shift * 16, ROWS * 16, 128) : da;
        const uint32_t d2 = (pair == 2 || pair == 4 || pair >= 5) ? tmem + n / 2 : tmem + 128;
        const uint32_t idesc = make_f16_idesc(128, n), idesc2 = make_f16_idesc(128, n / 2 < 16 ? 16 : n / 2);
        const long long t0 = clock64();
        if (pair == 6 || pair == 7 || pair == 8) {
            // the stage kernel's issue pattern: 9 taps unrolled (constant window shifts, 9 B slots), hi / lo' plane sets,
            // D = [ACC | LO] then LO; one "tile" per iteration, tiles advance by 128 rows (3 tiles, rotating)
            constexpr uint32_t SH[9] = {0, 1, 2, 33, 34, 35, 66, 67, 68};
            const uint64_t slotw = (uint64_t)((NB * 16 * 2) >> 4);
            for (int i = 0; i < reps; ++i) {
                const uint64_t tw = (uint64_t)((i % 3) * 128);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    umma_f16(tmem, da + tw + SH[tap], db + slotw * tap, idesc, tap > 0);
                    umma_f16(d2, da2 + tw + SH[tap], db + slotw * tap, idesc2, 1);
                }
                if (pair >= 7) umma_commit(smem_u32(&bar3[i & 7]));      // per-"tile" commit, nobody waits on it
                if (pair == 8) (void)mbar_try_wait_a(smem_u32(&bar3[(i + 3) & 7]), 0);   // plus one barrier poll per tile
            }
        } else if (pair == 5) {
            const int sh[9] = {0, 1, 2, 33, 34, 35, 66, 67, 68};
            for (int i = 0; i < reps; ++i) {
                const uint64_t o = (uint64_t)sh[i % 9];
                umma_f16(tmem, da + o, db, idesc, i > 0);
                umma_f16(d2, da2 + o, db, idesc2, 1);
            }
        } else
        for (int i = 0; i < reps; ++i) {
            umma_f16(tmem, da, db, idesc, i > 0);
            if (pair) umma_f16(d2, da2, db, idesc2, pair == 1 || pair == 3 ? i > 0 : 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait_a(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (cycles) *cycles = t1 - t0;
        stop_flag = 1;
    } else if (warp > 0 && noise) {
        // concurrent "epilogue" traffic while warp 0 issues MMAs: 1 = tcgen05.ld from other TMEM columns, 2 = st.shared.v4
        // into an unrelated shared-memory region, 3 = ld.shared.v4
        float acc = 0.f;
        unsigned char *scratch = g + 100 * 1024 + threadIdx.x * 16;
        while (!stop_flag) {
            if (noise == 1) {
                uint32_t rr[16];
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 192, rr);
                acc += __uint_as_float(rr[0]);
            } else if (noise == 2) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(scratch + u * 2048)), "f"(acc) : "memory");
                acc += 1.f;
            } else if (noise == 4) {
                asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(scratch)), "f"(acc) : "memory");
                fence_proxy_async();
                acc += 1.f;
            } else if (noise == 5) {
                asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(scratch)), "f"(acc) : "memory");
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bar2);
                acc += 1.f;
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(smem_u32(scratch + u * 2048)) : "memory");
                    acc += x0;
                }
            }
        }
        sink[threadIdx.x] = acc;
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
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
    std::vector<__half> a(ROWS * 16), b(NB * 16);
    std::vector<float> af(ROWS * 16), bf(NB * 16);
    srand(1);
    for (size_t i = 0; i < a.size(); ++i) { af[i] = (rand() % 2001 - 1000) / 1024.f; a[i] = __float2half(af[i]); af[i] = __half2float(a[i]); }
    for (size_t i = 0; i < b.size(); ++i) { bf[i] = (rand() % 2001 - 1000) / 1024.f; b[i] = __float2half(bf[i]); bf[i] = __half2float(b[i]); }
    __half *da, *db; float *dd; long long *dc;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 128 * NB * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 160 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int noise = 0;
    auto run = [&](int n, int shift, int pair, int reps, bool check) {
        cudaMemset(dd, 0, 128 * NB * 4);
        probe<<<1, 128, smem>>>(da, db, dd, n, shift, pair, reps, dc, noise);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("n %d shift %d: CUDA error %s\n", n, shift, cudaGetErrorString(e)); exit(1); }
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        double maxerr = -1;
        if (check) {
            std::vector<float> d(128 * n);
            cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
            maxerr = 0;
            for (int r = 0; r < 128; ++r)
                for (int cc = 0; cc < n; ++cc) {
                    double ref = 0;
                    for (int k = 0; k < 16; ++k) ref += (double)af[(r + shift) * 16 + k] * bf[cc * 16 + k];
                    maxerr = fmax(maxerr, fabs(ref * reps - d[r * n + cc]));
                }
        }
        printf("f16 no-swizzle n=%d shift=%d pair=%d noise=%d reps=%d: %.2f cyc/iter  %s\n", n, shift, pair, noise, reps, (double)c / reps,
               check ? (maxerr < 1e-3 ? "MATCH" : "mismatch") : "");
    };
    for (int shift : {0, 1, 3, 8, 33, 35}) run(32, shift, 0, 1, true);
    run(128, 5, 0, 1, true);
    for (int n : {16, 32, 64, 128}) run(n, 3, 0, 512, false);
    for (int pair : {1, 2, 3, 4, 5})
        for (int n : {32, 64, 128}) run(n, 3, pair, 512, false);
    for (int n : {32, 64, 128}) run(n, 0, 2, 512, false);          // aligned window
    for (noise = 0; noise <= 5; ++noise)
        for (int n : {32, 64, 128}) run(n, 0, 6, 90, false);           // cyc/iter = 9 pairs
    noise = 0;
    for (int pair : {7, 8})
        for (int n : {32, 64, 128}) run(n, 0, pair, 90, false);
    return 0;
}
