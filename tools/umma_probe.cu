// Probe (development aid, not part of the library): semantics of the NO-SWIZZLE K-major shared-memory descriptor for
// tcgen05.mma kind::tf32 and the issue cadence of small-N MMAs.  Layout under test ("planes"):
//   element (row r, k) lives at  base + (k/4)*PLANE + r*16 + (k%4)*4   -> 8-row core matrices are contiguous (SBO = 128 B)
//   and the two 16-byte K chunks of one UMMA_K = 8 step are PLANE bytes apart (LBO = PLANE).
// A shifted window (rows r+s) is then just base + s*16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ursabench_b200/csrc tools/umma_probe.cu -o gpurun_out/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "tc_common.cuh"

namespace ursa { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return -2; } int sm_count() { return 148; } }
using namespace ursa;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;   // layout type 0 = no swizzle
}

constexpr int ROWS = 512;                 // rows available in the A plane (for shifted windows)
constexpr int PLANE_A = ROWS * 16;        // bytes
constexpr int NB = 64;
constexpr int PLANE_B = NB * 16;

// variant 0: LBO = plane stride, SBO = 128 ; variant 1: swapped
__global__ void __launch_bounds__(128) probe_kernel(const float *a, const float *b, float *d, int n, int shift, int variant,
                                                    int reps, long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float *sa = reinterpret_cast<float *>(smem);                      // 2 planes
    float *sb = reinterpret_cast<float *>(smem + 2 * PLANE_A);        // 2 planes
    for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) {        // a is [ROWS][8] row-major
        const int r = i / 8, k = i % 8;
        sa[(k / 4) * (PLANE_A / 4) + r * 4 + (k % 4)] = a[i];
    }
    for (int i = threadIdx.x; i < NB * 8; i += blockDim.x) {
        const int r = i / 8, k = i % 8;
        sb[(k / 4) * (PLANE_B / 4) + r * 4 + (k % 4)] = b[i];
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
        const uint32_t abase = smem_u32(sa) + shift * 16, bbase = smem_u32(sb);
        const uint64_t da = variant == 0 ? make_desc(abase, PLANE_A, 128) : make_desc(abase, 128, PLANE_A);
        const uint64_t db = variant == 0 ? make_desc(bbase, PLANE_B, 128) : make_desc(bbase, 128, PLANE_B);
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
    std::vector<float> a(ROWS * 8), b(NB * 8);
    srand(1);
    for (auto &v : a) v = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : b) v = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    float *da, *db, *dd; long long *dc;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, 128 * 64 * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 2 * PLANE_A + 2 * PLANE_B;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // cadence: reps back-to-back MMAs into the same accumulator (variant 0)
    for (int n : {16, 32, 64})
        for (int reps : {64, 1024}) {
            probe_kernel<<<1, 128, smem>>>(da, db, dd, n, 0, 0, reps, dc);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            printf("cadence n=%d reps=%d: %lld cycles total, %.2f cycles/MMA (floor N/2 = %d)\n", n, reps, c, (double)c / reps, n / 2);
        }
    for (int variant = 0; variant < 2; ++variant)
        for (int n : {16, 32, 64})
            for (int shift : {0, 1, 33, 37}) {
                cudaMemset(dd, 0, 128 * 64 * 4);
                probe_kernel<<<1, 128, smem>>>(da, db, dd, n, shift, variant, 1, nullptr);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("variant %d n %d shift %d: CUDA error %s\n", variant, n, shift, cudaGetErrorString(e)); return 1; }
                std::vector<float> d(128 * n);
                cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0;
                for (int r = 0; r < 128; ++r)
                    for (int c = 0; c < n; ++c) {
                        double ref = 0;
                        for (int k = 0; k < 8; ++k) ref += (double)a[(r + shift) * 8 + k] * b[c * 8 + k];
                        maxerr = fmax(maxerr, fabs(ref - d[r * n + c]));
                    }
                printf("variant %d (%s) n=%d shift=%d max|err| = %.3g %s\n", variant, variant == 0 ? "LBO=plane,SBO=128" : "LBO=128,SBO=plane",
                       n, shift, maxerr, maxerr < 1e-5 ? "MATCH" : "mismatch");
            }
    return 0;
}
