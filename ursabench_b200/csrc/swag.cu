// K2: SWAG moment / deviation-ring update, variance, and the batched rank-K draw.
//  collect   reference inference/swa.py:79-90 + inference/subspaces.py:85-89     24 B/param, one pass
//  variance  reference inference/swa.py:106-108                                 12 B/param
//  draw      reference inference/swag.py:85-97 (formula; see header)            (K + 2 + S) * 4 B/param
//
// The draw stages [K x 1024-column] tiles of the deviation ring in shared memory with the TMA engine
// (cp.async.bulk + mbarrier, double buffered), keeps z2 [K, S] in shared memory (broadcast LDS.128) and all
// S accumulators of a thread's 4 columns in registers, so the ring is read from HBM exactly once for
// all S draws.  z1 comes from Philox in-register (or from memory in parity mode).
#include "async.cuh"
#include "common.cuh"

namespace ursa {

constexpr int kEwThreads = 256;

__global__ void __launch_bounds__(kEwThreads) swag_collect_kernel(const float *__restrict__ w, float *__restrict__ mean,
                                                                   float *__restrict__ sq, float *__restrict__ dev,
                                                                   int64_t n, float keep, float denom) {
    // element order of operations = reference: mul_(keep) ; add_(w/denom) ; pow(2)/denom ; w - mean
    auto one = [&](float wv, float &m, float &s, float &d) {
        m = __fadd_rn(__fmul_rn(m, keep), __fdiv_rn(wv, denom));                    // swa.py:83-84
        s = __fadd_rn(__fmul_rn(s, keep), __fdiv_rn(__fmul_rn(wv, wv), denom));     // swa.py:87-88
        d = __fsub_rn(wv, m);                                                       // swa.py:89
    };
    const int64_t nvec = n >> 2;
    const float4 *w4 = reinterpret_cast<const float4 *>(w);
    float4 *m4 = reinterpret_cast<float4 *>(mean), *s4 = reinterpret_cast<float4 *>(sq),
           *d4 = reinterpret_cast<float4 *>(dev);
    const int64_t stride = (int64_t)gridDim.x * kEwThreads;
    for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < nvec; i += stride) {
        const float4 wv = __ldg(w4 + i);
        float4 m = m4[i], s = s4[i], d;
        one(wv.x, m.x, s.x, d.x);
        one(wv.y, m.y, s.y, d.y);
        one(wv.z, m.z, s.z, d.z);
        one(wv.w, m.w, s.w, d.w);
        m4[i] = m;
        s4[i] = s;
        d4[i] = d;
    }
    const int tail = (int)(n & 3);
    if (blockIdx.x == 0 && threadIdx.x < tail) {
        const int64_t e = (nvec << 2) + threadIdx.x;
        float m = mean[e], s = sq[e], d;
        one(w[e], m, s, d);
        mean[e] = m;
        sq[e] = s;
        dev[e] = d;
    }
}

__global__ void __launch_bounds__(kEwThreads) swag_variance_kernel(const float *__restrict__ mean,
                                                                    const float *__restrict__ sq,
                                                                    float *__restrict__ var, int64_t n, float clamp) {
    const int64_t stride = (int64_t)gridDim.x * kEwThreads;
    for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < n; i += stride) {
        const float m = mean[i];
        var[i] = fmaxf(__fsub_rn(sq[i], __fmul_rn(m, m)), clamp);                   // swa.py:107
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int kDrawThreads = 256;
constexpr int kTileCols = kDrawThreads * 4;     // 1024 columns = 4 KB per ring row per stage
constexpr int kStages = 2;

struct DrawArgs {
    float *out;
    const float *mean, *var, *ring, *z2, *z1;
    int64_t ld_out, ld_ring, ld_z1, D;
    int K, S;
    float rank_div;
    uint2 key;
    uint64_t step;
};

// SP = padded number of draws, NG = sample groups: thread (tid % 256) owns 4 columns, group (tid / 256) owns the
// draws [g*SP/NG, (g+1)*SP/NG): more resident warps and fewer accumulator registers per thread for large S.
template <int SP, int NG>
__global__ void __launch_bounds__(kDrawThreads * NG, 1) swag_draw_kernel(const DrawArgs a) {
    constexpr int SPG = SP / NG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring_s = reinterpret_cast<float *>(smem_raw);                       // [kStages][K][kTileCols]
    __shared__ __align__(16) float z2s[URSA_DRAW_MAX_K * SP];                  // [K][SP], zero padded
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int tid = threadIdx.x % kDrawThreads;          // column owner
    const int grp = threadIdx.x / kDrawThreads;          // sample group (warp-uniform)
    const int s_lo = grp * SPG;
    const int K = a.K, S = a.S;
    const float inv_div = (K > 0) ? 1.0f / a.rank_div : 0.f;
    for (int i = threadIdx.x; i < K * SP; i += kDrawThreads * NG) {
        const int k = i / SP, s = i - k * SP;
        z2s[i] = (s < S) ? a.z2[(int64_t)s * K + k] * inv_div : 0.f;      // 1/sqrt(max_rank-1) folded in (swag.py:95)
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t ntiles = (a.D + kTileCols - 1) / kTileCols;
    const int64_t Dp4 = (a.D + 3) >> 2;                                        // Philox blocks per draw row
    auto issue = [&](int64_t tile, int stage) {                                 // one thread: K bulk copies
        const int64_t c0 = tile * kTileCols;
        int64_t rem = a.D - c0;
        const uint32_t cols = (uint32_t)(rem >= kTileCols ? kTileCols : ((rem + 3) & ~(int64_t)3));
        const uint32_t bytes = cols * 4u;
        mbar_arrive_expect_tx(&full_bar[stage], bytes * (uint32_t)K);
        float *dst = ring_s + (size_t)stage * K * kTileCols;
        for (int k = 0; k < K; ++k)
            bulk_g2s(dst + (size_t)k * kTileCols, a.ring + (int64_t)k * a.ld_ring + c0, bytes, &full_bar[stage]);
    };

    if (K > 0 && threadIdx.x == 0 && (int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        const uint32_t parity = (uint32_t)(it >> 1) & 1u;
        const int64_t next = tile + gridDim.x;
        if (K > 0 && threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);

        float acc[SPG][4];
#pragma unroll
        for (int s = 0; s < SPG; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;

        if (K > 0) {
            mbar_wait(&full_bar[stage], parity);
            const float4 *rs = reinterpret_cast<const float4 *>(ring_s + (size_t)stage * K * kTileCols) + tid;
            for (int k = 0; k < K; ++k) {
                const float4 r = rs[(size_t)k * (kTileCols / 4)];
                const float4 *zk = reinterpret_cast<const float4 *>(z2s + k * SP + s_lo);
#pragma unroll
                for (int q = 0; q < SPG / 4; ++q) {
                    const float4 z = zk[q];
                    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[4 * q + j][0] = fmaf(zz[j], r.x, acc[4 * q + j][0]);
                        acc[4 * q + j][1] = fmaf(zz[j], r.y, acc[4 * q + j][1]);
                        acc[4 * q + j][2] = fmaf(zz[j], r.z, acc[4 * q + j][2]);
                        acc[4 * q + j][3] = fmaf(zz[j], r.w, acc[4 * q + j][3]);
                    }
                }
            }
        }

        const int64_t c0 = tile * kTileCols + (int64_t)tid * 4;
        if (c0 < a.D) {
            const bool full4 = c0 + 4 <= a.D;
            float m[4], sd[4];
            if (full4) {
                const float4 mv = __ldg(reinterpret_cast<const float4 *>(a.mean + c0));
                const float4 vv = __ldg(reinterpret_cast<const float4 *>(a.var + c0));
                m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
                sd[0] = sqrtf(vv.x); sd[1] = sqrtf(vv.y); sd[2] = sqrtf(vv.z); sd[3] = sqrtf(vv.w);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool ok = c0 + j < a.D;
                    m[j] = ok ? a.mean[c0 + j] : 0.f;
                    sd[j] = ok ? sqrtf(a.var[c0 + j]) : 0.f;
                }
            }
#pragma unroll
            for (int sl = 0; sl < SPG; ++sl) {
                const int s = s_lo + sl;
                if (s < S) {
                    float z[4];
                    if (a.z1) {
                        const float *zr = a.z1 + (int64_t)s * a.ld_z1 + c0;
                        if (full4) {
                            const float4 zv = __ldg(reinterpret_cast<const float4 *>(zr));
                            z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) z[j] = (c0 + j < a.D) ? zr[j] : 0.f;
                        }
                    } else {
                        const float4 zv = philox_normal4((uint64_t)s * (uint64_t)Dp4 + (uint64_t)(c0 >> 2), a.step, a.key);
                        z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w;
                    }
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float r = __fmul_rn(sd[j], z[j]);                                  // swag.py:88-89
                        if (K > 0) r = __fadd_rn(r, acc[sl][j]);                           // swag.py:95-96 (scale folded)
                        o[j] = __fadd_rn(m[j], r);                                         // swag.py:97
                    }
                    float *orow = a.out + (int64_t)s * a.ld_out + c0;
                    if (full4) {
                        *reinterpret_cast<float4 *>(orow) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (c0 + j < a.D) orow[j] = o[j];
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with ring_s[stage] before it is refilled
    }
}

static int ew_grid(int64_t work_items) {
    const int64_t want = (work_items + kEwThreads - 1) / kEwThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <int SP, int NG>
static int launch_draw(const DrawArgs &a, cudaStream_t st) {
    const size_t smem = (size_t)kStages * (a.K > 0 ? a.K : 0) * kTileCols * sizeof(float);
    URSA_CUDA(cudaFuncSetAttribute(swag_draw_kernel<SP, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (a.D + kTileCols - 1) / kTileCols;
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    swag_draw_kernel<SP, NG><<<grid, kDrawThreads * NG, smem, st>>>(a);
    URSA_LAUNCH_CHECK("swag_draw_kernel");
    return URSA_OK;
}

}  // namespace ursa

using namespace ursa;

extern "C" int ursa_swag_collect(const float *w, float *mean, float *sq_mean, float *dev_row, int64_t n, float keep,
                                 float denom, void *stream) {
    URSA_REQUIRE(w && mean && sq_mean && dev_row && n >= 0, "ursa_swag_collect: bad arguments");
    URSA_REQUIRE(aligned16(w) && aligned16(mean) && aligned16(sq_mean) && aligned16(dev_row),
                 "ursa_swag_collect: buffers must be 16-byte aligned");
    URSA_REQUIRE(denom != 0.f, "ursa_swag_collect: denom == 0");
    if (n == 0) return URSA_OK;
    swag_collect_kernel<<<ew_grid(n >> 2), kEwThreads, 0, (cudaStream_t)stream>>>(w, mean, sq_mean, dev_row, n, keep,
                                                                                 denom);
    URSA_LAUNCH_CHECK("swag_collect_kernel");
    return URSA_OK;
}

extern "C" int ursa_swag_variance(const float *mean, const float *sq_mean, float *var, int64_t n, float clamp,
                                  void *stream) {
    URSA_REQUIRE(mean && sq_mean && var && n >= 0, "ursa_swag_variance: bad arguments");
    if (n == 0) return URSA_OK;
    swag_variance_kernel<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(mean, sq_mean, var, n, clamp);
    URSA_LAUNCH_CHECK("swag_variance_kernel");
    return URSA_OK;
}

extern "C" int ursa_swag_draw(float *out, int64_t ld_out, const float *mean, const float *var, const float *ring,
                              int64_t ld_ring, int K, const float *z2, const float *z1, int64_t ld_z1, int S,
                              int64_t D, float rank_div, uint64_t seed, uint64_t step, void *stream) {
    URSA_REQUIRE(out && mean && var && D >= 0, "ursa_swag_draw: bad arguments");
    URSA_REQUIRE(S >= 1 && S <= URSA_DRAW_MAX_S, "ursa_swag_draw: S must be in [1, %d]", URSA_DRAW_MAX_S);
    URSA_REQUIRE(K >= 0 && K <= URSA_DRAW_MAX_K, "ursa_swag_draw: K must be in [0, %d]", URSA_DRAW_MAX_K);
    URSA_REQUIRE(K == 0 || (ring && z2 && rank_div != 0.f), "ursa_swag_draw: ring, z2 and rank_div are required when K > 0");
    const int64_t d4 = (D + 3) & ~(int64_t)3;
    URSA_REQUIRE(ld_out % 4 == 0 && ld_out >= d4 && aligned16(out), "ursa_swag_draw: out rows must be 16-byte aligned (ld_out %% 4 == 0, ld_out >= roundup4(D))");
    URSA_REQUIRE(K == 0 || (ld_ring % 4 == 0 && ld_ring >= d4 && aligned16(ring)), "ursa_swag_draw: ring rows must be 16-byte aligned and padded to a multiple of 4");
    URSA_REQUIRE(!z1 || (ld_z1 % 4 == 0 && ld_z1 >= d4 && aligned16(z1)), "ursa_swag_draw: z1 rows must be 16-byte aligned");
    URSA_REQUIRE(aligned16(mean) && aligned16(var), "ursa_swag_draw: mean/var must be 16-byte aligned");
    if (D == 0) return URSA_OK;
    DrawArgs a;
    a.out = out; a.mean = mean; a.var = var; a.ring = ring; a.z2 = z2; a.z1 = z1;
    a.ld_out = ld_out; a.ld_ring = ld_ring; a.ld_z1 = ld_z1; a.D = D; a.K = K; a.S = S;
    a.rank_div = rank_div;
    a.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    a.step = step;
    cudaStream_t st = (cudaStream_t)stream;
    if (S <= 8) return launch_draw<8, 1>(a, st);
    if (S <= 16) return launch_draw<16, 2>(a, st);
    return launch_draw<32, 4>(a, st);
}
