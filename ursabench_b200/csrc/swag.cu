// K2: SWAG moment / deviation-ring update, variance, and the batched rank-K draw.
//  collect   reference inference/swa.py:79-90 + inference/subspaces.py:85-89     24 B/param, one pass
//  variance  reference inference/swa.py:106-108                                 12 B/param
//  draw      reference inference/swag.py:85-97 (formula; see header)            (K + 2 + S) * 4 B/param
//
// The draw stages [K x 1024-column] tiles of the deviation ring in shared memory with the TMA engine
// (cp.async.bulk + mbarrier, double buffered) and contracts them with z2 [S, K] on the warp-level tensor-core MMA
// (3xTF32), so the ring is read from HBM exactly once for all S draws.  z1 comes from Philox in-register (or from
// memory in parity mode).
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
// K2b.  CTA = 16 warps, tile = 1024 columns of the ring, double buffered in shared memory by the TMA engine.
// A warp owns 16-column blocks of the tile.  The K x S contraction  acc[s, d] = sum_k z2[s, k] ring[k, d]  runs on
// the warp-level tensor-core MMA (m16n8k8, 3xTF32: operands split hi + lo, lo*lo dropped, fp32 accumulate) because
// its accumulator fragments stay in the registers of the thread that also draws the Gaussians for the same (s, d):
//   A = z2 / rank_div  [16 draws x 8 ring rows]   constant per launch -> loaded once into registers (hi / lo)
//   B = ring tile      [8 ring rows x 8 columns]  LDS.64 from the staged tile (row pitch = 1032 floats: conflict free)
//   C                  [16 draws x 8 columns]     MMA column c of n-tile j <-> column 4*(c/2) + 2*(c%2) + j of the block,
//                                                 so a thread ends up with 4 CONSECUTIVE columns of 4 draws
// = one Philox4x32-10 block (4 normals along d) and one 16-byte store per (draw, thread).  Before this, the contraction
// was K*S FFMAs per column and the kernel was issue bound at 44 % of HBM peak.
constexpr int kDrawThreads = 512;
constexpr int kDrawWarps = kDrawThreads / 32;
constexpr int kTileCols = 1024;
constexpr int kPitch = kTileCols + 8;           // floats; pitch % 32 == 8
constexpr int kKP = 24;                         // ring rows padded to 3 k-steps of 8 (URSA_DRAW_MAX_K)
constexpr int kRows = kKP + 2;                  // + the mean and var rows of the tile (staged by the same bulk copies)
constexpr int kStages = 2;
static_assert(URSA_DRAW_MAX_K <= kKP && URSA_DRAW_MAX_S <= 32, "fragment shapes");

struct DrawArgs {
    float *out;
    const float *mean, *var, *ring, *z2, *z1;
    int64_t ld_out, ld_ring, ld_z1, D;
    int K, S;
    int s0;                 // index of the launch's first draw within the call: Philox block base (s0 + s) * ceil(D / 4)
    float rank_div;
    uint2 key;
    uint64_t step;
};

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT = 16-draw row tiles (1: S <= 16, 2: S <= 32); RING = low-rank term present (K > 0)
template <int MT, bool RING>
__global__ void __launch_bounds__(kDrawThreads, 1) swag_draw_kernel(const DrawArgs a) {
    constexpr int NI = 2 * MT;                                                  // draws per thread: rows nrow + 8 i
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TR = RING ? kRows : 2;                                        // staged rows per tile
    constexpr int MROW = RING ? kKP : 0;                                        // row index of the mean (var = MROW + 1)
    float *ring_s = reinterpret_cast<float *>(smem_raw);                       // [kStages][TR][kPitch]
    __shared__ __align__(16) uint4 afrag[2][3][MT][32];                         // [hi / lo][k-step][row tile][lane]
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane & 3, nrow = lane >> 2;
    const int K = a.K, S = a.S;
    const int KT = (K + 7) >> 3;

    if (RING) {
        // A fragments (z2 / rank_div, split hi + lo): a0 = (row nrow, col q), a1 = (row nrow + 8, col q),
        // a2 = (row nrow, col q + 4), a3 = (row nrow + 8, col q + 4); one uint4 per lane, conflict-free LDS.128
        const float inv_div = 1.0f / a.rank_div;                               // 1/sqrt(max_rank-1) folded in (swag.py:95)
        for (int e = threadIdx.x; e < 3 * MT * 32; e += kDrawThreads) {
            const int ln = e & 31, mt = (e >> 5) % MT, kt = (e >> 5) / MT;
            uint32_t h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s = mt * 16 + (ln >> 2) + (i & 1) * 8, k = kt * 8 + (ln & 3) + (i >> 1) * 4;
                const float z = (s < S && k < K) ? __ldg(a.z2 + (int64_t)s * K + k) * inv_div : 0.f;
                h[i] = __float_as_uint(z) & 0xFFFFE000u;
                l[i] = __float_as_uint(z - __uint_as_float(h[i]));
            }
            afrag[0][kt][mt][ln] = make_uint4(h[0], h[1], h[2], h[3]);
            afrag[1][kt][mt][ln] = make_uint4(l[0], l[1], l[2], l[3]);
        }
        // rows K .. 8*KT-1 of both stages are never written by the bulk copies: zero them once
        const int zr = 8 * KT - K;
        for (int i = threadIdx.x; i < kStages * zr * kPitch; i += kDrawThreads) {
            const int st = i / (zr * kPitch), r = i - st * zr * kPitch;
            ring_s[(st * TR + K) * kPitch + r] = 0.f;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t ntiles = (a.D + kTileCols - 1) / kTileCols;
    const int64_t Dp4 = (a.D + 3) >> 2;                                        // Philox blocks per draw row
    // per-thread draw rows: Philox block base and output row of draw i
    uint64_t ctr_base[NI];
    float *out_row[NI];
    bool act[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int s = (i >> 1) * 16 + (i & 1) * 8 + nrow;
        act[i] = s < S;
        ctr_base[i] = (uint64_t)(a.s0 + s) * (uint64_t)Dp4;
        out_row[i] = a.out + (int64_t)s * a.ld_out;
    }
    const bool dense = S > 16 * MT - 8 && a.z1 == nullptr;                      // (nearly) all fragment rows are real draws

    auto issue = [&](int64_t tile, int stage) {                                 // one thread: K bulk copies
        const int64_t c0 = tile * kTileCols;
        const int64_t rem = a.D - c0;
        const uint32_t cols = (uint32_t)(rem >= kTileCols ? kTileCols : ((rem + 3) & ~(int64_t)3));
        const uint32_t bytes = cols * 4u;
        // mean / var are only guaranteed D elements: copy whole quads, the ragged last quad is read directly
        const uint32_t mv_bytes = (uint32_t)(rem >= kTileCols ? kTileCols : (rem & ~(int64_t)3)) * 4u;
        mbar_arrive_expect_tx(&full_bar[stage], bytes * (uint32_t)K + 2u * mv_bytes);
        float *dst = ring_s + stage * TR * kPitch;
        for (int k = 0; k < K; ++k)
            bulk_g2s(dst + k * kPitch, a.ring + (int64_t)k * a.ld_ring + c0, bytes, &full_bar[stage]);
        if (mv_bytes) {
            bulk_g2s(dst + MROW * kPitch, a.mean + c0, mv_bytes, &full_bar[stage]);
            bulk_g2s(dst + (MROW + 1) * kPitch, a.var + c0, mv_bytes, &full_bar[stage]);
        }
    };
    if (threadIdx.x == 0 && (int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        const float *rs = ring_s + stage * TR * kPitch;
        const float *bfrag = rs + q * kPitch + 2 * nrow;                        // B fragment base of this lane
        {
            const int64_t next = tile + gridDim.x;
            if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);
            mbar_wait(&full_bar[stage], (uint32_t)(it >> 1) & 1u);
        }
#pragma unroll 1
        for (int blk = warp; blk < kTileCols / 16; blk += kDrawWarps) {
            const int cb = blk * 16;
            const int64_t cblk = tile * kTileCols + cb;
            if (cblk >= a.D) break;                                             // warp-uniform
            const int64_t c0 = cblk + 4 * q;                                    // this thread's 4 columns
            const bool whole = cblk + 16 <= a.D;                                // warp-uniform: no ragged edge in this block
            const bool live = c0 < a.D, full4 = c0 + 4 <= a.D;
            float m[4], sd[4];
            if (full4) {
                const float4 mv = *reinterpret_cast<const float4 *>(rs + MROW * kPitch + cb + 4 * q);
                const float4 vv = *reinterpret_cast<const float4 *>(rs + (MROW + 1) * kPitch + cb + 4 * q);
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
            float acc[MT][2][4];                                                // [row tile][n-tile j][c0..c3]
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
            if (RING) {
#pragma unroll
                for (int kt = 0; kt < 3; ++kt) {
                    if (kt < KT) {
                        // B fragments of both n-tiles: b0 = (k = q, n = nrow), b1 = (k = q + 4, n = nrow); columns cb + 2n + j
                        const float2 r0 = *reinterpret_cast<const float2 *>(bfrag + kt * 8 * kPitch + cb);
                        const float2 r1 = *reinterpret_cast<const float2 *>(bfrag + (kt * 8 + 4) * kPitch + cb);
                        const float bv[2][2] = {{r0.x, r1.x}, {r0.y, r1.y}};    // [j][b0 / b1]
                        uint32_t bh[2][2], bl[2][2];
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                bh[j][e] = __float_as_uint(bv[j][e]) & 0xFFFFE000u;
                                bl[j][e] = __float_as_uint(bv[j][e] - __uint_as_float(bh[j][e]));
                            }
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint4 ah4 = afrag[0][kt][mt][lane], al4 = afrag[1][kt][mt][lane];
                            const uint32_t ah[4] = {ah4.x, ah4.y, ah4.z, ah4.w}, al[4] = {al4.x, al4.y, al4.z, al4.w};
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                mma_tf32_16x8x8(acc[mt][j], al, bh[j][0], bh[j][1]);
                                mma_tf32_16x8x8(acc[mt][j], ah, bl[j][0], bl[j][1]);
                                mma_tf32_16x8x8(acc[mt][j], ah, bh[j][0], bh[j][1]);
                            }
                        }
                    }
                }
            }
            // draw i covers row s = 16 (i / 2) + 8 (i % 2) + nrow; column 4q + e of the block is C element
            // (n-tile e & 1, c = 2 (i % 2) + (e >> 1))
            if (dense && whole) {
                // straight-line path: the NI Philox / Box-Muller chains are independent and interleave
                const uint64_t blk4 = (uint64_t)(c0 >> 2);
                float4 zv[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) zv[i] = philox_normal4(ctr_base[i] + blk4, a.step, a.key);
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int mt = i >> 1, h = i & 1;
                    const float z[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w};
                    const float lr[4] = {acc[mt][0][2 * h], acc[mt][1][2 * h], acc[mt][0][2 * h + 1], acc[mt][1][2 * h + 1]};
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float r = __fmul_rn(sd[j], z[j]);                                  // swag.py:88-89
                        if (RING) r = __fadd_rn(r, lr[j]);                                 // swag.py:95-96 (scale folded)
                        o[j] = __fadd_rn(m[j], r);                                         // swag.py:97
                    }
                    if (act[i]) *reinterpret_cast<float4 *>(out_row[i] + c0) = make_float4(o[0], o[1], o[2], o[3]);
                }
                continue;
            }
            if (!live) continue;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (!act[i]) continue;
                const int mt = i >> 1, h = i & 1;
                float z[4];
                if (a.z1) {
                    const float *zr = a.z1 + (int64_t)((i >> 1) * 16 + (i & 1) * 8 + nrow) * a.ld_z1 + c0;
                    if (full4) {
                        const float4 zv = __ldg(reinterpret_cast<const float4 *>(zr));
                        z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) z[j] = (c0 + j < a.D) ? zr[j] : 0.f;
                    }
                } else {
                    const float4 zv = philox_normal4(ctr_base[i] + (uint64_t)(c0 >> 2), a.step, a.key);
                    z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w;
                }
                const float lr[4] = {acc[mt][0][2 * h], acc[mt][1][2 * h], acc[mt][0][2 * h + 1], acc[mt][1][2 * h + 1]};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float r = __fmul_rn(sd[j], z[j]);                                      // swag.py:88-89
                    if (RING) r = __fadd_rn(r, lr[j]);                                     // swag.py:95-96 (scale folded)
                    o[j] = __fadd_rn(m[j], r);                                             // swag.py:97
                }
                float *orow = out_row[i] + c0;
                if (full4) {
                    *reinterpret_cast<float4 *>(orow) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + j < a.D) orow[j] = o[j];
                }
            }
        }
        __syncthreads();   // everyone is done with ring_s[stage] before it is refilled
    }
}

// ---------------------------------------------------------------------------------------------
// Gram matrix of the deviation ring, G = R R^T (K x K, fp64), one streaming pass over [K, D]: the first half of the PCA
// subspace (reference inference/subspaces.py:116-131 runs sklearn's randomized SVD on the K x D matrix on the host; with
// K <= 24 rows the SVD is the eigen-decomposition of G, and s V^T = U^T R is one more K2b-shaped pass).
// CTA = 8 warps, slab = 128 columns staged [24][128] in shared memory (coalesced float4 loads, double buffered through
// registers); warp w owns up to three 4 x 4 blocks (ib <= jb) of G, its lanes own columns -- conflict-free LDS along a row,
// 16 FMAs per 8 LDS.  fp32 partial sums per lane, shuffle reduce, fp64 atomicAdd per entry.
constexpr int kGramCols = 128, kGramRows = URSA_DRAW_MAX_K, kGramThreads = 256;
static_assert(kGramRows == 24, "6 x 6 blocks of 4 rows");

__global__ void __launch_bounds__(kGramThreads) ring_gram_kernel(const float *__restrict__ ring, int64_t ld, int K, int64_t D,
                                                                 int vec_ok, double *__restrict__ gram) {
    __shared__ __align__(16) float sm[kGramRows][kGramCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the 21 upper-triangular 4 x 4 blocks, dealt round-robin to the 8 warps
    int ib[3], jb[3], nblk = 0;
    {
        int t = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j, ++t)
                if ((t & 7) == warp && nblk < 3) { ib[nblk] = i; jb[nblk] = j; ++nblk; }
    }
    float acc[3][16];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[b][e] = 0.f;
    const int64_t nslab = (D + kGramCols - 1) / kGramCols;
    auto load = [&](int64_t slab, float4 (&r)[3]) {                       // 24 rows x 32 float4 = 768 = 3 per thread
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int idx = threadIdx.x + u * kGramThreads;
            const int row = idx >> 5, c4 = (idx & 31) * 4;
            const int64_t c = slab * kGramCols + c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < K && c < D) {
                const float *src = ring + (int64_t)row * ld + c;
                if (vec_ok && c + 4 <= D) v = __ldg(reinterpret_cast<const float4 *>(src));
                else {
                    v.x = __ldg(src);
                    if (c + 1 < D) v.y = __ldg(src + 1);
                    if (c + 2 < D) v.z = __ldg(src + 2);
                    if (c + 3 < D) v.w = __ldg(src + 3);
                }
            }
            r[u] = v;
        }
    };
    float4 regs[3];
    int64_t slab = blockIdx.x;
    if (slab < nslab) load(slab, regs);
    for (; slab < nslab; slab += gridDim.x) {
        __syncthreads();                                                   // previous slab consumed
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int idx = threadIdx.x + u * kGramThreads;
            *reinterpret_cast<float4 *>(&sm[idx >> 5][(idx & 31) * 4]) = regs[u];
        }
        __syncthreads();
        if (slab + gridDim.x < nslab) load(slab + gridDim.x, regs);        // next slab's loads fly during the FMAs
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b < nblk) {
#pragma unroll
                for (int q = 0; q < kGramCols / 32; ++q) {
                    const int c = q * 32 + lane;
                    float vi[4], vj[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { vi[e] = sm[ib[b] * 4 + e][c]; vj[e] = sm[jb[b] * 4 + e][c]; }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int f = 0; f < 4; ++f) acc[b][e * 4 + f] = fmaf(vi[e], vj[f], acc[b][e * 4 + f]);
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        if (b >= nblk) continue;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            float v = acc[b][e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int i = ib[b] * 4 + (e >> 2), j = jb[b] * 4 + (e & 3);
            if (lane == 0 && i < K && j < K) {
                atomicAdd(gram + (int64_t)i * K + j, (double)v);
                if (ib[b] != jb[b]) atomicAdd(gram + (int64_t)j * K + i, (double)v);
            }
        }
    }
}

static int ew_grid(int64_t work_items) {
    const int64_t want = (work_items + kEwThreads - 1) / kEwThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <int MT, bool RING>
static int launch_draw(const DrawArgs &a, cudaStream_t st) {
    const size_t smem = (size_t)kStages * (RING ? kRows : 2) * kPitch * sizeof(float);
    URSA_CUDA(cudaFuncSetAttribute(swag_draw_kernel<MT, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (a.D + kTileCols - 1) / kTileCols;
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    swag_draw_kernel<MT, RING><<<grid, kDrawThreads, smem, st>>>(a);
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
    URSA_REQUIRE(S >= 1, "ursa_swag_draw: S must be positive");
    URSA_REQUIRE(K >= 0 && K <= URSA_DRAW_MAX_K, "ursa_swag_draw: K must be in [0, %d]", URSA_DRAW_MAX_K);
    URSA_REQUIRE(K == 0 || (ring && z2 && rank_div != 0.f), "ursa_swag_draw: ring, z2 and rank_div are required when K > 0");
    const int64_t d4 = (D + 3) & ~(int64_t)3;
    URSA_REQUIRE(ld_out % 4 == 0 && ld_out >= d4 && aligned16(out), "ursa_swag_draw: out rows must be 16-byte aligned (ld_out %% 4 == 0, ld_out >= roundup4(D))");
    URSA_REQUIRE(K == 0 || (ld_ring % 4 == 0 && ld_ring >= d4 && aligned16(ring)), "ursa_swag_draw: ring rows must be 16-byte aligned and padded to a multiple of 4");
    URSA_REQUIRE(!z1 || (ld_z1 % 4 == 0 && ld_z1 >= d4 && aligned16(z1)), "ursa_swag_draw: z1 rows must be 16-byte aligned");
    URSA_REQUIRE(aligned16(mean) && aligned16(var), "ursa_swag_draw: mean/var must be 16-byte aligned");
    if (D == 0) return URSA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // One launch holds the fragments of up to URSA_DRAW_MAX_S draws; more draws go out in groups (the ring is re-read once
    // per group of 32).  The Philox stream is indexed by the draw's position in the CALL, so grouping does not change it.
    for (int g0 = 0; g0 < S; g0 += URSA_DRAW_MAX_S) {
        const int sg = S - g0 < URSA_DRAW_MAX_S ? S - g0 : URSA_DRAW_MAX_S;
        DrawArgs a;
        a.out = out + (int64_t)g0 * ld_out; a.mean = mean; a.var = var; a.ring = ring;
        a.z2 = z2 ? z2 + (int64_t)g0 * K : nullptr;
        a.z1 = z1 ? z1 + (int64_t)g0 * ld_z1 : nullptr;
        a.ld_out = ld_out; a.ld_ring = ld_ring; a.ld_z1 = ld_z1; a.D = D; a.K = K; a.S = sg; a.s0 = g0;
        a.rank_div = rank_div;
        a.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        a.step = step;
        int rc;
        if (K == 0) rc = sg <= 16 ? launch_draw<1, false>(a, st) : launch_draw<2, false>(a, st);
        else rc = sg <= 16 ? launch_draw<1, true>(a, st) : launch_draw<2, true>(a, st);
        if (rc) return rc;
    }
    return URSA_OK;
}

extern "C" int ursa_swag_gram(const float *ring, int64_t ld_ring, int K, int64_t D, double *gram, void *stream) {
    URSA_REQUIRE(ring && gram && D >= 0, "ursa_swag_gram: bad arguments");
    URSA_REQUIRE(K >= 1 && K <= URSA_DRAW_MAX_K, "ursa_swag_gram: K must be in [1, %d]", URSA_DRAW_MAX_K);
    URSA_REQUIRE(ld_ring >= D, "ursa_swag_gram: ld_ring < D");
    cudaStream_t st = (cudaStream_t)stream;
    URSA_CUDA(cudaMemsetAsync(gram, 0, sizeof(double) * K * K, st));
    if (D == 0) return URSA_OK;
    const int64_t nslab = (D + kGramCols - 1) / kGramCols;
    const int64_t cap = (int64_t)sm_count() * 4;
    const int vec_ok = aligned16(ring) && (ld_ring & 3) == 0;
    ring_gram_kernel<<<(int)(nslab < cap ? nslab : cap), kGramThreads, 0, st>>>(ring, ld_ring, K, D, vec_ok, gram);
    URSA_LAUNCH_CHECK("ring_gram_kernel");
    return URSA_OK;
}
