// K2: SWAG moment / deviation-ring update, variance, and the batched rank-K draw.
//  collect   reference inference/swa.py:79-90 + inference/subspaces.py:85-89     24 B/param, one pass
//  variance  reference inference/swa.py:106-108                                 12 B/param
//  draw      reference inference/swag.py:85-97 (formula; see header)            (K + 2 + S) * 4 B/param
//
//  gram      reference inference/subspaces.py:116-131 (first half of the PCA fit) K * 4 B/param, one pass
//
// The draw stages [K x 512-column] tiles of the deviation ring in shared memory with the TMA engine
// (cp.async.bulk + mbarrier, double buffered) and contracts them with z2 [S, K] on tcgen05 (3xTF32, accumulators in TMEM),
// so the ring is read from HBM exactly once per 30 draws.  z1 comes from Philox in-register (or from memory in parity mode).
#include "async.cuh"
#include "common.cuh"
#include "tc_common.cuh"

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
// K2b.  out[s, d] = mean[d] + sd[d] z1[s, d] + sum_k (z2[s, k] / rank_div) ring[k, d]      (swag.py:85-97)
// One persistent CTA per SM, tile = 512 columns d of the ring, all S <= 30 draws of the tile per pass, so the ring crosses
// HBM once.  Round 2 rebuilt the kernel around tcgen05: the round-1 version ran the K x S contraction on the warp-level
// mma.sync and spent 45 issue slots per Gaussian (17 of them on fragment loads / splits / moves around the HMMAs; ncu r04,
// r2s15: IPC 2.15, no pipe above 45 %) -- it was bound by instruction issue at 0.47 of the HBM roofline.
//   warp 16  (one lane) = producer: K + 2 row copies (ring rows, mean, var) per tile into a 2-stage ring by the TMA engine
//            (cp.async.bulk + mbarrier), and the tile's 36 tcgen05.mma: D[128 columns x 32 draws] (+)= A[128 x 8] * B[32 x 8]^T,
//            kind::tf32, three products per k-step (lo*hi, hi*lo, hi*hi), four 128-column row tiles, into one of two TMEM
//            buffers; B = z2 / rank_div, split once per launch
//   warps 0..15, thread = column:
//     split  its K ring values become the rows of the A operand, split x = hi + lo (3xTF32), written in the NO-SWIZZLE
//            K-major UMMA layout (row d at 16 B stride, one 4-k chunk per 8 KB plane: 6 STS.128 per part, conflict free);
//            the same thread takes sqrt(var) once -- no lane recomputes another lane's value
//     Gauss  TMEM lane = column: tcgen05.ld hands the thread its 32 low-rank terms; ceil(S / 6) Philox4x32-10 blocks, two at
//            a time in straight-line code -> Box-Muller -> mean + (sd z + lr), one coalesced 128-byte store per warp and draw
// No CTA-wide barrier in the loop: split(i + 1) -> a_ready (mbarrier, 512 arrivals) -> MMA(i + 1) -> mma_bar -> Gauss(i + 1);
// the MMAs of tile i + 1 and the copies of tiles i + 2, i + 3 run under the Gaussian generation of tile i.
// Philox stream: element (s, d) of a call is normal s % 6 of block (s / 6) * D + d at `step` (common.cuh::box_muller6) -- a block
// serves six consecutive DRAWS of one column, so a thread never needs another thread's normals, and the generator work per
// Gaussian is 2/3 of the K1 stream's.
constexpr int kDrawThreads = 512;               // compute threads (the ring kernel adds one producer warp)
constexpr int kTileCols = 512;
constexpr int kKP = 24;                         // ring rows padded to 3 k-steps of 8 (URSA_DRAW_MAX_K)
constexpr int kRows = kKP + 2;                  // + the mean and var rows of the tile (staged by the same bulk copies)
constexpr int kStages = 2;
constexpr int kDrawN = 32;                      // MMA N: TMEM columns per accumulator tile
constexpr int kDrawGroup = URSA_DRAW_MAX_S;     // draws per launch: five Philox blocks of six normals
constexpr uint32_t kAPlane = kTileCols * 16;    // bytes between two 4-k chunks of the A operand (LBO)
constexpr uint32_t kABytes = (kKP / 4) * kAPlane;            // 48 KB per part (hi / lo)
constexpr uint32_t kBPlane = kDrawN * 16;       // LBO of the B operand
constexpr uint32_t kBBytes = (kKP / 4) * kBPlane;            // 3 KB per part
constexpr uint32_t kTmemCols = 2 * (kTileCols / 128) * kDrawN;   // two buffers of four 32-column accumulators
constexpr uint32_t kStageBytes = kRows * kTileCols * 4;
constexpr uint32_t kAOff = kStages * kStageBytes, kBOff = kAOff + 2 * kABytes, kDrawSmem = kBOff + 2 * kBBytes;
static_assert(URSA_DRAW_MAX_K <= kKP && kDrawGroup <= kDrawN && kDrawGroup == 30, "operand shapes");
static_assert(kTileCols == kDrawThreads, "thread = column");
static_assert(kTmemCols == 256, "power-of-two TMEM allocation");
static_assert(kDrawSmem <= 227 * 1024 - 1024, "shared memory");

struct DrawArgs {
    float *out;
    const float *mean, *var, *ring, *z2, *z1;
    int64_t ld_out, ld_ring, ld_z1, D;
    int K, S;
    int s0;                 // index of the launch's first draw within the call (a multiple of 30): Philox block row (s0 + s) / 6
    float rank_div;
    uint2 key;
    uint64_t step;
};

// no-swizzle K-major UMMA descriptor: start | LBO = stride between two 16-byte K chunks | SBO = 128 B between 8-row groups
__device__ __forceinline__ uint64_t draw_desc(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}

// The S <= 30 draws of column c (c < D): groups of SIX draws share a Philox block (box_muller6); two groups at a time run as
// straight-line code so that the two Philox / Box-Muller chains interleave.
template <bool RING, bool EXTZ>
__device__ __forceinline__ void emit_draws(const DrawArgs &a, int64_t c, float m, float sd, const uint32_t (&lr)[kDrawN]) {
    const int S = a.S;
    float *o = a.out + c;
    const float *zp = EXTZ ? a.z1 + c : nullptr;
    uint64_t blk = (uint64_t)(a.s0 / 6) * (uint64_t)a.D + (uint64_t)c;
    auto apply = [&](float z, int s) {
        float r = __fmul_rn(sd, z);                                                        // swag.py:88-89
        if (RING) r = __fadd_rn(r, __uint_as_float(lr[s]));                                // swag.py:95-96 (scale folded)
        *o = __fadd_rn(m, r);                                                              // swag.py:97
        o += a.ld_out;
    };
#pragma unroll
    for (int gp = 0; gp < 3; ++gp) {                                                       // groups 2 gp, 2 gp + 1 (there is no group 5)
        if (gp < 2 && 12 * gp + 12 <= S) {                                                 // uniform: twelve draws, no predicates
            float za[6], zb[6];
            if (EXTZ) {
#pragma unroll
                for (int e = 0; e < 6; ++e) { za[e] = __ldg(zp); zp += a.ld_z1; }
#pragma unroll
                for (int e = 0; e < 6; ++e) { zb[e] = __ldg(zp); zp += a.ld_z1; }
            } else {
                philox_normal6(blk, a.step, a.key, za);
                philox_normal6(blk + (uint64_t)a.D, a.step, a.key, zb);
                blk += 2 * (uint64_t)a.D;
            }
#pragma unroll
            for (int e = 0; e < 6; ++e) apply(za[e], 12 * gp + e);
#pragma unroll
            for (int e = 0; e < 6; ++e) apply(zb[e], 12 * gp + 6 + e);
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int s0 = 12 * gp + 6 * h;
                if (s0 < kDrawGroup && s0 < S) {                                           // uniform; the last group may be ragged
                    float z[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (EXTZ) {
#pragma unroll
                        for (int e = 0; e < 6; ++e)
                            if (s0 + e < S) { z[e] = __ldg(zp); zp += a.ld_z1; }
                    } else {
                        philox_normal6(blk, a.step, a.key, z);
                        blk += (uint64_t)a.D;
                    }
#pragma unroll
                    for (int e = 0; e < 6; ++e)
                        if (s0 + e < S) apply(z[e], s0 + e);
                }
            }
        }
    }
}

// K = 0 (diagonal draw) runs through the same kernel: the producer stages only the mean / var rows, no MMA is issued (the
// commit then arrives at once) and the Gaussians skip the low-rank term.  A stand-alone thread-per-column kernel with global loads
// was slower (1.80 vs 1.65 ms at S = 30, D = 36.5 M): its loads queued behind the 30 stores per column in the LSU (ncu:
// long_scoreboard 5.1, lg_throttle 2.3 warp-cycles per issued instruction).
template <bool EXTZ, bool RING>
__global__ void __launch_bounds__(kDrawThreads + 32, 1) swag_draw_kernel(const DrawArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int MROW = kKP;                                                   // row index of the mean (var = MROW + 1)
    float *raw = reinterpret_cast<float *>(smem_raw);                          // [kStages][kRows][kTileCols]
    __shared__ __align__(8) uint64_t full_bar[kStages], mma_bar[2], a_ready;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const bool producer = warp == kDrawThreads / 32;
    const int K = RING ? a.K : 0, S = a.S;                                      // RING = false: the diagonal draw (K = 0)
    const int KS = (K + 7) >> 3, KC = 2 * KS;                                   // 8-k MMA steps / 4-k chunks in use
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t a_hi = sbase + kAOff, a_lo = a_hi + kABytes, b_hi = sbase + kBOff, b_lo = b_hi + kBBytes;

    if (!producer) {
        // B operand: z2 / rank_div (swag.py:95), split hi + lo; element (s, k) at chunk (k / 4) * kBPlane + s * 16 + (k % 4) * 4
        const float inv_div = K > 0 ? 1.0f / a.rank_div : 0.f;
        for (int e = tid; e < (K > 0 ? kDrawN * kKP : 0); e += kDrawThreads) {
            const int s = e & (kDrawN - 1), k = e / kDrawN;
            const float z = (s < S && k < K) ? __ldg(a.z2 + (int64_t)s * K + k) * inv_div : 0.f;
            const uint32_t h = __float_as_uint(z) & 0xFFFFE000u;
            const uint32_t off = (uint32_t)(k >> 2) * kBPlane + (uint32_t)s * 16u + (uint32_t)(k & 3) * 4u;
            *reinterpret_cast<uint32_t *>(smem_raw + kBOff + off) = h;
            *reinterpret_cast<float *>(smem_raw + kBOff + kBBytes + off) = z - __uint_as_float(h);
        }
        // ring rows K .. 8 KS - 1 of both stages are never written by the bulk copies: zero them once
        for (int st = 0; st < kStages; ++st)
            for (int r = K; r < 8 * KS; ++r) raw[(st * kRows + r) * kTileCols + tid] = 0.f;
        fence_proxy_async();
        if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
    } else if (tid == kDrawThreads) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        mbar_init(&a_ready, kDrawThreads);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int64_t ntiles = (a.D + kTileCols - 1) / kTileCols;
    const int n_my = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // tiles blockIdx.x + i * gridDim.x, i < n_my

    if (producer) {
        if (tid == kDrawThreads) {
            auto issue = [&](int i) {                                           // K + 2 bulk copies of local tile i
                const int stage = i & 1;
                const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTileCols;
                const int64_t rem = a.D - c0;
                const uint32_t bytes = (uint32_t)(rem >= kTileCols ? kTileCols : ((rem + 3) & ~(int64_t)3)) * 4u;
                // mean / var are only guaranteed D elements: copy whole quads, the ragged last quad is read directly
                const uint32_t mv_bytes = (uint32_t)(rem >= kTileCols ? kTileCols : (rem & ~(int64_t)3)) * 4u;
                mbar_arrive_expect_tx(&full_bar[stage], bytes * (uint32_t)K + 2u * mv_bytes);
                float *dst = raw + stage * kRows * kTileCols;
                const float *src = a.ring + c0;
                for (int k = 0; k < K; ++k, src += a.ld_ring, dst += kTileCols) bulk_g2s(dst, src, bytes, &full_bar[stage]);
                if (mv_bytes) {
                    dst = raw + (stage * kRows + MROW) * kTileCols;
                    bulk_g2s(dst, a.mean + c0, mv_bytes, &full_bar[stage]);
                    bulk_g2s(dst + kTileCols, a.var + c0, mv_bytes, &full_bar[stage]);
                }
            };
            if (n_my > 0) issue(0);
            if (n_my > 1) issue(1);
            const uint32_t idesc = make_tf32_idesc(128, kDrawN);
            // descriptors differ only in the start-address field (bits 0-13, 16-byte units): one 32-bit add per operand and MMA.
            // The loop is straight-line code -- this lane shares its scheduler with four busy compute warps, and every
            // instruction it spends between a_ready and the commit delays the tile's Gaussians.
            const uint64_t ah0 = draw_desc(a_hi, kAPlane), al0 = draw_desc(a_lo, kAPlane);
            const uint64_t bh0 = draw_desc(b_hi, kBPlane), bl0 = draw_desc(b_lo, kBPlane);
            for (int i = 0; i < n_my; ++i) {
                mbar_wait(&a_ready, (uint32_t)i & 1u);                          // A operand of tile i written, raw stage i & 1 consumed
                tc_fence_after();
                const uint32_t dbase = tmem + (uint32_t)(i & 1) * (kTmemCols / 2);
#pragma unroll
                for (int ks = 0; ks < kKP / 8; ++ks) {
                    if (ks < KS) {
#pragma unroll
                        for (int t = 0; t < kTileCols / 128; ++t) {
                            const uint32_t d = dbase + (uint32_t)t * kDrawN;
                            const uint32_t ao = ((uint32_t)(2 * ks) * kAPlane + (uint32_t)t * 2048u) >> 4;
                            const uint32_t bo = ((uint32_t)(2 * ks) * kBPlane) >> 4;
                            umma_tf32(d, al0 + ao, bh0 + bo, idesc, ks > 0);
                            umma_tf32(d, ah0 + ao, bl0 + bo, idesc, 1);
                            umma_tf32(d, ah0 + ao, bh0 + bo, idesc, 1);
                        }
                    }
                }
                if (RING) umma_commit(smem_u32(&mma_bar[i & 1]));
                else mbar_arrive(&mma_bar[i & 1]);                               // diagonal draw: nothing to wait for
                if (i + 2 < n_my) issue(i + 2);
            }
        }
    } else {
        float m_next = 0.f, sd_next = 0.f;
        // local tile i has landed: this thread's column -> (mean, sd) registers and its rows of the A operand
        auto split = [&](int i) {
            const int stage = i & 1;
            mbar_wait(&full_bar[stage], (uint32_t)(i >> 1) & 1u);
            const float *rs = raw + stage * kRows * kTileCols + tid;
            const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTileCols;
            const int64_t c = c0 + tid, rem = a.D - c0;
            if (c < a.D) {
                const bool staged = tid < (rem >= kTileCols ? kTileCols : (int)(rem & ~(int64_t)3));
                m_next = staged ? rs[MROW * kTileCols] : a.mean[c];
                sd_next = sqrtf(staged ? rs[(MROW + 1) * kTileCols] : a.var[c]);   // swag.py:88 var.sqrt()
            }
            float x[kKP];                                                       // all loads first: their latencies overlap
#pragma unroll
            for (int r = 0; r < kKP; ++r) x[r] = (r >> 2) < KC ? rs[r * kTileCols] : 0.f;
#pragma unroll
            for (int kc = 0; kc < kKP / 4; ++kc) {
                if (kc < KC) {
                    uint32_t h[4];
                    float l[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        h[e] = __float_as_uint(x[4 * kc + e]) & 0xFFFFE000u;
                        l[e] = x[4 * kc + e] - __uint_as_float(h[e]);
                    }
                    const uint32_t off = kAOff + (uint32_t)kc * kAPlane + (uint32_t)tid * 16u;
                    *reinterpret_cast<uint4 *>(smem_raw + off) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(smem_raw + off + kABytes) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            fence_proxy_async();                                                // generic writes -> tensor-core (async proxy) reads
            tc_fence_before();                                                  // this thread's tcgen05.ld of tile i - 2 (same TMEM buffer) precede the MMAs
            mbar_arrive(&a_ready);
        };
        if (n_my > 0) split(0);
        for (int i = 0; i < n_my; ++i) {
            const float m = m_next, sd = sd_next;
            mbar_wait(&mma_bar[i & 1], (uint32_t)(i >> 1) & 1u);                // tile i's MMAs done: A is free, TMEM buffer full
            if (i + 1 < n_my) split(i + 1);
            tc_fence_after();
            // TMEM lane = row of the 128-column accumulator tile = 32 (warp % 4) + lane; tile warp / 4 at columns 32 (warp / 4)
            const uint32_t taddr = tmem + (uint32_t)(i & 1) * (kTmemCols / 2) + (uint32_t)(warp >> 2) * kDrawN +
                                   ((uint32_t)((warp & 3) * 32) << 16);
            const int64_t c = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTileCols + tid;
            uint32_t lr[kDrawN];
            if (RING) {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(taddr, r0);
                tmem_ld16_nowait(taddr + 16, r1);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) { lr[e] = r0[e]; lr[16 + e] = r1[e]; }
                if (c < a.D) emit_draws<true, EXTZ>(a, c, m, sd, lr);
            } else {
                if (c < a.D) emit_draws<false, EXTZ>(a, c, m, sd, lr);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// K2c.  Gram matrix of the deviation ring, G = R R^T (K x K, fp64), one streaming pass over [K, D]: the first half of the PCA
// subspace (reference inference/subspaces.py:116-131 runs sklearn's randomized SVD on the K x D matrix on the host; with
// K <= 24 rows the SVD is the eigen-decomposition of G, and s V^T = U^T R is one more K2b-shaped pass).
// Round 2: the contraction moved from FFMA (0.35 of the HBM roofline, issue bound) to the warp-level tensor-core MMA
// (m16n8k8, 3xTF32: x = hi + lo, lo*lo dropped) -- the accumulators must come back to registers every few steps anyway
// (below), so the TMEM round trip of tcgen05 has nothing to offer here -- and the reduction order is FIXED (no atomics).
//   producer warp (one lane): ONE 2-D TMA tensor copy per slab -- box = 248 columns x K rows (<= 23 KB) -- into a 6-stage
//            ring; columns beyond D are zero-filled by the TMA engine.  (Row-by-row bulk copies, 20 per slab, kept the
//            producer lane busier than the slab's HBM time: 0.38 of the roofline, 28 % of all samples waiting for data.)
//   16 MMA warps: warp w owns k-steps w, w + 16 of a slab (8 columns each; the order of the contraction index is free, so
//            k = t is column 2 t and k = t + 4 column 2 t + 1).  Per k-step a lane loads three float2 R[g + 8 r][c + 2 t ..]
//            (LDS.64, bank-conflict free at the box's row pitch of 248 floats); they are BOTH the A fragments of the row
//            tiles {0-15, 16-31} and the B fragments of the column tiles {0-7, 8-15, 16-23}, because A = B^T = R.
//            12 mma.sync per k-step: tiles (0-15) x {0-7, 8-15, 16-23} and (16-23) x (16-23); the rest is the transpose.
//   accumulation: the tensor core adds into its accumulator with truncation (a bias that grows with the chain), so a chain
//            covers 16 k-steps = 128 columns, then goes into per-lane fp64 sums; warps -> CTA in warp order through shared
//            memory, CTAs -> G in CTA order by gram_finish_kernel: bit-reproducible.
constexpr int kGramSlab = 248, kGramKSteps = kGramSlab / 8, kGramRows = URSA_DRAW_MAX_K, kGramStages = 6;
constexpr int kGramWarps = 16, kGramThreads = (kGramWarps + 1) * 32;
constexpr int kGramFrag = 16 * 32;                                           // doubles per warp / CTA partial (fragment layout)
constexpr uint32_t kGramStageBytes = kGramRows * kGramSlab * 4;
static_assert(kGramRows == 24, "two 16-row tiles (the second half padded), three 8-column tiles");
static_assert(kGramSlab % 32 == 24 || kGramSlab % 32 == 8, "LDS.64 fragment loads: rows g = 0..3 of a half-warp on distinct bank groups");
static_assert(kGramSlab <= 256 && kGramSlab % 8 == 0, "TMA box");
static_assert(kGramStages * kGramStageBytes + kGramWarps * kGramFrag * 8 <= 227 * 1024 - 1024, "shared memory");

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kGramThreads, 1) ring_gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, int K, int64_t D,
                                                                        double *__restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // TMA destinations: 128-byte aligned
    float *stage_s = reinterpret_cast<float *>(base);                         // [kGramStages][kGramRows][kGramSlab]
    double *red = reinterpret_cast<double *>(base + kGramStages * kGramStageBytes);       // [kGramWarps][kGramFrag]
    __shared__ __align__(8) uint64_t full_bar[kGramStages], empty_bar[kGramStages];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nslab = (D + kGramSlab - 1) / kGramSlab;
    const int n_my = (int)((nslab - blockIdx.x + gridDim.x - 1) / gridDim.x);  // slabs blockIdx.x + i * gridDim.x

    // rows K .. 23 are never written by the copies: zero them once in every stage
    for (int i = tid; i < kGramStages * (kGramRows - K) * kGramSlab; i += kGramThreads) {
        const int st = i / ((kGramRows - K) * kGramSlab), r = i - st * (kGramRows - K) * kGramSlab;
        stage_s[(st * kGramRows + K) * kGramSlab + r] = 0.f;
    }
    if (tid == 0) {
        for (int s = 0; s < kGramStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kGramWarps); }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == kGramWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap);
            for (int i = 0; i < n_my; ++i) {
                const int st = i % kGramStages;
                if (i >= kGramStages) mbar_wait(&empty_bar[st], (uint32_t)(i / kGramStages - 1) & 1u);
                const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kGramSlab;
                mbar_arrive_expect_tx(&full_bar[st], (uint32_t)K * kGramSlab * 4u);   // the whole box, zero fill included
                tma_load_2d(stage_s + st * kGramRows * kGramSlab, &tmap, (int)c0, 0, &full_bar[st]);
            }
        }
        return;
    }

    const int g = lane >> 2, t = lane & 3;
    double sum[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) sum[e] = 0.0;
    float acc[4][4];                                                           // tiles (m0,n0) (m0,n1) (m0,n2) (m1,n2)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    auto drain = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) { sum[4 * q + e] += (double)acc[q][e]; acc[q][e] = 0.f; }
    };
    for (int i = 0; i < n_my; ++i) {
        const int st = i % kGramStages;
        mbar_wait(&full_bar[st], (uint32_t)(i / kGramStages) & 1u);
        const float *rs = stage_s + st * kGramRows * kGramSlab + g * kGramSlab + 2 * t;
#pragma unroll
        for (int j = 0; j < (kGramKSteps + kGramWarps - 1) / kGramWarps; ++j) {
            const int ks = warp + kGramWarps * j;
            if (ks < kGramKSteps) {                                            // warp-uniform
                uint32_t hi[3][2], lo[3][2];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float2 x = *reinterpret_cast<const float2 *>(rs + r * 8 * kGramSlab + 8 * ks);
                    hi[r][0] = __float_as_uint(x.x) & 0xFFFFE000u;
                    hi[r][1] = __float_as_uint(x.y) & 0xFFFFE000u;
                    lo[r][0] = __float_as_uint(x.x - __uint_as_float(hi[r][0]));
                    lo[r][1] = __float_as_uint(x.y - __uint_as_float(hi[r][1]));
                }
                const uint32_t a0h[4] = {hi[0][0], hi[1][0], hi[0][1], hi[1][1]}, a0l[4] = {lo[0][0], lo[1][0], lo[0][1], lo[1][1]};
                const uint32_t a1h[4] = {hi[2][0], 0u, hi[2][1], 0u}, a1l[4] = {lo[2][0], 0u, lo[2][1], 0u};
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    mma_tf32_16x8x8(acc[n], a0l, hi[n][0], hi[n][1]);
                    mma_tf32_16x8x8(acc[n], a0h, lo[n][0], lo[n][1]);
                    mma_tf32_16x8x8(acc[n], a0h, hi[n][0], hi[n][1]);
                }
                mma_tf32_16x8x8(acc[3], a1l, hi[2][0], hi[2][1]);
                mma_tf32_16x8x8(acc[3], a1h, lo[2][0], lo[2][1]);
                mma_tf32_16x8x8(acc[3], a1h, hi[2][0], hi[2][1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        if ((i & 7) == 7) drain();                                             // 8 slabs x <= 2 k-steps per accumulation chain
    }
    drain();
    // warps -> CTA in warp order (fragment layout: entry e of lane l at e * 32 + l), then one coalesced row of the partials
#pragma unroll
    for (int e = 0; e < 16; ++e) red[warp * kGramFrag + e * 32 + lane] = sum[e];
    asm volatile("bar.sync 1, %0;" ::"n"(kGramWarps * 32) : "memory");       // the producer warp has left
    for (int e = tid; e < kGramFrag; e += kGramWarps * 32) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kGramWarps; ++w) v += red[w * kGramFrag + e];
        partial[(int64_t)blockIdx.x * kGramFrag + e] = v;
    }
}

// CTA partials -> G in CTA order.  Fragment entry (q, e) of lane (g, t): tile q = (m0,n0) (m0,n1) (m0,n2) (m1,n2),
// row = 16 (q == 3) + g + 8 (e >> 1), column = 8 min(q, 2) + 2 t + (e & 1); rows 24..31 of tile m1 are padding.
__global__ void __launch_bounds__(kGramFrag) gram_finish_kernel(const double *__restrict__ partial, int nparts, int K,
                                                                double *__restrict__ gram) {
    const int idx = threadIdx.x, lane = idx & 31, qe = idx >> 5;
    const int q = qe >> 2, e = qe & 3, g = lane >> 2, t = lane & 3;
    double v = 0.0;
    for (int p = 0; p < nparts; ++p) v += partial[(int64_t)p * kGramFrag + idx];
    const int row = (q == 3 ? 16 : 0) + g + 8 * (e >> 1), col = 8 * (q < 2 ? q : 2) + 2 * t + (e & 1);
    if (q == 3 && (e >> 1)) return;
    // the tiles hold some entries twice, as (i, j) and (j, i), summed in a different order: the upper triangle is the one kept,
    // so that the result is exactly symmetric
    if (row <= col && col < K) {
        gram[(int64_t)row * K + col] = v;
        gram[(int64_t)col * K + row] = v;
    }
}

// Fallback for rings that are not 16-byte aligned (bulk copies need it): CUDA cores, same fixed-order reduction.
// CTA = 8 warps, slab = 128 columns staged [24][128] in shared memory (double buffered through registers); warp w owns up to
// three 4 x 4 blocks (ib <= jb) of G, its lanes own columns -- conflict-free LDS along a row, 16 FMAs per 8 LDS.
constexpr int kGramCols = 128, kGramFmaThreads = 256;

__global__ void __launch_bounds__(kGramFmaThreads) ring_gram_kernel(const float *__restrict__ ring, int64_t ld, int K, int64_t D,
                                                                    double *__restrict__ partial) {
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
    auto load = [&](int64_t slab, float (&r)[12]) {                       // 24 rows x 128 columns = 3072 = 12 per thread
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int idx = threadIdx.x + u * kGramFmaThreads;
            const int row = idx >> 7;
            const int64_t c = slab * kGramCols + (idx & 127);
            r[u] = (row < K && c < D) ? __ldg(ring + (int64_t)row * ld + c) : 0.f;
        }
    };
    float regs[12];
    int64_t slab = blockIdx.x;
    if (slab < nslab) load(slab, regs);
    for (; slab < nslab; slab += gridDim.x) {
        __syncthreads();                                                   // previous slab consumed
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int idx = threadIdx.x + u * kGramFmaThreads;
            sm[idx >> 7][idx & 127] = regs[u];
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
    // per-CTA partial in matrix layout [24][24] (upper blocks; lanes in butterfly order: deterministic)
    for (int b = 0; b < 3; ++b) {
        if (b >= nblk) continue;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            double v = (double)acc[b][e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int i = ib[b] * 4 + (e >> 2), j = jb[b] * 4 + (e & 3);
            if (lane == 0) partial[(int64_t)blockIdx.x * (kGramRows * kGramRows) + i * kGramRows + j] = v;
        }
    }
}

__global__ void __launch_bounds__(kGramRows * kGramRows) gram_finish_fma_kernel(const double *__restrict__ partial, int nparts,
                                                                               int K, double *__restrict__ gram) {
    const int i = threadIdx.x / kGramRows, j = threadIdx.x % kGramRows;
    if (i > j || j >= K) return;                                           // blocks with ib <= jb: element (i, j), i <= j within the diagonal blocks too
    double v = 0.0;
    const int src = (i / 4 <= j / 4) ? i * kGramRows + j : j * kGramRows + i;
    for (int p = 0; p < nparts; ++p) v += partial[(int64_t)p * (kGramRows * kGramRows) + src];
    gram[(int64_t)i * K + j] = v;
    gram[(int64_t)j * K + i] = v;
}

static int ew_grid(int64_t work_items) {
    const int64_t want = (work_items + kEwThreads - 1) / kEwThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <bool EXTZ, bool RING>
static int launch_draw(const DrawArgs &a, cudaStream_t st) {
    const int64_t ntiles = (a.D + kTileCols - 1) / kTileCols;
    const size_t smem = kDrawSmem + 1024;                                      // + slack for the 1 KB alignment
    URSA_CUDA(cudaFuncSetAttribute(swag_draw_kernel<EXTZ, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    swag_draw_kernel<EXTZ, RING><<<grid, kDrawThreads + 32, smem, st>>>(a);
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
        const int rc = K > 0 ? (z1 ? launch_draw<true, true>(a, st) : launch_draw<false, true>(a, st))
                             : (z1 ? launch_draw<true, false>(a, st) : launch_draw<false, false>(a, st));
        if (rc) return rc;
    }
    return URSA_OK;
}

extern "C" int ursa_swag_gram(const float *ring, int64_t ld_ring, int K, int64_t D, double *gram, void *stream) {
    URSA_REQUIRE(ring && gram && D >= 0, "ursa_swag_gram: bad arguments");
    URSA_REQUIRE(K >= 1 && K <= URSA_DRAW_MAX_K, "ursa_swag_gram: K must be in [1, %d]", URSA_DRAW_MAX_K);
    URSA_REQUIRE(ld_ring >= D, "ursa_swag_gram: ld_ring < D");
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 0) {
        URSA_CUDA(cudaMemsetAsync(gram, 0, sizeof(double) * K * K, st));
        return URSA_OK;
    }
    // per-CTA partial sums live in stream-ordered scratch, so that the reduction order is fixed and concurrent calls on
    // different streams do not share state
    const bool tc = aligned16(ring) && (ld_ring & 3) == 0;
    const int64_t nslab = tc ? (D + kGramSlab - 1) / kGramSlab : (D + kGramCols - 1) / kGramCols;
    const int64_t cap = tc ? (int64_t)sm_count() : (int64_t)sm_count() * 4;
    const int grid = (int)(nslab < cap ? nslab : cap);
    const size_t per = tc ? (size_t)kGramFrag : (size_t)kGramRows * kGramRows;
    double *partial = nullptr;
    {
        // keep the default pool's memory across synchronisations (its default is to hand everything back to the OS)
        static bool pool_set[64] = {};
        int dev = 0;
        URSA_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && !pool_set[dev]) {
            cudaMemPool_t pool;
            uint64_t keep = 64ull << 20;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pool_set[dev] = true;
        }
    }
    URSA_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&partial), sizeof(double) * per * grid, st));
    int rc = URSA_OK;
    if (tc) {
        const size_t smem = kGramStages * kGramStageBytes + kGramWarps * kGramFrag * sizeof(double) + 1024;
        cudaError_t e = cudaFuncSetAttribute(ring_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaFuncSetAttribute(ring_gram_tc_kernel)");
        CUtensorMap tm;
        if (!rc) {
            // ring as a [K rows, D columns] tensor: columns >= D (the padding up to ld_ring included) are out of bounds = zero
            const uint64_t dims[2] = {(uint64_t)D, (uint64_t)K}, strides[1] = {(uint64_t)ld_ring * 4};
            const uint32_t box[2] = {(uint32_t)kGramSlab, (uint32_t)K};
            rc = make_tensor_map(&tm, ring, 2, dims, strides, box, 0);
        }
        if (!rc) {
            ring_gram_tc_kernel<<<grid, kGramThreads, smem, st>>>(tm, K, D, partial);
            count_launch();
            gram_finish_kernel<<<1, kGramFrag, 0, st>>>(partial, grid, K, gram);
            count_launch();
        }
    } else {
        URSA_CUDA(cudaMemsetAsync(partial, 0, sizeof(double) * per * grid, st));
        ring_gram_kernel<<<grid, kGramFmaThreads, 0, st>>>(ring, ld_ring, K, D, partial);
        count_launch();
        gram_finish_fma_kernel<<<1, kGramRows * kGramRows, 0, st>>>(partial, grid, K, gram);
        count_launch();
    }
    if (!rc) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "ring_gram_kernel");
    }
    cudaFreeAsync(partial, st);
    return rc;
}
