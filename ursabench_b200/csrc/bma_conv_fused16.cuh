// K3 (PreResNet), fused stage kernel, second generation: 2xFP16-split operands and a tile-granular MMA / epilogue
// wavefront.  Same role and same activation layout idea as bma_conv_fused.cuh (a run of same-shape 3x3 convolutions for
// G images of ONE posterior sample with the activations resident in shared memory as shifted-window "planes"), with
// three changes that remove what bounded the first generation (ncu: tensor pipe 16-43 % active, MMA issue and epilogue
// strictly alternating):
//
//  * operands are FP16 hi + lo pairs instead of TF32 hi + lo: x = hi + lo' * 2^-11 with hi = rn_f16(x),
//    lo' = rn_f16((x - hi) * 2^11).  Both parts carry 11 significant bits (22 together, the same as the 3xTF32 split)
//    but one kind::f16 MMA covers K = 16 channels for the same 4 KB of A operand, and streaming A from shared memory is
//    what bounds these small-N MMAs (tools/umma_probe3.cu: the [N = 2C ; N = C] pair costs 80 / 89 / 113 clk per K = 16 at
//    C = 16 / 32 / 64, against 158 / 176 / 224 clk for the two TF32 pairs it replaces).  Activations are kept in a /16
//    domain so that values up to ~1e6 stay finite in FP16; anything larger becomes inf -> NaN logits (loud, and the
//    Python layer re-runs such a batch on the TF32 engine).  lo' products accumulate in their own TMEM columns (scaled by
//    2^11) and are folded in by the epilogue.
//  * the residual stream lives in the epilogue threads' REGISTERS (each thread owns fixed rows x channels for the whole
//    pass), so TMEM only holds [ACC | LO] per tile, nothing is zeroed (the first MMA of a tile overwrites) and conv2 no
//    longer needs its own filter layout.
//  * C <= 32: all 9 taps of a conv are resident (18 one-tap slots = two convs), the MMA loop is tile-outer, and every
//    tile has its own pair of mbarriers: the epilogue of tile t starts as soon as tile t+1's MMAs are done (tile t's
//    planes are then no longer read as a halo) and conv k+1's MMAs on tile t start as soon as tile t+1's new
//    activations are in the planes -- the tensor pipe does not wait for a whole-conv epilogue any more.  The prologue of
//    the next pass is just one more "epilogue" in that wavefront.  C = 64 (16 KB per tap) keeps the tap-outer ring.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "bma_conv_fused.cuh"

namespace ursa {

constexpr float kLoScale = 2048.f, kLoUnscale = 1.f / 2048.f;     // 2^11
constexpr float kActDown = 0.0625f, kActUp = 16.f;               // activation planes hold y / 16

template <int C>
struct F16Cfg {
    static constexpr int H = 512 / C;                          // 32, 16, 8
    static constexpr int G = C == 16 ? 1 : (C == 32 ? 2 : 3);  // images per pass
    static constexpr int PITCH = H + 1;
    static constexpr int ROWS = G * (H + 1) - 1;
    static constexpr int F0 = PITCH + 1;
    static constexpr int SPAN = (ROWS - 1) * PITCH + H;
    static constexpr int T = (SPAN + 127) / 128;               // 9, 5, 2
    static constexpr int NPOS = ((F0 + T * 128 + PITCH + 2) + 7) & ~7;
    static constexpr int PLANE_BYTES = NPOS * 16;              // 8 halves per position
    static constexpr int NPLANES = C / 8;
    static constexpr int SLOT_BYTES = 4 * C * C;               // one tap: [C/8][2C rows][8 halves]
    static constexpr bool TILE_OUTER = C <= 32;
    static constexpr int NSLOT = TILE_OUTER ? 18 : 6;
    // epilogue: 16 warps = NGT tile groups x NGC channel groups x 4 lane quarters; group (gt, gc) owns the tiles
    // t = gt, gt + NGT, .. and the channels [gc * CPT, (gc + 1) * CPT) -- NGT tiles are in flight at once
    static constexpr int NGT = C == 16 ? 4 : 2;
    static constexpr int NGC = 4 / NGT;
    static constexpr int CPT = C / NGC;                        // channels per epilogue thread: 16, 16, 32
    static constexpr int TPG = (T + NGT - 1) / NGT;            // tiles per group: 3, 3, 1
    static constexpr int TILE_COLS = 2 * C;                    // [ACC | LO]
    static constexpr int BN_FLOATS = (kFusedMaxConvs + 1) * 2 * C;
    // "plane image" of one pass in GLOBAL memory: the activation planes of the tile range [F0, F0 + 128 T) in exactly the
    // shared-memory layout, [hi | lo'][plane][position][8 halves], zero at every pad position -- written by the producer
    // kernel (stem / stride-2 transition conv), pulled in by 2 KB bulk copies per (part, plane, tile)
    static constexpr int IMG_POS = T * 128;
    static constexpr int PASS_BYTES = 2 * NPLANES * IMG_POS * 16;
    static constexpr int TILE_TX = 2 * NPLANES * 2048;
    // output staging (C <= 32, the stages that feed a stride-2 conv): every epilogue warp owns a 32-row x 64-byte tile --
    // its rows' [hi(16 ch) | lo'(16 ch)] halves -- 64-byte swizzled so that row-per-lane 16-byte stores are conflict free,
    // drained by one TMA tensor store per warp and tile.  A global row is OUT_ROW_BYTES = 64 B per 16-channel group.
    static constexpr int OUT_ROW_BYTES = 4 * C;
    static constexpr int STG_SLOTS = C == 16 ? 2 : 1;                           // per warp: y and R tiles (C = 32: one, reused)
    static constexpr int STG_BYTES = C <= 32 ? 16 * 2048 * STG_SLOTS : 0;       // 16 epilogue warps x 2 KB x slots
    static constexpr size_t SMEM = (size_t)STG_BYTES + (size_t)2 * NPLANES * PLANE_BYTES + (size_t)NSLOT * SLOT_BYTES +
                                   2 * BN_FLOATS * 4 + 1024;
    static_assert(T * TILE_COLS <= 512, "TMEM columns");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

// cute::UMMA::InstrDescriptor for kind::f16: D = F32 (bits 4-5 = 1), A = B = F16 (0), K-major both
__device__ __forceinline__ uint32_t make_f16_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// (y0, y1) -> packed hi pair and packed lo' pair
__device__ __forceinline__ void split_h2(float y0, float y1, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(y0, y1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((y0 - hf.x) * kLoScale, (y1 - hf.y) * kLoScale);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

constexpr int kF16Threads = 64 + 512;           // producer warp, MMA warp, 16 epilogue warps

// Warp-level mbarrier wait for the epilogue warps: one lane polls with the NON-blocking test_wait plus a back-off, then the
// warp reconverges.  mbarrier.try_wait may park inside the memory-instruction queue of its SM sub-partition until the phase
// flips or a time limit expires; the MMA-issuing warp shares that queue with a quarter of the epilogue warps (TMEM lane
// quarter = warp % 4 = sub-partition), so parked try_waits that are waiting for MMAs were sitting in front of the very
// tcgen05.mma instructions they waited for (ncu: mio_throttle on every UTCHMMA, MMA cadence 2x the stand-alone probe).
__device__ __forceinline__ bool mbar_test_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        for (uint32_t spin = 0; !mbar_test_wait_a(bar, parity); ++spin) {
            __nanosleep(64);
            if (spin > (1u << 24)) __trap();
        }
    }
    __syncwarp();
}
// add `bytes` to the barrier's pending transaction count WITHOUT arriving (the issuing thread arrives later, with the others)
__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void epi16_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// FusedStageArgs as in bma_conv_fused.cuh with two differences: w_off[] point at type-6 packed filters, and the input
// activations of a pass do not pass through registers at all: they arrive as the pass's PLANE IMAGE (a.pi_in, see F16Cfg)
// by bulk copies straight into the planes, tile by tile, as soon as the previous pass's last MMAs on the neighbouring
// tiles have completed.  Round 1 computed them in the epilogue warps from global loads (r_in -> BN -> ReLU -> split):
// three serial global round trips per thread at every pass boundary, during which the tensor pipe idled for ~30 k clk per
// pass (live event timing: 66 / 75 / 74 k clk per pass against an MMA floor of 39 / 40 / 41 k at C = 16 / 32 / 64).
// r_in only feeds the residual registers and is consumed one conv later.
// DBG: development build with clock64 accounting per role (URSA_STAGE_DBG=1, see launch_stage16); compiled out otherwise
template <int C, bool DBG>
__global__ void __launch_bounds__(kF16Threads, 1) preresnet_stage16_kernel(const __grid_constant__ FusedStageArgs a) {
    using Cfg = F16Cfg<C>;
    constexpr int H = Cfg::H, G = Cfg::G, PITCH = Cfg::PITCH, F0 = Cfg::F0, T = Cfg::T, CPT = Cfg::CPT, TPG = Cfg::TPG;
    constexpr int NGT = Cfg::NGT, NGC = Cfg::NGC;
    constexpr int PLANE = Cfg::PLANE_BYTES, NPL = Cfg::NPLANES, NSLOT = Cfg::NSLOT, BNF = Cfg::BN_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NSLOT];
    __shared__ __align__(8) uint64_t empty_bar[NSLOT];
    __shared__ __align__(8) uint64_t acc_full[T];              // MMA -> epilogue: the accumulators of tile t are complete
    __shared__ __align__(8) uint64_t act_ready[T];             // epilogue -> MMA: the planes of tile t hold the next input
    __shared__ uint32_t tmem_base_s;

    const uint32_t smem_base0 = (smem_u32(smem_raw) + 1023u) & ~1023u;       // staging tiles first (swizzle atoms: 1 KB aligned)
    const uint32_t smem_base = smem_base0 + Cfg::STG_BYTES;
    const uint32_t planes_hi = smem_base;
    const uint32_t planes_lo = smem_base + NPL * PLANE;
    const uint32_t ring = smem_base + 2 * NPL * PLANE;
    unsigned char *stg_base = smem_raw + (smem_base0 - smem_u32(smem_raw));
    unsigned char *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    float *bn_all = reinterpret_cast<float *>(gen_base + 2 * NPL * PLANE + NSLOT * Cfg::SLOT_BYTES);   // [2][BNF]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = (a.n_images + G - 1) / G;
    const int n_pass = n_groups * a.n_samples;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < T; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&act_ready[i], 128 * NGC);
        }
        fence_barrier_init();
    }
    if (threadIdx.x == 0 && a.stagger_ns > 0) {
        // All CTAs run passes of equal length, so without a phase offset every SM reaches its pass boundary -- the burst of
        // output stores and input copies -- at the same moment and the bursts queue up in L2 / HBM while the tensor pipes
        // wait (clock64: the first MMA of a pass waited ~14 k clk for its planes; ~0.7 k with the stores removed).
        const long long t_end = (long long)globaltimer_ns() + (long long)a.stagger_ns * blockIdx.x / gridDim.x;
        while ((long long)globaltimer_ns() < t_end) __nanosleep(256);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    {   // zero both plane sets once: pad positions are never written afterwards
        float4 *z = reinterpret_cast<float4 *>(gen_base);
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < 2 * NPL * Cfg::NPOS; i += kF16Threads) z[i] = zero;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ================= filter producer =================
        // TILE_OUTER: one bulk copy per conv (its 9 taps are contiguous) into one of two 9-slot bundles, one full / empty
        // barrier pair per bundle; otherwise one copy and one barrier pair per tap.
        if (elect_one()) {
            uint32_t it = 0;
            for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                const int s = pass / n_groups;
                const float *pk = a.packed + (int64_t)s * a.ld_packed;
                for (int k = 0; k < a.n_convs; ++k) {
                    const unsigned char *w = reinterpret_cast<const unsigned char *>(pk + a.w_off[k]);
                    if (Cfg::TILE_OUTER) {
                        const uint32_t b = it & 1u, ph = (it >> 1) & 1u;
                        mbar_wait_a(smem_u32(&empty_bar[b]), ph ^ 1u);
                        mbar_expect_tx_a(smem_u32(&full_bar[b]), 9 * Cfg::SLOT_BYTES);
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                ring + b * 9 * Cfg::SLOT_BYTES),
                            "l"(w), "r"((uint32_t)(9 * Cfg::SLOT_BYTES)), "r"(smem_u32(&full_bar[b]))
                            : "memory");
                        ++it;
                    } else
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const uint32_t slot = it % NSLOT, ph = (it / NSLOT) & 1u;
                        mbar_wait_a(smem_u32(&empty_bar[slot]), ph ^ 1u);
                        mbar_expect_tx_a(smem_u32(&full_bar[slot]), Cfg::SLOT_BYTES);
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                ring + slot * Cfg::SLOT_BYTES),
                            "l"(w + (size_t)tap * Cfg::SLOT_BYTES), "r"((uint32_t)Cfg::SLOT_BYTES), "r"(smem_u32(&full_bar[slot]))
                            : "memory");
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc_cat = make_f16_idesc(128, 2 * C), idesc_lo = make_f16_idesc(128, C);
            constexpr uint32_t DESC_HI = (uint32_t)(128 >> 4) | (1u << 14);          // SBO = 128 B, version 1
            constexpr uint32_t PL16 = (uint32_t)(PLANE >> 4);
            const uint32_t a_hi_w = ((planes_hi >> 4) + (uint32_t)F0) | (PL16 << 16);  // LBO = plane stride
            const uint32_t a_lo_w = ((planes_lo >> 4) + (uint32_t)F0) | (PL16 << 16);
            const uint32_t b_w0 = (ring >> 4) | ((uint32_t)(2 * C) << 16);             // LBO = 2C rows x 16 B
            auto mk = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo; };
            // One tap of one tile: C/16 K steps of [A_hi x (B_hi ; B_lo') -> ACC | LO] and [A_lo' x B_hi -> LO].  TAP is a
            // compile-time constant and the descriptors differ from per-tile / per-slot base words by constants, so the
            // issue work per MMA is one integer add: tcgen05.mma is issued by ONE thread and any arithmetic between two MMAs
            // that takes longer than the MMA itself (40 clk) leaves the tensor pipe idle.
            auto tap_mmas = [&](auto tap_c, uint32_t ahw, uint32_t alw, uint32_t bw, uint32_t d) {
                constexpr int TAP = decltype(tap_c)::value;
                constexpr uint32_t SHIFT = (uint32_t)((TAP / 3 - 1) * PITCH + (TAP % 3 - 1));
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks) {
                    umma_f16(d, mk(ahw + SHIFT + ks * 2 * PL16), mk(bw + ks * 4 * C), idesc_cat, (TAP | ks) != 0);
                    umma_f16(d + C, mk(alw + SHIFT + ks * 2 * PL16), mk(bw + ks * 4 * C), idesc_lo, 1);
                }
            };
            constexpr uint32_t SLOTW = (uint32_t)(Cfg::SLOT_BYTES >> 4);
            // Barrier waits of this thread queue up behind its own tcgen05.mma instructions, so a wait that is issued when
            // its result is needed costs a drain of the MMA queue.  Every barrier is therefore TESTED (non-blocking) one step
            // early -- the answer arrives while the MMAs of the current step are being issued -- and only waited for if that
            // early test failed.
            uint32_t it = 0, item = 0;
            bool ok_a = false, ok_b = false;        // early-test results carried into the next step
            long long d_wait_full = 0, d_wait_act = 0, d_wait_act0 = 0, d_t0 = clock64();
#define DBG_T(var, stmt) do { if (DBG && a.dbg) { const long long c0__ = clock64(); stmt; var += clock64() - c0__; } else { stmt; } } while (0)
            for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                for (int k = 0; k < a.n_convs; ++k, ++item) {
                    const uint32_t iph = item & 1u;
                    if (Cfg::TILE_OUTER) {
                        const uint32_t b = iph, fph = (item >> 1) & 1u;     // bundle b holds this conv's 9 taps
                        DBG_T(d_wait_full, mbar_wait_a(smem_u32(&full_bar[b]), fph));
                        const uint32_t bw0 = b_w0 + b * 9u * SLOTW;
                        // ok_a / ok_b: act_ready[0] / act_ready[1] of this conv were seen complete by the previous conv's tail
                        if (!ok_a) DBG_T(d_wait_act0, mbar_wait_a(smem_u32(&act_ready[0]), iph));
                        bool ok_next = ok_b;
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            // tile t reads the planes of tiles t-1 .. t+1 (halo); tile groups publish independently
                            if (t + 1 < T && !ok_next) { if (k == 0) DBG_T(d_wait_act0, mbar_wait_a(smem_u32(&act_ready[t + 1]), iph)); else DBG_T(d_wait_act, mbar_wait_a(smem_u32(&act_ready[t + 1]), iph)); }
                            tc_fence_after();
                            if (t + 2 < T) {
                                ok_next = mbar_test_wait_a(smem_u32(&act_ready[t + 2]), iph);
                            } else if (t + 2 == T) {                        // early tests for the next conv's first two tiles
                                ok_a = mbar_test_wait_a(smem_u32(&act_ready[0]), iph ^ 1u);
                                ok_next = true;
                            } else {
                                ok_b = T > 1 && mbar_test_wait_a(smem_u32(&act_ready[T > 1 ? 1 : 0]), iph ^ 1u);
                            }
                            const uint32_t ahw = a_hi_w + 128u * t, alw = a_lo_w + 128u * t, d = tmem + (uint32_t)(t * Cfg::TILE_COLS);
                            tap_mmas(std::integral_constant<int, 0>{}, ahw, alw, bw0 + 0 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 1>{}, ahw, alw, bw0 + 1 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 2>{}, ahw, alw, bw0 + 2 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 3>{}, ahw, alw, bw0 + 3 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 4>{}, ahw, alw, bw0 + 4 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 5>{}, ahw, alw, bw0 + 5 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 6>{}, ahw, alw, bw0 + 6 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 7>{}, ahw, alw, bw0 + 7 * SLOTW, d);
                            tap_mmas(std::integral_constant<int, 8>{}, ahw, alw, bw0 + 8 * SLOTW, d);
                            umma_commit(smem_u32(&acc_full[t]));
                        }
                        umma_commit(smem_u32(&empty_bar[b]));
                    } else {
                        auto one_tap = [&](auto tap_c) {
                            constexpr int TAP = decltype(tap_c)::value;
                            const uint32_t slot = it % NSLOT;
                            if (!ok_a) DBG_T(d_wait_full, mbar_wait_a(smem_u32(&full_bar[slot]), (it / NSLOT) & 1u));
                            if (TAP == 0)
                                for (int t = 0; t < T; ++t) DBG_T(d_wait_act, mbar_wait_a(smem_u32(&act_ready[t]), iph));
                            tc_fence_after();
                            ok_a = mbar_test_wait_a(smem_u32(&full_bar[(it + 1) % NSLOT]), ((it + 1) / NSLOT) & 1u);
                            const uint32_t bw = b_w0 + slot * SLOTW;
#pragma unroll
                            for (int t = 0; t < T; ++t)
                                tap_mmas(tap_c, a_hi_w + 128u * t, a_lo_w + 128u * t, bw, tmem + (uint32_t)(t * Cfg::TILE_COLS));
                            umma_commit(smem_u32(&empty_bar[slot]));
                            ++it;
                        };
                        one_tap(std::integral_constant<int, 0>{});
                        one_tap(std::integral_constant<int, 1>{});
                        one_tap(std::integral_constant<int, 2>{});
                        one_tap(std::integral_constant<int, 3>{});
                        one_tap(std::integral_constant<int, 4>{});
                        one_tap(std::integral_constant<int, 5>{});
                        one_tap(std::integral_constant<int, 6>{});
                        one_tap(std::integral_constant<int, 7>{});
                        one_tap(std::integral_constant<int, 8>{});
                        for (int t = 0; t < T; ++t) umma_commit(smem_u32(&acc_full[t]));
                    }
                }
            }
            if (DBG && a.dbg) {
                a.dbg[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - d_t0);
                a.dbg[blockIdx.x * 8 + 1] = (unsigned long long)d_wait_full;
                a.dbg[blockIdx.x * 8 + 2] = (unsigned long long)d_wait_act;
                a.dbg[blockIdx.x * 8 + 5] = (unsigned long long)d_wait_act0;
            }
        }
    } else {
        // ================= prologue / epilogue warps =================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;
        const int gt = grp % NGT, gc = grp / NGT;     // tile group, channel group
        const int m = q * 32 + lane;                  // row within a tile
        const int ch0 = gc * CPT;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ch0;
        const int etid = threadIdx.x - 64;            // 0..511
        const bool to_global = Cfg::STG_BYTES > 0 && a.has_out_map != 0 && a.bn_off[a.n_convs - 1] >= 0;
        const bool copier = gc == 0 && m == 0;        // ONE thread per tile group issues the plane-image copies of its tiles
        const unsigned char *pimg = reinterpret_cast<const unsigned char *>(a.pi_in);

        // this thread's rows (tile slot j <-> tile t = gt + j NGT): element offset inside a pass's image group and image
        // index within the group; -1 = padding or no such tile
        int32_t rel[TPG], rel00[TPG];                 // rel00: offset in the compact even-pixel output (r_out_compact)
        int gimg[TPG];
#pragma unroll
        for (int j = 0; j < TPG; ++j) {
            const int t = gt + j * NGT;
            const int f = F0 + 128 * t + m;
            const int prow = f / PITCH, pcol = f - prow * PITCH;
            const int r = prow - 1;
            const int g = r / (H + 1), h = r - g * (H + 1);
            const bool ok = t < T && pcol >= 1 && r >= 0 && g < G && h < H;
            rel[j] = ok ? ((g * H + h) * H + (pcol - 1)) * C + ch0 : -1;
            rel00[j] = (ok && !(h & 1) && !((pcol - 1) & 1)) ? ((g * (H / 2) + h / 2) * (H / 2) + (pcol - 1) / 2) * C + ch0 : -1;
            gimg[j] = g;
        }
        auto plane_off = [&](int t, int i) { return (uint32_t)(((ch0 + i) >> 3) * PLANE + (F0 + 128 * t + m) * 16); };
        auto goff_of = [&](int pass, int j) -> int32_t {
            const int s = pass / n_groups, n0 = (pass - s * n_groups) * G;
            if (rel[j] < 0 || n0 + gimg[j] >= a.n_images) return -1;
            return (int32_t)((s * a.n_images + n0) * (H * H * C)) + rel[j];
        };
        // BatchNorm (a, b) of the pass's sample -> shared memory buffer `buf`, rescaled for the /16 activation domain:
        //   entries after a mode-1 conv act on the true-domain residual R:  y/16 = relu(a/16 R + b/16)
        //   entries after a mode-0 conv act on x = conv/16:                  y/16 = relu(a x + b/16)
        auto load_bn = [&](int pass, int buf) {
            const int s = pass / n_groups;
            const float *pk = a.packed + (int64_t)s * a.ld_packed;
            float *bs = bn_all + buf * BNF;
            for (int k = 0; k < a.n_convs; ++k)
                if (a.bn_off[k] >= 0)
                    for (int i = etid; i < 2 * C; i += 512)
                        bs[(k + 1) * 2 * C + i] = __ldg(pk + a.bn_off[k] + i) * ((i >= C || a.mode[k] == 1) ? kActDown : 1.f);
        };
        auto load_rows = [&](const float *src, int32_t goff, float *v) {
            if (goff >= 0 && src != nullptr) {
                const float4 *rp = reinterpret_cast<const float4 *>(src + goff);
#pragma unroll
                for (int i = 0; i < CPT / 4; ++i) {
                    const float4 x4 = __ldg(rp + i);
                    v[4 * i] = x4.x; v[4 * i + 1] = x4.y; v[4 * i + 2] = x4.z; v[4 * i + 3] = x4.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < CPT; ++i) v[i] = 0.f;
            }
        };
        // y[0..16) (already / 16) -> fp16 hi / lo' planes of tile t, channels ch0 + c0 ..
        auto store_planes16 = [&](int t, int c0, const float *y) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) split_h2(y[i + 2 * jj], y[i + 2 * jj + 1], hw[jj], lw[jj]);
                *reinterpret_cast<uint4 *>(gen_base + plane_off(t, c0 + i)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(gen_base + NPL * PLANE + plane_off(t, c0 + i)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        };
        auto publish = [&](int t) {
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&act_ready[t]);
        };
        // the planes of tile t <- plane image of `pass`: 2 KB per (part, plane), completion counted on act_ready[t] next
        // to the arrivals of the tile's epilogue threads (which guard the TMEM columns of the tile)
        auto fetch_planes = [&](int pass, int t) {
            const int img = a.pi_per_image ? pass % n_groups : pass;
            const unsigned char *src = pimg + (int64_t)img * Cfg::PASS_BYTES + (size_t)t * 2048;
            const uint32_t bar = smem_u32(&act_ready[t]);
            mbar_expect_tx_only(bar, Cfg::TILE_TX);
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
                bulk_g2s_a(planes_hi + pl * PLANE + (F0 + 128 * t) * 16, src + (size_t)pl * Cfg::IMG_POS * 16, 2048, bar);
                bulk_g2s_a(planes_lo + pl * PLANE + (F0 + 128 * t) * 16, src + (size_t)(NPL + pl) * Cfg::IMG_POS * 16, 2048, bar);
            }
        };
        // where this thread's row of tile slot j goes in r_out (full NHWC, or the compact even-pixel tensor)
        auto rout_of = [&](int pass, int j) -> int32_t {
            if (!a.r_out_compact) return goff_of(pass, j);
            const int s = pass / n_groups, n0 = (pass - s * n_groups) * G;
            if (rel00[j] < 0 || n0 + gimg[j] >= a.n_images) return -1;
            return (int32_t)((s * a.n_images + n0) * (H * H * C / 4)) + rel00[j];
        };
        auto prefetch_l2 = [&](const float *src, int32_t g) {
            if (g >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + g));
        };

        float R[TPG][CPT];                             // residual stream (true domain)
        int32_t goff[TPG] = {};
        long long d_wait_acc = 0, d_work = 0, d_sync = 0, d_work_last = 0;
        const bool dbg_me = DBG && a.dbg != nullptr && m == 0 && gc == 0 && (gt == 0 || gt == NGT - 1);
        uint32_t item = 0;
        int buf = 0;
        {
            const int pass = blockIdx.x;
            load_bn(pass, 0);
#pragma unroll
            for (int j = 0; j < TPG; ++j) {
                const int t = gt + j * NGT;
                if (t < T && copier) fetch_planes(pass, t);
                goff[j] = goff_of(pass, j);
                load_rows(a.r_in, goff[j], R[j]);
            }
            epi16_bar_sync();
#pragma unroll
            for (int j = 0; j < TPG; ++j)
                if (gt + j * NGT < T) publish(gt + j * NGT);
        }
        for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
            for (int k = 0; k < a.n_convs; ++k, ++item) {
                const bool last = k == a.n_convs - 1;
                const int mode = a.mode[k];
                const float *bn = bn_all + buf * BNF + (k + 1) * 2 * C;
                const int np = pass + gridDim.x;
                const bool has_next = last && np < n_pass;
                if (k == 0 && np < n_pass) {
                    // Off the critical path (the MMA queue is full at this point): everybody has left the previous pass, so
                    // its BatchNorm buffer can be refilled for the NEXT pass, and the next pass's plane image and residual
                    // rows are pulled into L2 so that the copies at the pass boundary are short.
                    epi16_bar_sync();
                    load_bn(np, buf ^ 1);
                    if (etid == 0)
                        bulk_prefetch_l2(pimg + (int64_t)(a.pi_per_image ? np % n_groups : np) * Cfg::PASS_BYTES, Cfg::PASS_BYTES);
                    if (a.r_in != nullptr) {
#pragma unroll
                        for (int j = 0; j < TPG; ++j) prefetch_l2(a.r_in, goff_of(np, j));
                    }
                }
                if (has_next) epi16_bar_sync();      // bn_all[buf ^ 1] is complete
#pragma unroll
                for (int j = 0; j < TPG; ++j) {
                    const int t = gt + j * NGT;
                    if (t >= T) continue;
                    // tile t's planes are read as a halo by tile t+1's MMAs: wait for those before overwriting them
                    DBG_T(d_wait_acc, mbar_wait_warp(smem_u32(&acc_full[t + 1 < T ? t + 1 : T - 1]), item & 1u));   // in-order commits: tile t too
                    const long long w0__ = (DBG && a.dbg) ? clock64() : 0;
                    tc_fence_after();
                    // pass boundary: the planes of tile t are free from here on -- start pulling in the next pass's input
                    // BEFORE touching the accumulators, so the copy flies under this tile's epilogue
                    if (has_next && copier && !(DBG && (a.dbg_skip & 8))) fetch_planes(np, t);
                    const uint32_t tcol = t_lane + (uint32_t)(t * Cfg::TILE_COLS);
#pragma unroll
                    for (int c0 = 0; c0 < CPT; c0 += 16) {
                        uint32_t ra[16], rl[16];
                        tmem_ld<16>(tcol + c0, ra);
                        tmem_ld<16>(tcol + C + c0, rl);
                        const float4 *ba4 = reinterpret_cast<const float4 *>(bn + ch0 + c0);
                        const float4 *bb4 = reinterpret_cast<const float4 *>(bn + C + ch0 + c0);
                        tmem_ld_wait();
                        float x[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            x[i] = fmaf(__uint_as_float(rl[i]), kLoUnscale, __uint_as_float(ra[i]));   // conv / 16
                            if (mode == 1) {
                                R[j][c0 + i] = fmaf(kActUp, x[i], R[j][c0 + i]);                       // residual, true domain
                                x[i] = R[j][c0 + i];
                            } else if (last) {
                                R[j][c0 + i] = x[i] * kActUp;                                           // a run ending in a mode-0 conv
                            }
                        }
                        if (!last && goff[j] >= 0) {
                            float y[16];
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 a4 = ba4[i >> 2], b4 = bb4[i >> 2];
                                y[i] = relu_nan(fmaf(a4.x, x[i], b4.x));
                                y[i + 1] = relu_nan(fmaf(a4.y, x[i + 1], b4.y));
                                y[i + 2] = relu_nan(fmaf(a4.z, x[i + 2], b4.z));
                                y[i + 3] = relu_nan(fmaf(a4.w, x[i + 3], b4.w));
                            }
                            store_planes16(t, c0, y);
                        }
                    }
                    if (!last) {
                        publish(t);
                        if (DBG && a.dbg) d_work += clock64() - w0__;
                        continue;
                    }
                    // ---- last conv of the pass: the run's output sits in R[j].  Release the tile at once (the next pass's MMAs on
                    // it only wait for the accumulator reads above and for the plane copy); the output is written from the
                    // registers AFTER all tiles of this thread have been released (below), under the next pass's first MMAs.
                    tc_fence_before();
                    if (has_next) mbar_arrive(&act_ready[t]);
                    if (DBG && a.dbg) d_work_last += clock64() - w0__;
                }
                if (last) {
                    const long long s0__ = (DBG && a.dbg) ? clock64() : 0;
#pragma unroll
                    for (int j = 0; j < TPG; ++j) {
                        const int t = gt + j * NGT;
                        if (t >= T) continue;                       // warp-uniform
                        const int32_t my_r = (Cfg::STG_BYTES > 0 && a.has_rrow_map) ? -1 : rout_of(pass, j);
                        if (my_r >= 0 && !(DBG && (a.dbg_skip & 2))) {
                            float4 *op = reinterpret_cast<float4 *>(a.r_out + my_r);
#pragma unroll
                            for (int i = 0; i < CPT; i += 4) op[i >> 2] = make_float4(R[j][i], R[j][i + 1], R[j][i + 2], R[j][i + 3]);
                        }
                        if (Cfg::STG_BYTES > 0 && to_global && !(DBG && (a.dbg_skip & 1))) {
                            // y / 16 = relu(bn_next(R)) / 16, split like the planes, leaves through the warp's own staging tile
                            // and ONE asynchronous TMA tensor store per warp: no thread waits on a global store (round 1: 8
                            // STG.128 per row held every thread in the store queue for ~5 k clk per tile at the pass boundary)
                            // and no warp waits for another (a tile-wide staging slot with group barriers cost ~6 k clk/pass)
                            unsigned char *slot = stg_base + (warp - 2) * (2048 * Cfg::STG_SLOTS);
                            const int sw = (lane >> 1) & 3;
                            const int yrow = pass * Cfg::IMG_POS + 128 * t + 32 * q;
                            if (lane == 0) bulk_wait_group_read0();                 // this warp's previous stores have read the slots
                            __syncwarp();
                            if (goff[j] >= 0) {
#pragma unroll
                                for (int i = 0; i < 16; i += 8) {
                                    uint32_t hw[4], lw[4];
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) {
                                        const int c = i + 2 * jj;
                                        split_h2(relu_nan(fmaf(bn[ch0 + c], R[j][c], bn[C + ch0 + c])),
                                                 relu_nan(fmaf(bn[ch0 + c + 1], R[j][c + 1], bn[C + ch0 + c + 1])), hw[jj], lw[jj]);
                                    }
                                    const int ck = i >> 3;                         // 16-byte chunk: hi {0, 1}, lo' {2, 3}
                                    *reinterpret_cast<uint4 *>(slot + lane * 64 + ((ck ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                    *reinterpret_cast<uint4 *>(slot + lane * 64 + (((2 + ck) ^ sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                                }
                            }
                            const bool both = Cfg::STG_SLOTS == 2 && a.has_rrow_map;     // two slots: one fence, two stores
                            if (!both) {
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_2d(&a.out_map, smem_u32(slot), gc * 16, yrow);
                                    bulk_commit_group();
                                }
                            }
                            if (a.has_rrow_map) {
                                // the raw residual R / 16 in the same row format (the next stage's shortcut conv reads it)
                                unsigned char *rslot = slot + (Cfg::STG_SLOTS == 2 ? 2048 : 0);
                                if (Cfg::STG_SLOTS == 1) {
                                    if (lane == 0) bulk_wait_group_read0();
                                    __syncwarp();
                                }
                                if (goff[j] >= 0) {
#pragma unroll
                                    for (int i = 0; i < 16; i += 8) {
                                        uint32_t hw[4], lw[4];
#pragma unroll
                                        for (int jj = 0; jj < 4; ++jj)
                                            split_h2(R[j][i + 2 * jj] * kActDown, R[j][i + 2 * jj + 1] * kActDown, hw[jj], lw[jj]);
                                        const int ck = i >> 3;
                                        *reinterpret_cast<uint4 *>(rslot + lane * 64 + ((ck ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                        *reinterpret_cast<uint4 *>(rslot + lane * 64 + (((2 + ck) ^ sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                                    }
                                }
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) {
                                    if (both) tma_store_2d(&a.out_map, smem_u32(slot), gc * 16, yrow);
                                    tma_store_2d(&a.rrow_map, smem_u32(rslot), gc * 16, yrow);
                                    bulk_commit_group();
                                }
                            }
                        }
                        if (has_next) {
                            goff[j] = goff_of(np, j);
                            if (!(DBG && (a.dbg_skip & 4))) load_rows(a.r_in, goff[j], R[j]);
                        }
                    }
                    if (DBG && a.dbg) d_sync += clock64() - s0__;
                }
                if (has_next) buf ^= 1;
            }
        }
        if (lane == 0) bulk_wait_group0();           // the warp's last tensor store has left shared memory
        if (dbg_me) {
            if (gt == 0) {
                a.dbg[blockIdx.x * 8 + 3] = (unsigned long long)d_wait_acc;
                a.dbg[blockIdx.x * 8 + 4] = (unsigned long long)d_work;
                a.dbg[blockIdx.x * 8 + 6] = (unsigned long long)d_work_last;
                a.dbg[blockIdx.x * 8 + 7] = (unsigned long long)d_sync;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int C>
static int launch_stage16(const FusedStageArgs &a_in, cudaStream_t st) {
    using Cfg = F16Cfg<C>;
    const int n_groups = (a_in.n_images + Cfg::G - 1) / Cfg::G;
    const int n_pass = n_groups * a_in.n_samples;
    if (n_pass <= 0) return URSA_OK;
    const int sms = sm_count();
    const int grid = n_pass < sms ? n_pass : sms;
    URSA_CUDA(cudaFuncSetAttribute(preresnet_stage16_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    static const bool dbg_on = getenv("URSA_STAGE_DBG") != nullptr;
    static const int stagger_env = getenv("URSA_STAGE_STAGGER_NS") ? atoi(getenv("URSA_STAGE_STAGGER_NS")) : -1;
    FusedStageArgs a = a_in;
    a.stagger_ns = stagger_env >= 0 ? stagger_env : 0;
    if (n_pass < 2 * grid) a.stagger_ns = 0;        // short launches: nothing to spread
    if (dbg_on) {
        // development aid: clock64 breakdown of CTA 0 and the average over CTAs (synchronous; never on in production)
        FusedStageArgs b = a;
        if (const char *e = getenv("URSA_STAGE_SKIP")) b.dbg_skip = atoi(e);
        URSA_CUDA(cudaMalloc(&b.dbg, (size_t)grid * 8 * sizeof(unsigned long long)));
        URSA_CUDA(cudaMemsetAsync(b.dbg, 0, (size_t)grid * 8 * sizeof(unsigned long long), st));
        URSA_CUDA(cudaFuncSetAttribute(preresnet_stage16_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        preresnet_stage16_kernel<C, true><<<grid, kF16Threads, Cfg::SMEM, st>>>(b);
        URSA_LAUNCH_CHECK("preresnet_stage16_kernel");
        URSA_CUDA(cudaStreamSynchronize(st));
        unsigned long long *h = (unsigned long long *)malloc((size_t)grid * 8 * sizeof(unsigned long long));
        URSA_CUDA(cudaMemcpy(h, b.dbg, (size_t)grid * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double avg[8] = {};
        for (int i = 0; i < grid; ++i) for (int k = 0; k < 8; ++k) avg[k] += (double)h[i * 8 + k] / grid;
        const double np_cta = (double)n_pass / grid;
        fprintf(stderr, "[stage16<%d>] passes/CTA %.1f convs %d | per pass (clk): mma total %.0f wait_full %.0f wait_act(k>0) %.0f | epi g0: wait_acc %.0f work(k<last) %.0f | mma wait_act(k=0) %.0f | epi g0 work(last, release) %.0f store phase %.0f\n",
                C, np_cta, a.n_convs, avg[0] / np_cta, avg[1] / np_cta, avg[2] / np_cta, avg[3] / np_cta, avg[4] / np_cta, avg[5] / np_cta, avg[6] / np_cta, avg[7] / np_cta);
        free(h);
        cudaFree(b.dbg);
        return URSA_OK;
    }
    preresnet_stage16_kernel<C, false><<<grid, kF16Threads, Cfg::SMEM, st>>>(a);
    URSA_LAUNCH_CHECK("preresnet_stage16_kernel");
    return URSA_OK;
}

}  // namespace ursa
