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
    static constexpr int NCH = C / 2;                          // channels per epilogue thread
    static constexpr int TILE_COLS = 2 * C;                    // [ACC | LO]
    static constexpr int BN_FLOATS = (kFusedMaxConvs + 1) * 2 * C;
    static constexpr size_t SMEM = (size_t)2 * NPLANES * PLANE_BYTES + (size_t)NSLOT * SLOT_BYTES + 2 * BN_FLOATS * 4 + 128;
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

// FusedStageArgs as in bma_conv_fused.cuh with two differences: w_off[] point at type-6 packed filters, and the a_in
// path (bn_in_off < 0) reads ONE plain fp32 plane of activations (a_in_hi; a_in_lo is ignored).
template <int C>
__global__ void __launch_bounds__(kFusedThreads, 1) preresnet_stage16_kernel(const FusedStageArgs a) {
    using Cfg = F16Cfg<C>;
    constexpr int H = Cfg::H, G = Cfg::G, PITCH = Cfg::PITCH, F0 = Cfg::F0, T = Cfg::T, NCH = Cfg::NCH;
    constexpr int PLANE = Cfg::PLANE_BYTES, NPL = Cfg::NPLANES, NSLOT = Cfg::NSLOT, BNF = Cfg::BN_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NSLOT];
    __shared__ __align__(8) uint64_t empty_bar[NSLOT];
    __shared__ __align__(8) uint64_t acc_full[T];              // MMA -> epilogue: the accumulators of tile t are complete
    __shared__ __align__(8) uint64_t act_ready[T];             // epilogue -> MMA: the planes of tile t hold the next input
    __shared__ uint32_t tmem_base_s;

    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t planes_hi = smem_base;
    const uint32_t planes_lo = smem_base + NPL * PLANE;
    const uint32_t ring = smem_base + 2 * NPL * PLANE;
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
            mbar_init(&act_ready[i], 256);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    {   // zero both plane sets once: pad positions are never written afterwards
        float4 *z = reinterpret_cast<float4 *>(gen_base);
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < 2 * NPL * Cfg::NPOS; i += kFusedThreads) z[i] = zero;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ================= filter producer: one bulk copy per tap =================
        if (elect_one()) {
            uint32_t it = 0;
            for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                const int s = pass / n_groups;
                const float *pk = a.packed + (int64_t)s * a.ld_packed;
                for (int k = 0; k < a.n_convs; ++k) {
                    const unsigned char *w = reinterpret_cast<const unsigned char *>(pk + a.w_off[k]);
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
            // one tap of one tile: C/16 K steps, [A_hi x (B_hi ; B_lo') -> ACC | LO] then [A_lo' x B_hi -> LO]
            auto tap_mmas = [&](uint32_t slot, int tap, int t) {
                const uint32_t shift = (uint32_t)((tap / 3 - 1) * PITCH + (tap % 3 - 1)) + (uint32_t)(128 * t);
                const uint32_t bw = b_w0 + slot * (uint32_t)(Cfg::SLOT_BYTES >> 4);
                const uint32_t d = tmem + (uint32_t)(t * Cfg::TILE_COLS);
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks) {
                    umma_f16(d, mk(a_hi_w + shift + ks * 2 * PL16), mk(bw + ks * 4 * C), idesc_cat, (tap | ks) != 0);
                    umma_f16(d + C, mk(a_lo_w + shift + ks * 2 * PL16), mk(bw + ks * 4 * C), idesc_lo, 1);
                }
            };
            uint32_t it = 0, item = 0;
            for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                for (int k = 0; k < a.n_convs; ++k, ++item) {
                    const uint32_t iph = item & 1u;
                    if (Cfg::TILE_OUTER) {
                        for (int tap = 0; tap < 9; ++tap)
                            mbar_wait_a(smem_u32(&full_bar[(it + tap) % NSLOT]), ((it + tap) / NSLOT) & 1u);
                        for (int t = 0; t < T; ++t) {
                            mbar_wait_a(smem_u32(&act_ready[t + 1 < T ? t + 1 : T - 1]), iph);
                            tc_fence_after();
                            for (int tap = 0; tap < 9; ++tap) tap_mmas((it + tap) % NSLOT, tap, t);
                            umma_commit(smem_u32(&acc_full[t]));
                        }
                        for (int tap = 0; tap < 9; ++tap) umma_commit(smem_u32(&empty_bar[(it + tap) % NSLOT]));
                        it += 9;
                    } else {
                        for (int tap = 0; tap < 9; ++tap, ++it) {
                            const uint32_t slot = it % NSLOT;
                            mbar_wait_a(smem_u32(&full_bar[slot]), (it / NSLOT) & 1u);
                            if (tap == 0) mbar_wait_a(smem_u32(&act_ready[T - 1]), iph);      // in-order arrivals: all tiles
                            tc_fence_after();
                            for (int t = 0; t < T; ++t) tap_mmas(slot, tap, t);
                            umma_commit(smem_u32(&empty_bar[slot]));
                        }
                        for (int t = 0; t < T; ++t) umma_commit(smem_u32(&acc_full[t]));
                    }
                }
            }
        }
    } else {
        // ================= prologue / epilogue warps =================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // channel half
        const int m = q * 32 + lane;                  // row within a tile
        const int ch0 = half * NCH;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ch0;
        const int etid = threadIdx.x - 64;            // 0..255
        const bool from_r = a.bn_in_off >= 0;         // planes = split(relu(bn_in(r_in))) ; else planes = split(a_in)
        const bool to_global = a.a_out_hi != nullptr && a.bn_off[a.n_convs - 1] >= 0;

        // this thread's T rows: element offset inside a pass's image group, image index within the group; -1 = padding
        int32_t rel[T];
        int gimg[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int f = F0 + 128 * t + m;
            const int prow = f / PITCH, pcol = f - prow * PITCH;
            const int r = prow - 1;
            const int g = r / (H + 1), h = r - g * (H + 1);
            const bool ok = pcol >= 1 && r >= 0 && g < G && h < H;
            rel[t] = ok ? ((g * H + h) * H + (pcol - 1)) * C + ch0 : -1;
            gimg[t] = g;
        }
        auto plane_off = [&](int t, int i) { return (uint32_t)(((ch0 + i) >> 3) * PLANE + (F0 + 128 * t + m) * 16); };
        auto goff_of = [&](int pass, int t) -> int32_t {
            const int s = pass / n_groups, n0 = (pass - s * n_groups) * G;
            if (rel[t] < 0 || n0 + gimg[t] >= a.n_images) return -1;
            return (int32_t)((s * a.n_images + n0) * (H * H * C)) + rel[t];
        };
        // BatchNorm (a, b) of the pass's sample -> shared memory buffer `buf`, rescaled for the /16 activation domain:
        //   entry 0 (bn_in) and entries after a mode-1 conv act on the true-domain residual R:  y/16 = relu(a/16 R + b/16)
        //   entries after a mode-0 conv act on x = conv/16:                                      y/16 = relu(a x + b/16)
        auto load_bn = [&](int pass, int buf) {
            const int s = pass / n_groups;
            const float *pk = a.packed + (int64_t)s * a.ld_packed;
            float *bs = bn_all + buf * BNF;
            for (int i = etid; i < 2 * C; i += 256)
                if (from_r) bs[i] = __ldg(pk + a.bn_in_off + i) * kActDown;
            for (int k = 0; k < a.n_convs; ++k)
                if (a.bn_off[k] >= 0)
                    for (int i = etid; i < 2 * C; i += 256)
                        bs[(k + 1) * 2 * C + i] = __ldg(pk + a.bn_off[k] + i) * ((i >= C || a.mode[k] == 1) ? kActDown : 1.f);
        };
        auto load_rows = [&](const float *src, int32_t goff, float *v) {
            if (goff >= 0) {
                const float4 *rp = reinterpret_cast<const float4 *>(src + goff);
#pragma unroll
                for (int i = 0; i < NCH / 4; ++i) {
                    const float4 x4 = __ldg(rp + i);
                    v[4 * i] = x4.x; v[4 * i + 1] = x4.y; v[4 * i + 2] = x4.z; v[4 * i + 3] = x4.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < NCH; ++i) v[i] = 0.f;
            }
        };
        // y[0..NCH) (already / 16) -> fp16 hi / lo' planes of tile t
        auto store_planes = [&](int t, const float *y) {
#pragma unroll
            for (int i = 0; i < NCH; i += 8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_h2(y[i + 2 * j], y[i + 2 * j + 1], hw[j], lw[j]);
                *reinterpret_cast<uint4 *>(gen_base + plane_off(t, i)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(gen_base + NPL * PLANE + plane_off(t, i)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        };
        auto publish = [&](int t) {
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&act_ready[t]);
        };

        float R[T][NCH];                               // residual stream (true domain) -- or the staged prologue source
        int32_t goff[T];
        // the prologue of a pass: R holds the raw source rows (r_in, or the plain a_in plane)
        auto prologue = [&](int pass, int buf) {
            const float *bn = bn_all + buf * BNF;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (goff[t] >= 0) {
                    float y[NCH];
#pragma unroll
                    for (int i = 0; i < NCH; ++i)
                        y[i] = from_r ? fmaxf(fmaf(bn[ch0 + i], R[t][i], bn[C + ch0 + i]), 0.f) : R[t][i] * kActDown;
                    store_planes(t, y);
                }
                publish(t);
                if (!from_r) load_rows(a.r_in, goff[t], R[t]);       // lands long before the first mode-1 epilogue
            }
        };

        uint32_t item = 0;
        int buf = 0;
        {
            const int pass = blockIdx.x;
            load_bn(pass, 0);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                goff[t] = goff_of(pass, t);
                load_rows(from_r ? a.r_in : a.a_in_hi, goff[t], R[t]);
            }
            epi_bar_sync();
            prologue(pass, 0);
        }
        for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
            for (int k = 0; k < a.n_convs; ++k, ++item) {
                const bool last = k == a.n_convs - 1;
                const int mode = a.mode[k];
                const float *bn = bn_all + buf * BNF + (k + 1) * 2 * C;
                const bool has_bn = a.bn_off[k] >= 0;
                const int np = pass + gridDim.x;
                const bool has_next = last && np < n_pass;
                if (has_next) load_bn(np, buf ^ 1);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    mbar_wait_a(smem_u32(&acc_full[t + 1 < T ? t + 1 : T - 1]), item & 1u);
                    tc_fence_after();
                    const uint32_t tcol = t_lane + (uint32_t)(t * Cfg::TILE_COLS);
                    uint32_t ra[NCH], rl[NCH];
                    tmem_ld<NCH>(tcol, ra);
                    tmem_ld<NCH>(tcol + C, rl);
                    tmem_ld_wait();
                    float x[NCH];
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        x[i] = fmaf(__uint_as_float(rl[i]), kLoUnscale, __uint_as_float(ra[i]));     // conv / 16
                        if (mode == 1) {
                            R[t][i] = fmaf(kActUp, x[i], R[t][i]);                                   // residual, true domain
                            x[i] = R[t][i];
                        }
                    }
                    if (!last) {
                        if (goff[t] >= 0) {
                            float y[NCH];
#pragma unroll
                            for (int i = 0; i < NCH; ++i) y[i] = fmaxf(fmaf(bn[ch0 + i], x[i], bn[C + ch0 + i]), 0.f);
                            store_planes(t, y);
                        }
                        publish(t);
                    } else {
                        tc_fence_before();
                        if (goff[t] >= 0) {
                            if (mode == 0) {
#pragma unroll
                                for (int i = 0; i < NCH; ++i) x[i] *= kActUp;
                            }
                            float4 *op = reinterpret_cast<float4 *>(a.r_out + goff[t]);
#pragma unroll
                            for (int i = 0; i < NCH; i += 4) op[i >> 2] = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
                            if (to_global) {
#pragma unroll
                                for (int i = 0; i < NCH; i += 4) {
                                    float y[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j)      // bn entry is in the /16 domain (mode 1): undo
                                        y[j] = fmaxf(fmaf(bn[ch0 + i + j], x[i + j], bn[C + ch0 + i + j]), 0.f) * kActUp;
                                    float4 hv, lv;
                                    split4(y, hv, lv);
                                    reinterpret_cast<float4 *>(a.a_out_hi + goff[t])[i >> 2] = hv;
                                    reinterpret_cast<float4 *>(a.a_out_lo + goff[t])[i >> 2] = lv;
                                }
                            }
                        }
                        if (has_next) {                 // R[t] is dead: fetch the next pass's prologue source under the tail
                            goff[t] = goff_of(np, t);
                            load_rows(from_r ? a.r_in : a.a_in_hi, goff[t], R[t]);
                        }
                    }
                }
                if (has_next) {
                    epi_bar_sync();                     // the next pass's BatchNorm parameters are in bn_all[buf ^ 1]
                    buf ^= 1;
                    prologue(np, buf);
                }
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
static int launch_stage16(const FusedStageArgs &a, cudaStream_t st) {
    using Cfg = F16Cfg<C>;
    const int n_groups = (a.n_images + Cfg::G - 1) / Cfg::G;
    const int n_pass = n_groups * a.n_samples;
    if (n_pass <= 0) return URSA_OK;
    const int sms = sm_count();
    const int grid = n_pass < sms ? n_pass : sms;
    URSA_CUDA(cudaFuncSetAttribute(preresnet_stage16_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    preresnet_stage16_kernel<C><<<grid, kFusedThreads, Cfg::SMEM, st>>>(a);
    URSA_LAUNCH_CHECK("preresnet_stage16_kernel");
    return URSA_OK;
}

}  // namespace ursa
