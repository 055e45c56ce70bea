// K3 (PreResNet), fused stage kernel: a whole run of same-shape 3x3 convolutions (the 2n convs of stage 1, the 2n-1
// convs after the stride-2 conv of stages 2 / 3) for G images of ONE posterior sample, with the activations resident in
// SHARED MEMORY and the residual stream resident in TENSOR MEMORY.  Per (sample, image) pair the layer-by-layer path
// moved ~3.7 MB through HBM and re-read every activation tile 9x from L2; here a stage reads its input once and
// writes its output once.
//
// Activation layout in shared memory ("planes"): the G images are stacked vertically with ONE zero row between them and
// ONE zero column per row (pitch = W + 1: the right neighbour of the last pixel of a row is the zero column of the next
// row, which is also that row's left neighbour), so a 3x3 tap is a constant shift of the flat position index f and the
// zero padding is free.  Position f, channel c lives at  plane[c / 4] + f * 16 + (c % 4) * 4  bytes: this is the
// NO-SWIZZLE K-major UMMA layout with SBO = 128 B (8 rows x 16 B contiguous) and LBO = plane stride, so the A operand of
// tap (kh, kw) for output tile t is just a descriptor with start address plane + (F0 + 128 t + (kh-1) pitch + (kw-1)) * 16
// (tools/umma_probe.cu verified these semantics on B200).  tf32 hi / lo planes are separate.
//
// 3xTF32 with two MMAs per K step: the filters of a tap are stored as B' = [B_hi ; B_lo] (2C rows), so
//   MMA1  A_hi x B'   -> [main | lo-term] (N = 2C adjacent TMEM columns)
//   MMA2  A_lo x B_hi -> main             (N = C)
// and the epilogue adds the lo-term columns.  SS-mode tcgen05.mma streams A from shared memory at 128 B/clk, which is the
// binding limit at N = 16..64 (umma_probe: 39 / 40 / 48 clk per MMA at N = 16 / 32 / 64), so fewer, wider MMAs win.
//
// TMEM columns per 128-position tile t: [ACC (C) | LO (C) | R (C)] at 3C*t.  conv1 of a block (mode 0) accumulates into
// [ACC | LO]; conv2 (mode 1, filter rows stored [lo ; hi]) into [LO | R], i.e. straight ONTO the residual.  The epilogue
// warps read the accumulators, fold LO in, zero ACC / LO for the next conv, apply the next BatchNorm + ReLU, split into
// tf32 hi / lo and store the result IN PLACE into the activation planes (all MMAs of the conv have completed).
//
// Warp roles (320 threads, 1 CTA / SM, persistent over (sample, image-group) passes):
//   warp 0  filter producer: one cp.async.bulk per tap (8 C^2 bytes, prepacked in exactly the shared-memory layout)
//   warp 1  TMEM alloc + MMA issuer
//   warps 2-9  prologue (global -> TMEM / planes) and epilogues; warp w owns TMEM lanes 32 (w % 4) and half the channels
#pragma once
#include "preresnet_plan.cuh"
#include "tc_common.cuh"

namespace ursa {

constexpr int kFusedMaxConvs = 13;

struct FusedStageArgs {
    const float *packed;
    int64_t ld_packed;
    int n_convs;
    int64_t w_off[kFusedMaxConvs];     // type-5 packed filters
    int mode[kFusedMaxConvs];          // 0: -> [ACC | LO], result goes to the planes; 1: -> [LO | R], residual update
    int64_t bn_off[kFusedMaxConvs];    // BatchNorm applied by this conv's epilogue to produce the next A (< 0: none)
    int64_t bn_in_off;                 // >= 0: the prologue computes A = split(relu(bn(r_in))); < 0: A comes from a_in_*
    const float *r_in;                 // [S_c][N_c][H][W][C] raw residual stream entering the run
    const float *a_in_hi, *a_in_lo;    // [S_c][N_c][H][W][C] pre-activated tf32 planes (bn_in_off < 0)
    float *r_out;                      // [S_c][N_c][H][W][C]
    float *a_out_hi, *a_out_lo;        // nullable: split(relu(bn_off[last](r_out))) for the next stage's stride-2 conv
    int n_images, n_samples;
    // FP16-split stage kernels: y = relu(bn_off[last](r_out)) / 16 for the next stage's stride-2 conv leaves as FP16 hi / lo'
    // rows [pass][position][hi(C) | lo'(C)] halves through a swizzled staging tile and ONE TMA tensor store per tile
    // (out_map: 2-D fp32 view, C floats x positions, box C x 128, swizzle = row bytes); replaces a_out_hi / a_out_lo
    CUtensorMap out_map;
    int has_out_map = 0;
    // same row format for the RAW residual R / 16 (what the next stage's 1x1 stride-2 shortcut reads as a tenth K block of
    // conv3x3s2_f16_kernel); when set, r_out is not written at all
    CUtensorMap rrow_map;
    int has_rrow_map = 0;
    int pi_per_image = 0;              // pi_in holds one plane image per IMAGE GROUP (shared by all samples), not per pass:
                                       // stage 1, whose input planes carry the test images for the stem conv; r_in == nullptr
                                       // then starts the residual stream at zero
    int r_out_compact = 0;             // FP16-split stage kernels: r_out holds only the even-row / even-column pixels,
                                       // [S_c][N_c][H/2][W/2][C] -- all the 1x1 stride-2 shortcut of the next stage reads
    int stagger_ns = 0;                // CTA i starts i / gridDim * stagger_ns late (see launch_stage16)
    int dbg_skip = 0;                  // development aid: 1 skip a_out stores, 2 skip r_out stores, 4 skip next-pass R loads, 8 skip plane fetch
    unsigned long long *dbg = nullptr; // development aid (URSA_STAGE_DBG=1): per-CTA clock64 breakdown, see launch_stage16
    const void *pi_in;                 // FP16-split stage kernels only: plane images of the passes (see F16Cfg), replaces a_in_* / bn_in_off
};

template <int C>
struct FusedCfg {
    static constexpr int H = 512 / C;                          // 32, 16, 8
    static constexpr int G = C == 16 ? 1 : (C == 32 ? 2 : 3);  // images per pass
    static constexpr int PITCH = H + 1;
    static constexpr int ROWS = G * (H + 1) - 1;               // stacked rows incl. the separators
    static constexpr int F0 = PITCH + 1;                       // flat position of pixel (0, 0)
    static constexpr int SPAN = (ROWS - 1) * PITCH + H;        // positions from the first to the last pixel
    static constexpr int T = (SPAN + 127) / 128;               // 9, 5, 2
    static constexpr int NPOS = ((F0 + T * 128 + PITCH + 2) + 7) & ~7;
    static constexpr int PLANE_BYTES = NPOS * 16;
    static constexpr int NPLANES = C / 4;
    static constexpr int SLOT_BYTES = 8 * C * C;               // one tap: [C/4][2C][4] floats
    static constexpr int NSLOT = C == 16 ? 8 : (C == 32 ? 4 : 2);
    static constexpr int NCH = C / 2;                          // channels per epilogue thread
    static constexpr int TILE_COLS = 3 * C;
    static constexpr int BN_FLOATS = (kFusedMaxConvs + 1) * 2 * C;
    static constexpr size_t SMEM = (size_t)2 * NPLANES * PLANE_BYTES + (size_t)NSLOT * SLOT_BYTES + BN_FLOATS * 4 + 128;
    static_assert(T * TILE_COLS <= 512, "TMEM columns");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

constexpr int kFusedThreads = 320;

// no-swizzle K-major descriptor: start | LBO (stride between the two 16-byte K chunks) | SBO (stride between 8-row groups)
__device__ __forceinline__ uint64_t make_plane_desc(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t *r);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t *r) {
    tmem_ld<16>(taddr, r);
    tmem_ld<16>(taddr + 16, r + 16);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t *r);
template <>
__device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
            "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<32>(uint32_t taddr, const uint32_t *r) {
    tmem_st<16>(taddr, r);
    tmem_st<16>(taddr + 16, r + 16);
}
template <int N>
__device__ __forceinline__ void tmem_st_zero(uint32_t taddr) {
    const uint32_t z = 0;
#pragma unroll
    for (int c = 0; c < N; c += 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(z)
                     : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// y[0..3] -> tf32 hi / lo quads.  hi = y with the low 13 mantissa bits cleared (exactly representable in tf32; one LOP3
// instead of the 4-instruction emulation of cvt.rna), lo = y - hi exactly; kind::tf32 reads only the upper 19 bits of an
// operand, so lo needs no explicit rounding.  |y - (hi + tf32(lo))| <= 2^-21 |y|.
__device__ __forceinline__ void split4(const float *y, float4 &hv, float4 &lv) {
    hv.x = __uint_as_float(__float_as_uint(y[0]) & 0xFFFFE000u);
    hv.y = __uint_as_float(__float_as_uint(y[1]) & 0xFFFFE000u);
    hv.z = __uint_as_float(__float_as_uint(y[2]) & 0xFFFFE000u);
    hv.w = __uint_as_float(__float_as_uint(y[3]) & 0xFFFFE000u);
    lv.x = y[0] - hv.x; lv.y = y[1] - hv.y; lv.z = y[2] - hv.z; lv.w = y[3] - hv.w;
}

template <int C>
__global__ void __launch_bounds__(kFusedThreads, 1) preresnet_stage_kernel(const FusedStageArgs a) {
    using Cfg = FusedCfg<C>;
    constexpr int H = Cfg::H, G = Cfg::G, PITCH = Cfg::PITCH, F0 = Cfg::F0, T = Cfg::T, NCH = Cfg::NCH;
    constexpr int PLANE = Cfg::PLANE_BYTES, NPL = Cfg::NPLANES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t mma_bar, epi_bar;
    __shared__ uint32_t tmem_base_s;

    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t planes_hi = smem_base;
    const uint32_t planes_lo = smem_base + NPL * PLANE;
    const uint32_t ring = smem_base + 2 * NPL * PLANE;
    unsigned char *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    float *bn_s = reinterpret_cast<float *>(gen_base + 2 * NPL * PLANE + Cfg::NSLOT * Cfg::SLOT_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = (a.n_images + G - 1) / G;
    const int n_pass = n_groups * a.n_samples;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::NSLOT; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&mma_bar, 1);
        mbar_init(&epi_bar, 256);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    // zero both plane sets once: pad positions are never written afterwards
    {
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
        // ================= filter producer =================
        if (elect_one()) {
            uint32_t it = 0;
            for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                const int s = pass / n_groups;
                const float *pk = a.packed + (int64_t)s * a.ld_packed;
                for (int k = 0; k < a.n_convs; ++k) {
                    const float *w = pk + a.w_off[k];
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const int slot = it % Cfg::NSLOT;
                        const uint32_t ph = (it / Cfg::NSLOT) & 1u;
                        mbar_wait_a(smem_u32(&empty_bar[slot]), ph ^ 1u);
                        mbar_expect_tx_a(smem_u32(&full_bar[slot]), Cfg::SLOT_BYTES);
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                ring + slot * Cfg::SLOT_BYTES),
                            "l"(w + (int64_t)tap * (Cfg::SLOT_BYTES / 4)), "r"((uint32_t)Cfg::SLOT_BYTES),
                            "r"(smem_u32(&full_bar[slot]))
                            : "memory");
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc_cat = make_tf32_idesc(128, 2 * C), idesc_main = make_tf32_idesc(128, C);
            // Descriptors are built from 32-bit words in 16-byte units so that one integer add per MMA is all the issue
            // work: tcgen05.mma holds its operand registers until the tensor pipe accepts it, so any arithmetic between two
            // MMAs that is longer than the MMA itself (~40 clk at N <= 32) leaves the pipe idle.
            constexpr uint32_t DESC_HI = (uint32_t)(128 >> 4) | (1u << 14);          // SBO = 128 B, version 1
            constexpr uint32_t PL16 = (uint32_t)(PLANE >> 4);
            const uint32_t a_hi_w = ((planes_hi >> 4) + (uint32_t)F0) | (PL16 << 16);  // LBO = plane stride
            const uint32_t a_lo_w = ((planes_lo >> 4) + (uint32_t)F0) | (PL16 << 16);
            const uint32_t b_w0 = (ring >> 4) | ((uint32_t)(2 * C) << 16);             // LBO = 2C rows x 16 B
            auto mk = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo; };
            uint32_t it = 0, epi_phase = 0;
            for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
                for (int k = 0; k < a.n_convs; ++k) {
                    mbar_wait_a(smem_u32(&epi_bar), epi_phase);      // planes written, ACC / LO zeroed
                    epi_phase ^= 1u;
                    tc_fence_after();
                    const int mode = a.mode[k];
                    const uint32_t d_cat = tmem + (mode == 0 ? 0u : (uint32_t)C);
                    const uint32_t d_main = tmem + (mode == 0 ? 0u : (uint32_t)(2 * C));
                    const uint32_t bhi_rows = mode == 0 ? 0u : (uint32_t)C;
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const int slot = it % Cfg::NSLOT;
                        const uint32_t ph = (it / Cfg::NSLOT) & 1u;
                        mbar_wait_a(smem_u32(&full_bar[slot]), ph);
                        tc_fence_after();
                        const uint32_t shift = (uint32_t)((tap / 3 - 1) * PITCH + (tap % 3 - 1));
                        const uint32_t bw = b_w0 + (uint32_t)slot * (Cfg::SLOT_BYTES >> 4);
                        uint32_t ahw = a_hi_w + shift, alw = a_lo_w + shift;
                        uint32_t dc = d_cat, dm = d_main;
#pragma unroll 1
                        for (int t = 0; t < T; ++t) {
#pragma unroll
                            for (int ks = 0; ks < C / 8; ++ks) {
                                umma_tf32(dc, mk(ahw + ks * 2 * PL16), mk(bw + ks * 4 * C), idesc_cat, 1);
                                umma_tf32(dm, mk(alw + ks * 2 * PL16), mk(bw + ks * 4 * C + bhi_rows), idesc_main, 1);
                            }
                            ahw += 128; alw += 128;
                            dc += Cfg::TILE_COLS; dm += Cfg::TILE_COLS;
                        }
                        umma_commit(smem_u32(&empty_bar[slot]));
                    }
                    umma_commit(smem_u32(&mma_bar));
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
        uint32_t mma_phase = 0;
        // plane byte offset of (this thread's row of tile t, channel quad i)
        auto plane_off = [&](int t, int i) { return (uint32_t)(((ch0 + i) >> 2) * PLANE + (F0 + 128 * t + m) * 16); };

        // position bookkeeping of this thread's T rows for a pass: element offset into the [S_c][N_c][H][W][C] tensors
        // (< 2^31: a chunk holds at most 8 x 512 pairs of 16 K elements) or -1 for padding / missing images
        auto locate = [&](int pass, int32_t *goff) {
            const int s = pass / n_groups, ig = pass - s * n_groups;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int f = F0 + 128 * t + m;
                const int prow = f / PITCH, pcol = f - prow * PITCH;
                const int r = prow - 1;
                const int g = r / (H + 1), h = r - g * (H + 1);
                const int n = ig * G + g;
                const bool ok = pcol >= 1 && r >= 0 && g < G && h < H && n < a.n_images;
                goff[t] = ok ? (int32_t)(((((s * a.n_images + n) * H + h) * H + (pcol - 1)) * C) + ch0) : -1;
            }
        };
        // all T rows of the incoming residual stream in one batch of independent loads
        auto load_rows = [&](const float *src, const int32_t *goff, float (*v)[NCH]) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (goff[t] >= 0) {
                    const float4 *rp = reinterpret_cast<const float4 *>(src + goff[t]);
#pragma unroll
                    for (int i = 0; i < NCH / 4; ++i) {
                        const float4 x4 = __ldg(rp + i);
                        v[t][4 * i] = x4.x; v[t][4 * i + 1] = x4.y; v[t][4 * i + 2] = x4.z; v[t][4 * i + 3] = x4.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NCH; ++i) v[t][i] = 0.f;
                }
            }
        };

        int32_t goff[T];
        float rin[T][NCH];
        if ((int)blockIdx.x < n_pass) {
            locate(blockIdx.x, goff);
            load_rows(a.r_in, goff, rin);
        }
        for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
            const int s = pass / n_groups;
            const float *pk = a.packed + (int64_t)s * a.ld_packed;
            // ---- BatchNorm (a, b) of the whole run -> shared memory
            epi_bar_sync();                            // everybody is done with the previous pass's parameters
            if (a.bn_in_off >= 0)
                for (int i = etid; i < 2 * C; i += 256) bn_s[i] = __ldg(pk + a.bn_in_off + i);
            for (int k = 0; k < a.n_convs; ++k)
                if (a.bn_off[k] >= 0)
                    for (int i = etid; i < 2 * C; i += 256) bn_s[(k + 1) * 2 * C + i] = __ldg(pk + a.bn_off[k] + i);
            epi_bar_sync();

            // ---- prologue: residual -> TMEM R, activation -> planes
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const uint32_t tcol = t_lane + (uint32_t)(t * Cfg::TILE_COLS);
                tmem_st<NCH>(tcol + 2 * C, reinterpret_cast<const uint32_t *>(rin[t]));
                tmem_st_zero<NCH>(tcol);
                tmem_st_zero<NCH>(tcol + C);
                if (a.bn_in_off >= 0 && goff[t] >= 0) {
#pragma unroll
                    for (int i = 0; i < NCH; i += 4) {
                        float y[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            y[j] = relu_nan(fmaf(bn_s[ch0 + i + j], rin[t][i + j], bn_s[C + ch0 + i + j]));
                        float4 hv, lv;
                        split4(y, hv, lv);
                        *reinterpret_cast<float4 *>(gen_base + plane_off(t, i)) = hv;
                        *reinterpret_cast<float4 *>(gen_base + NPL * PLANE + plane_off(t, i)) = lv;
                    }
                }
            }
            if (a.bn_in_off < 0) {
                // pre-activated planes from the stride-2 conv: one batch of loads per plane set (rin is free again)
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    load_rows(part == 0 ? a.a_in_hi : a.a_in_lo, goff, rin);
#pragma unroll
                    for (int t = 0; t < T; ++t)
                        if (goff[t] >= 0) {
#pragma unroll
                            for (int i = 0; i < NCH; i += 4)
                                *reinterpret_cast<float4 *>(gen_base + part * NPL * PLANE + plane_off(t, i)) =
                                    make_float4(rin[t][i], rin[t][i + 1], rin[t][i + 2], rin[t][i + 3]);
                        }
                }
            }
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&epi_bar);

            // ---- one epilogue per conv
            for (int k = 0; k < a.n_convs; ++k) {
                const bool last = k == a.n_convs - 1;
                const int mode = a.mode[k];
                const float *bn = bn_s + (k + 1) * 2 * C;
                const bool has_bn = a.bn_off[k] >= 0;
                const bool to_global = last && has_bn && a.a_out_hi != nullptr;
                int32_t gnext[T];
                if (last) {
                    // the MMAs of the last conv are running: fetch the next pass's residual rows under them
                    const int np = pass + gridDim.x;
                    if (np < n_pass) {
                        locate(np, gnext);
                        load_rows(a.r_in, gnext, rin);
                    }
                }
                mbar_wait_a(smem_u32(&mma_bar), mma_phase);
                mma_phase ^= 1u;
                tc_fence_after();
                // double-buffered accumulator reads (the next tile's tcgen05.ld is in flight while this tile is processed)
                // where the register budget allows it
                constexpr bool PIPE = NCH <= 16;
                constexpr int NBUF = PIPE ? 2 : 1;
                uint32_t ra[NBUF][NCH], rl[NBUF][NCH];
                if (PIPE) {
                    tmem_ld<NCH>(t_lane + (mode == 0 ? 0 : 2 * C), ra[0]);
                    tmem_ld<NCH>(t_lane + C, rl[0]);
                }
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t tcol = t_lane + (uint32_t)(t * Cfg::TILE_COLS);
                    const int cur = PIPE ? (t & 1) : 0;
                    if (!PIPE) {
                        tmem_ld<NCH>(tcol + (mode == 0 ? 0 : 2 * C), ra[0]);
                        tmem_ld<NCH>(tcol + C, rl[0]);
                    }
                    tmem_ld_wait();
                    if (PIPE && t + 1 < T) {
                        tmem_ld<NCH>(tcol + Cfg::TILE_COLS + (mode == 0 ? 0 : 2 * C), ra[(t + 1) & (NBUF - 1)]);
                        tmem_ld<NCH>(tcol + Cfg::TILE_COLS + C, rl[(t + 1) & (NBUF - 1)]);
                    }
                    float v[NCH];
#pragma unroll
                    for (int i = 0; i < NCH; ++i) v[i] = __uint_as_float(ra[cur][i]) + __uint_as_float(rl[cur][i]);
                    tmem_st_zero<NCH>(tcol + C);
                    if (mode == 0) tmem_st_zero<NCH>(tcol);
                    else if (!last) tmem_st<NCH>(tcol + 2 * C, reinterpret_cast<const uint32_t *>(v));
                    if (goff[t] < 0) continue;
                    if (last) {
                        float4 *op = reinterpret_cast<float4 *>(a.r_out + goff[t]);
#pragma unroll
                        for (int i = 0; i < NCH; i += 4) op[i >> 2] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (!to_global) continue;
                    }
#pragma unroll
                    for (int i = 0; i < NCH; i += 4) {
                        float y[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x = v[i + j];
                            if (has_bn) x = relu_nan(fmaf(bn[ch0 + i + j], x, bn[C + ch0 + i + j]));
                            y[j] = x;
                        }
                        float4 hv, lv;
                        split4(y, hv, lv);
                        if (last) {
                            reinterpret_cast<float4 *>(a.a_out_hi + goff[t])[i >> 2] = hv;
                            reinterpret_cast<float4 *>(a.a_out_lo + goff[t])[i >> 2] = lv;
                        } else {
                            *reinterpret_cast<float4 *>(gen_base + plane_off(t, i)) = hv;
                            *reinterpret_cast<float4 *>(gen_base + NPL * PLANE + plane_off(t, i)) = lv;
                        }
                    }
                }
                tmem_st_wait();
                if (!last) {
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(&epi_bar);
                } else {
#pragma unroll
                    for (int t = 0; t < T; ++t) goff[t] = gnext[t];
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
static int launch_stage(const FusedStageArgs &a, cudaStream_t st) {
    using Cfg = FusedCfg<C>;
    const int n_groups = (a.n_images + Cfg::G - 1) / Cfg::G;
    const int n_pass = n_groups * a.n_samples;
    if (n_pass <= 0) return URSA_OK;
    const int sms = sm_count();
    const int grid = n_pass < sms ? n_pass : sms;
    URSA_CUDA(cudaFuncSetAttribute(preresnet_stage_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    preresnet_stage_kernel<C><<<grid, kFusedThreads, Cfg::SMEM, st>>>(a);
    URSA_LAUNCH_CHECK("preresnet_stage_kernel");
    return URSA_OK;
}

}  // namespace ursa
