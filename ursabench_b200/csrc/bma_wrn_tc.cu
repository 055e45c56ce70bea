// K3 (WideResNet) on 5th-generation tensor cores: BMA forward of models/wideresnet.py:78-120 (WRN-d-k, WideBasic blocks:
// bn1-relu-conv1(+bias)-[dropout: identity in eval]-bn2-relu-conv2(stride, +bias) + shortcut(x)) straight from the
// [S, D] weight bank and the [S, nb] BatchNorm running-statistics bank (BASELINE.json configs[2]: WRN-28-10, C = 100).
//
// One posterior sample at a time over a chunk of images (a WRN-28-10 sample is 146 MB of filters; its 3x3 convs are
// 0.47 GFLOP per image each -- there is nothing to gain from batching samples inside a launch):
//   pack     bank row -> K-major filters [Cout][tap*Cin_p + ci | shortcut ci], split into TF32 hi / lo planes
//            (3xTF32: hi*hi + hi*lo + lo*hi, fp32 accumulate -> fp32-level accuracy, see bma_mlp_tc.cu); eval-mode BN
//            folded to (a, b); conv bias (+ the 1x1 shortcut's bias) as an epilogue vector
//   stem     3 -> 16 on CUDA cores, writes A = split(relu(bn(R0))) and X = split(R0) padded to 32 channels
//   conv     persistent implicit-GEMM kernel, one CTA per SM: M = 128 output pixels (4 rows x 32 | 8 x 16 | 2 images x 8 x 8),
//            N = a tile of <= 160 output channels, K = 9 taps x Cin (+ Cin_block for a folded 1x1 shortcut).  Every
//            [128 pixels x 32 channels] A slice is ONE 4-D tiled TMA box at shifted coordinates (zero padding = TMA
//            out-of-bounds fill, applied AFTER the activation like PyTorch pads relu(bn(x))); stride-2 convs read four
//            parity-split tensor maps.  The 1x1 (strided) shortcut conv of a transition block is folded into conv2's
//            accumulation as extra K blocks over the split RAW block input X -- no separate kernel, no extra pass.
//            warp 0: TMA producer | warp 1: tcgen05.mma issuer (3 MMAs per K = 8 step) | warps 2-9: epilogue.
//            CTA PAIRS (default): the two CTAs of a cluster take adjacent M tiles of one N tile and run ONE M = 256
//            tcgen05.mma.cta_group::2 -- each SM streams its own A tile and HALF of B (52 KB instead of 72 KB per K block, so the
//            ring has 4 stages); both CTAs' TMA copies are counted on the leader's barrier, the leader's commits are multicast
//            to both CTAs, both epilogues release the leader's accumulator barrier.  One chain per tile: 259 -> 280 TFLOP/s;
//            with the accumulation segments: 228-238 -> 242.
//            TWO-LEVEL ACCUMULATION: the tensor core adds into TMEM with truncated alignment, a bias that grows with the
//            length of the MMA chain (measured: 2e-5 on WRN-16-2 probabilities with K = 1152 chains, 20x the fp32 noise
//            floor).  So a chain covers only WRN_SEG K blocks (K = 128, 48 MMAs); the epilogue warps drain each segment
//            from TMEM and add it to fp32 register accumulators (round-to-nearest).  Measured on WRN-28-10 (max |p - p_fp64|,
//            cuDNN fp32 = 1.9e-5 on the same inputs): SEG 2 / 4 / 9 / 18 / 45 / one chain = 0.8 / 1.1 / 2.5 / 4.7 / 10.6 / 48 e-5
//            at 192 / 212 / 238-250 / 234 / 235 / 250 TFLOP/s.  Restarting a chain / rotating the accumulator costs 3-5 % by itself
//            (timing experiment without the hand-shakes); the rest of the gap is the epilogue BETWEEN two tiles (bias, residual,
//            BN, split, stores: ~10 k cycles) during which nobody drains, while three rotating accumulators only cover
//            2 segments - the issuer's lead.  So a tile's first segment is WRN_SEG0_FACTOR x longer (12 K blocks: 207 -> 224
//            TFLOP/s at unchanged accuracy, 0.8e-5 on the same inputs); prefetching the residual one chunk ahead was tried and
//            lost to register spills (the kernel sits at ptxas' 168-register ceiling for 320 threads).  min(4, 512 / N) TMEM accumulators
//            rotate per segment, so draining segment j overlaps the MMAs of the following segments and a tile's epilogue
//            (bias, residual, BN, split, stores) overlaps the next tile's first segments.
//   epilogue v = acc + bias (+ residual);  raw fp32 v | split(v) (next block's shortcut input) | split(relu(bn_next(v)))
//   head     BN + ReLU + 8x8 average pool + linear -> logits;  accumulate (bma_metrics.cu) in sample order
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "bma_epilogue.cuh"

namespace ursa {

constexpr int WRN_THREADS = 320, WRN_MAX_STAGES = 4, WRN_MAX_BLOCKS = 8;
constexpr int WRN_SEG0_FACTOR = 3;                 // a tile's FIRST segment is this many times longer, see the kernel comment
constexpr int WRN_SEG = 4;                         // K blocks (of 32) per TMEM accumulation segment, see the kernel comment
constexpr int WRN_MAX_TBUF = 4;                    // TMEM accumulators: min(4, 512 / bn_tile) at a column pitch of bn_tile
constexpr int WRN_EPI_CHUNKS = 5;                  // 16-column chunks per epilogue thread: bn_tile / 2 <= 80
constexpr uint32_t WRN_A_BYTES = 128 * 128;        // 128 pixels x 32 channels x fp32, 128-byte swizzle rows
constexpr int WRN_CHUNK_IMAGES = 512;
// FP16-split variant (F16 = true, URSA_ALGO_TCGEN05_F16): planes and filters are halves, x = hi + lo' 2^-11 (22 significant bits like
// the TF32 split); a K block is still 32 channels = a 64-byte swizzle row; per K = 16 step hi*hi -> ACC columns, hi*lo' and lo'*hi ->
// the LO accumulator, result = ACC + LO 2^-11 in the epilogue registers.  Half the operand bytes and half the tensor-pipe time per
// MAC of 3xTF32 (kind::f16 runs at twice the TF32 rate).  Range: fp16's (|x| <= 65 504), overflow surfaces as NaN logits.
constexpr uint32_t WRN_A_BYTES_H = 128 * 64;       // 128 pixels x 32 channels x fp16, 64-byte swizzle rows
constexpr int WRN_MAX_STAGES_H = 8;
constexpr float kWrnLoScale = 2048.f, kWrnLoUnscale = 1.f / 2048.f;

__device__ __forceinline__ uint32_t wrn_f16_idesc(int m, int n) {        // D = F32, A = B = F16, K-major both
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int NCTA>
__device__ __forceinline__ void wrn_umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (NCTA == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// four floats -> packed hi halves and packed lo' halves (8 bytes each)
__device__ __forceinline__ void wrn_split_h4(float y0, float y1, float y2, float y3, uint2 &hi, uint2 &lo) {
    const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((y0 - f01.x) * kWrnLoScale, (y1 - f01.y) * kWrnLoScale);
    const __half2 l23 = __floats2half2_rn((y2 - f23.x) * kWrnLoScale, (y3 - f23.y) * kWrnLoScale);
    hi = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
}

struct WrnMaps {
    CUtensorMap a_hi[4], a_lo[4];      // index = h-parity * 2 + w-parity for stride 2; [0] only for stride 1
    CUtensorMap x_hi, x_lo;            // shortcut input (raw block input, split), parity (0, 0) lattice for stride 2
    CUtensorMap b_hi, b_lo;            // filters [Cout][Ktot]
};

struct WrnConvArgs {
    int cout, bn_tile, hout, stride, n_images;
    int kchunks, xchunks, m_tiles, n_tiles, stages, seg, seg0;
    const float *bias;                 // [cout]
    const float *bn;                   // [2 * cout] (a, b) of the BN that follows, or null
    const float *res;                  // identity shortcut: raw block input [P][hout][hout][cout], or null
    float *out_raw;                    // v
    float *out_hi, *out_lo;            // split(relu(bn(v)))
    float *outx_hi, *outx_lo;          // split(v)
    double *stats;                     // train mode (BatchNorm re-estimation): [batches][2][cout] sums of v and v^2, or null
    int batch;                         // images per batch (train mode)
};

// NCTA = 2: CTA pair (cta_group::2).  The two CTAs of a cluster take adjacent M tiles of the same N tile and run ONE M = 256 MMA:
// each streams its own A tile and HALF of the B tile from its shared memory (52 KB instead of 72 KB per K block -> a 4-stage
// ring), the leader (cluster rank 0) issues the MMAs and multicasts its commits; both epilogues release the leader's accumulator.
template <int NCTA, bool F16 = false>
__global__ void __launch_bounds__(WRN_THREADS, 1)
wrn_conv_tc_kernel(const __grid_constant__ WrnMaps maps, const WrnConvArgs a) {
    extern __shared__ unsigned char smem_raw[];
    constexpr uint32_t A_BYTES = F16 ? WRN_A_BYTES_H : WRN_A_BYTES;
    constexpr uint32_t ROW_BYTES = F16 ? 64u : 128u;               // one K block (32 channels) of an operand row
    __shared__ __align__(8) uint64_t full_bar[WRN_MAX_STAGES_H];
    __shared__ __align__(8) uint64_t empty_bar[WRN_MAX_STAGES_H];
    __shared__ __align__(8) uint64_t tfull_bar[WRN_MAX_TBUF];
    __shared__ __align__(8) uint64_t tempty_bar[WRN_MAX_TBUF];
    __shared__ __align__(8) uint64_t lo_empty_bar;                 // FP16-split: the tile's LO accumulator has been drained
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int WT = a.hout, HT = a.hout >= 16 ? 128 / a.hout : a.hout;
    const int tpi = (a.hout * a.hout) / 128;                 // tiles per image (0 when one tile spans 2 images)
    const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
    const uint32_t b_rows = (uint32_t)a.bn_tile / NCTA;        // rows of the B tile this CTA stages
    const uint32_t b_bytes = b_rows * ROW_BYTES;
    const uint32_t stage_bytes = 2 * A_BYTES + 2 * b_bytes;
    // FP16-split TMEM layout: two rotating ACC accumulators at columns [0, bn) and [bn, 2 bn), ONE LO accumulator at [2 bn, 3 bn).
    // LO holds the 2^-11-scaled cross terms: the truncation bias of its chain enters the result 2^-11 times smaller, so it is one
    // chain per tile (no two-level accumulation) and is drained once, after the tile's last ACC segment -- which is what lets
    // bn_tile stay at 160 (3 x 160 = 480 of the 512 columns) instead of 80 and halves the activation re-reads of stages 1 and 2.
    const uint32_t acc_pitch = (uint32_t)a.bn_tile;
    const uint32_t lo_col = 2u * (uint32_t)a.bn_tile;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int kb3 = 9 * a.kchunks, k_blocks = kb3 + a.xchunks;
    // work units: (M tile [pair], N tile), N fastest; unit u -> this CTA's tile (mt, nt)
    const int total_tiles = ((a.m_tiles + NCTA - 1) / NCTA) * a.n_tiles;
    const int first_unit = (int)blockIdx.x / NCTA, unit_stride = (int)gridDim.x / NCTA;
    const uint32_t ntbuf = F16 ? 2u : (512u / acc_pitch < (uint32_t)WRN_MAX_TBUF ? 512u / acc_pitch : (uint32_t)WRN_MAX_TBUF);

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < WRN_MAX_TBUF; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], NCTA * (WRN_THREADS - 64) / 32);   // one arrival per epilogue warp (of both CTAs)
        }
        mbar_init(&lo_empty_bar, NCTA * (WRN_THREADS - 64) / 32);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (NCTA == 2) tmem_alloc_2cta(&tmem_base_s, 512);
        else tmem_alloc(&tmem_base_s, 512);
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();                        // the peer's barriers are initialised before anybody signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            uint32_t st = 0, ph = 0;                                   // ring position / phase, advanced at the loop bottom
            const uint32_t nstages = (uint32_t)a.stages;
            for (int t = first_unit; t < total_tiles; t += unit_stride) {
                const int nt = t % a.n_tiles, mt = (t / a.n_tiles) * NCTA + (int)cta_rank;
                int n0, h0;
                if (tpi > 0) { n0 = mt / tpi; h0 = (mt % tpi) * HT; } else { n0 = mt * 2; h0 = 0; }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_a(smem_u32(&empty_bar[st]), ph ^ 1u);
                    uint32_t fb = smem_u32(&full_bar[st]);
                    uint32_t base = smem_base + st * stage_bytes;
                    if (NCTA == 2) {
                        // both CTAs' copies are counted on the LEADER's barrier (cluster addresses), which expects both stages
                        if (cta_rank == 0) mbar_expect_tx_a(fb, 2 * stage_bytes);
                        fb = mapa_u32(fb, 0);
                        base = mapa_u32(base, cta_rank);
                    } else {
                        mbar_expect_tx_a(fb, stage_bytes);
                    }
                    const int brow = nt * a.bn_tile + (int)(cta_rank * b_rows);
                    if (kb < kb3) {
                        const int tap = kb / a.kchunks, cc = kb - tap * a.kchunks;
                        const int kh = tap / 3, kw = tap - kh * 3;
                        int mi = 0, cw = kw - 1, ch = h0 + kh - 1;
                        if (a.stride == 2) {
                            mi = ((kh + 1) & 1) * 2 + ((kw + 1) & 1);        // parity of (kh-1, kw-1)
                            cw = (kw - 1) >> 1;                                // floor((kw-1)/2)
                            ch = h0 + ((kh - 1) >> 1);
                        }
                        if (NCTA == 2) {
                            tma_load_4d_2cta(base, &maps.a_hi[mi], cc * 32, cw, ch, n0, fb);
                            tma_load_4d_2cta(base + A_BYTES, &maps.a_lo[mi], cc * 32, cw, ch, n0, fb);
                        } else {
                            tma_load_4d_a(base, &maps.a_hi[mi], cc * 32, cw, ch, n0, fb);
                            tma_load_4d_a(base + A_BYTES, &maps.a_lo[mi], cc * 32, cw, ch, n0, fb);
                        }
                    } else {
                        const int cc = kb - kb3;
                        if (NCTA == 2) {
                            tma_load_4d_2cta(base, &maps.x_hi, cc * 32, 0, h0, n0, fb);
                            tma_load_4d_2cta(base + A_BYTES, &maps.x_lo, cc * 32, 0, h0, n0, fb);
                        } else {
                            tma_load_4d_a(base, &maps.x_hi, cc * 32, 0, h0, n0, fb);
                            tma_load_4d_a(base + A_BYTES, &maps.x_lo, cc * 32, 0, h0, n0, fb);
                        }
                    }
                    if (NCTA == 2) {
                        tma_load_2d_2cta(base + 2 * A_BYTES, &maps.b_hi, kb * 32, brow, fb);
                        tma_load_2d_2cta(base + 2 * A_BYTES + b_bytes, &maps.b_lo, kb * 32, brow, fb);
                    } else {
                        tma_load_2d_a(base + 2 * A_BYTES, &maps.b_hi, kb * 32, brow, fb);
                        tma_load_2d_a(base + 2 * A_BYTES + b_bytes, &maps.b_lo, kb * 32, brow, fb);
                    }
                    if (++st == nstages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (cta_rank == 0 && elect_one()) {
            // ===== MMA issuer (the leader CTA of a pair) =====
            const uint32_t idesc = F16 ? wrn_f16_idesc(128 * NCTA, a.bn_tile) : make_tf32_idesc(128 * NCTA, a.bn_tile);
            uint32_t st = 0, ph = 0, buf = 0, bph = 0, lph = 0;
            const uint32_t nstages = (uint32_t)a.stages;
            for (int t = first_unit; t < total_tiles; t += unit_stride) {
                uint32_t lo_acc = 0;
                if (F16) {                                                               // the previous tile's LO has been drained
                    mbar_wait_a(smem_u32(&lo_empty_bar), lph ^ 1u);
                    lph ^= 1u;
                }
                for (int kb0 = 0, len = a.seg0; kb0 < k_blocks; kb0 += len, len = a.seg) {
                    mbar_wait_a(smem_u32(&tempty_bar[buf]), bph ^ 1u);                   // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * acc_pitch;
                    const int kb1 = kb0 + len < k_blocks ? kb0 + len : k_blocks;
                    uint32_t acc = 0;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_a(smem_u32(&full_bar[st]), ph);
                        tc_fence_after();
                        const uint32_t base = smem_base + st * stage_bytes;
                        const uint64_t d_ahi = make_kmajor_desc<(int)ROW_BYTES>(base), d_alo = make_kmajor_desc<(int)ROW_BYTES>(base + A_BYTES);
                        const uint64_t d_bhi = make_kmajor_desc<(int)ROW_BYTES>(base + 2 * A_BYTES);
                        const uint64_t d_blo = make_kmajor_desc<(int)ROW_BYTES>(base + 2 * A_BYTES + b_bytes);
                        if (F16) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {                                // K = 16 halves = 32 bytes of the 64-byte row
                                const uint64_t koff = (uint64_t)((k * 32) >> 4);
                                wrn_umma_f16<NCTA>(d_tmem, d_ahi + koff, d_bhi + koff, idesc, acc);                       // ACC
                                wrn_umma_f16<NCTA>(tmem_base + lo_col, d_ahi + koff, d_blo + koff, idesc, lo_acc);       // LO
                                acc = 1;
                                lo_acc = 1;
                                wrn_umma_f16<NCTA>(tmem_base + lo_col, d_alo + koff, d_bhi + koff, idesc, 1);
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t koff = (uint64_t)((k * 32) >> 4);
                            if (NCTA == 2) {
                                umma_tf32_2cta(d_tmem, d_alo + koff, d_bhi + koff, idesc, acc);
                                acc = 1;
                                umma_tf32_2cta(d_tmem, d_ahi + koff, d_blo + koff, idesc, 1);
                                umma_tf32_2cta(d_tmem, d_ahi + koff, d_bhi + koff, idesc, 1);
                            } else {
                                umma_tf32(d_tmem, d_alo + koff, d_bhi + koff, idesc, acc);
                                acc = 1;
                                umma_tf32(d_tmem, d_ahi + koff, d_blo + koff, idesc, 1);
                                umma_tf32(d_tmem, d_ahi + koff, d_bhi + koff, idesc, 1);
                            }
                        }
                        }
                        if (NCTA == 2) umma_commit_2cta(smem_u32(&empty_bar[st]), 3);     // frees the stage in BOTH CTAs
                        else umma_commit(smem_u32(&empty_bar[st]));
                        if (++st == nstages) { st = 0; ph ^= 1u; }
                    }
                    if (NCTA == 2) umma_commit_2cta(smem_u32(&tfull_bar[buf]), 3);         // both epilogues drain their half
                    else umma_commit(smem_u32(&tfull_bar[buf]));
                    if (++buf == ntbuf) { buf = 0; bph ^= 1u; }
                }
            }
        }
    } else {
        // ===== epilogue: thread = output pixel (TMEM lane) x one half of the tile's channels (3xTF32: a contiguous half, a
        // multiple of 16 <= 80; FP16-split: the even or the odd 16-column chunks of a tile of <= 128) =====
        const int q = warp & 3, half_id = (warp - 2) >> 2;
        const int half = a.bn_tile >> 1;
        auto chunk_col = [&](int j) { return half_id * half + j * 16; };   // first column within the tile
        auto chunk_on = [&](int j) { return j * 16 < half; };
        const int qi = lane & 3, qb = lane & ~3;
        const int r = q * 32 + qb;                            // first pixel of this lane's quad (4 consecutive pixels of a row)
        const int w = r % WT, h = (r / WT) % HT, nl = r / (WT * HT);
        uint32_t buf = 0, bph = 0;
        for (int t = first_unit; t < total_tiles; t += unit_stride) {
            const int nt = t % a.n_tiles, mt = (t / a.n_tiles) * NCTA + (int)cta_rank;
            int n0, h0;
            if (tpi > 0) { n0 = mt / tpi; h0 = (mt % tpi) * HT; } else { n0 = mt * 2; h0 = 0; }
            const int n = n0 + nl;
            const bool valid = n < a.n_images;
            const int tbase = nt * a.bn_tile;
            const int64_t off = (((int64_t)n * a.hout + (h0 + h)) * a.hout + w) * a.cout + tbase;
            float accr[WRN_EPI_CHUNKS][16];
#pragma unroll
            for (int j = 0; j < WRN_EPI_CHUNKS; ++j)
#pragma unroll
                for (int e = 0; e < 16; ++e) accr[j][e] = 0.f;
            for (int kb0 = 0, len = a.seg0; kb0 < k_blocks; kb0 += len, len = a.seg) {
                mbar_wait_a(smem_u32(&tfull_bar[buf]), bph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_pitch;
#pragma unroll
                for (int j = 0; j < WRN_EPI_CHUNKS; ++j) {
                    if (chunk_on(j)) {
                        uint32_t rr[16];
                        tmem_ld16(taddr + (uint32_t)chunk_col(j), rr);
#pragma unroll
                        for (int e = 0; e < 16; ++e) accr[j][e] += __uint_as_float(rr[e]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (NCTA == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
                    else mbar_arrive(&tempty_bar[buf]);
                }
                if (++buf == ntbuf) { buf = 0; bph ^= 1u; }
            }
            if (F16) {
                // the commit of the tile's last segment also covers every LO MMA: add the cross terms and free the LO accumulator
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + lo_col;
#pragma unroll
                for (int j = 0; j < WRN_EPI_CHUNKS; ++j) {
                    if (chunk_on(j)) {
                        uint32_t rl[16];
                        tmem_ld16(taddr + (uint32_t)chunk_col(j), rl);
#pragma unroll
                        for (int e = 0; e < 16; ++e) accr[j][e] = fmaf(__uint_as_float(rl[e]), kWrnLoUnscale, accr[j][e]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (NCTA == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&lo_empty_bar), 0));
                    else mbar_arrive(&lo_empty_bar);
                }
            }
            // The TMEM layout gives a thread one pixel (row) x 16 columns: stored as is, every warp-level access touches 32
            // different 128-byte lines (32 L1 tag cycles each -- measured: the epilogue then takes ~3/4 of a tile's MMA time and
            // the MMA issuer stalls on it).  A 4 x 4 transpose inside each lane quad makes lane (qb + qi) own columns
            // [4 qi, 4 qi + 4) of the quad's 4 rows, so a quad reads / writes 64 contiguous bytes of one pixel per access.
#pragma unroll
            for (int j = 0; j < WRN_EPI_CHUNKS; ++j) {
                if (!chunk_on(j)) continue;
                float tr[4][4];
#pragma unroll
                for (int rd = 0; rd < 4; ++rd) {
                    const int sb = (qi + rd) & 3, jr = (qi - rd) & 3;       // block sent / quad row received in this round
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float sv = sb == 0 ? accr[j][e] : (sb == 1 ? accr[j][4 + e] : (sb == 2 ? accr[j][8 + e] : accr[j][12 + e]));
                        const float rv = rd == 0 ? sv : __shfl_sync(0xffffffffu, sv, qb + jr);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (jr == jj) tr[jj][e] = rv;
                    }
                }
                const int c0 = chunk_col(j) + 4 * qi;
                const int cbase = tbase;
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.bias + cbase + c0));
                    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = a4;
                    if (a.out_hi) {
                        a4 = __ldg(reinterpret_cast<const float4 *>(a.bn + cbase + c0));
                        s4 = __ldg(reinterpret_cast<const float4 *>(a.bn + a.cout + cbase + c0));
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int64_t o = off + (int64_t)jj * a.cout + c0;
                        float4 v = make_float4(tr[jj][0] + b4.x, tr[jj][1] + b4.y, tr[jj][2] + b4.z, tr[jj][3] + b4.w);
                        if (a.res) {
                            const float4 t4 = __ldg(reinterpret_cast<const float4 *>(a.res + o));
                            v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
                        }
                        if (a.out_raw) *reinterpret_cast<float4 *>(a.out_raw + o) = v;
                        if (a.stats) {
                            s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
                            s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
                            s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
                        }
                        if (a.outx_hi) {
                            if (F16) {
                                uint2 hh, ll;
                                wrn_split_h4(v.x, v.y, v.z, v.w, hh, ll);
                                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(a.outx_hi) + o) = hh;
                                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(a.outx_lo) + o) = ll;
                            } else {
                            float4 hv, lv;
                            hv.x = rn_tf32(v.x); hv.y = rn_tf32(v.y); hv.z = rn_tf32(v.z); hv.w = rn_tf32(v.w);
                            lv.x = rn_tf32(v.x - hv.x); lv.y = rn_tf32(v.y - hv.y); lv.z = rn_tf32(v.z - hv.z); lv.w = rn_tf32(v.w - hv.w);
                            *reinterpret_cast<float4 *>(a.outx_hi + o) = hv;
                            *reinterpret_cast<float4 *>(a.outx_lo + o) = lv;
                            }
                        }
                        if (a.out_hi) {
                            const float y0 = relu_nan(fmaf(a4.x, v.x, s4.x)), y1 = relu_nan(fmaf(a4.y, v.y, s4.y));
                            const float y2 = relu_nan(fmaf(a4.z, v.z, s4.z)), y3 = relu_nan(fmaf(a4.w, v.w, s4.w));
                            if (F16) {
                                uint2 hh, ll;
                                wrn_split_h4(y0, y1, y2, y3, hh, ll);
                                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(a.out_hi) + o) = hh;
                                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(a.out_lo) + o) = ll;
                                continue;
                            }
                            float4 hv, lv;
                            hv.x = rn_tf32(y0); hv.y = rn_tf32(y1); hv.z = rn_tf32(y2); hv.w = rn_tf32(y3);
                            lv.x = rn_tf32(y0 - hv.x); lv.y = rn_tf32(y1 - hv.y); lv.z = rn_tf32(y2 - hv.z); lv.w = rn_tf32(y3 - hv.w);
                            *reinterpret_cast<float4 *>(a.out_hi + o) = hv;
                            *reinterpret_cast<float4 *>(a.out_lo + o) = lv;
                        }
                    }
                }
                if (a.stats) {
                    // a warp's 32 pixels belong to one image, hence to one batch: reduce over the 8 quads, then one fp64 atomic
                    // per (channel, moment) from the lanes of quad 0
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) {
                            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
                            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
                        }
                    }
                    if (qb == 0 && valid) {
                        double *sp = a.stats + (int64_t)(n / a.batch) * 2 * a.cout + cbase + c0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            atomicAdd(sp + e, (double)s1[e]);
                            atomicAdd(sp + a.cout + e, (double)s2[e]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();                        // nobody frees TMEM / exits while the peer may still signal it
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (NCTA == 2) tmem_dealloc_2cta(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ---- pack / fold (per sample) --------------------------------------------------------------------------------------
// filters [co][ci][taps] (PyTorch) -> K-major rows dst[co][koff + tap * cin_p + ci], ci >= cin zero-filled, TF32 hi / lo
template <bool F16>
__global__ void __launch_bounds__(256) wrn_pack_filter_kernel(const float *__restrict__ src, float *__restrict__ dhi,
                                                              float *__restrict__ dlo, int cin, int cin_p, int cout, int taps,
                                                              int ktot, int koff) {
    const int64_t per_row = (int64_t)taps * cin_p, total = per_row * cout;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int co = (int)(i / per_row);
        const int rem = (int)(i - (int64_t)co * per_row);
        const int tap = rem / cin_p, ci = rem - tap * cin_p;
        const float wv = ci < cin ? __ldg(src + ((int64_t)co * cin + ci) * taps + tap) : 0.f;
        const int64_t d = (int64_t)co * ktot + koff + rem;
        if (F16) {                                             // the same regions of `packed`, viewed as halves
            const __half hh = __float2half_rn(wv);
            reinterpret_cast<__half *>(dhi)[d] = hh;
            reinterpret_cast<__half *>(dlo)[d] = __float2half_rn((wv - __half2float(hh)) * kWrnLoScale);
        } else {
            const float hv = rn_tf32(wv);
            dhi[d] = hv;
            dlo[d] = rn_tf32(wv - hv);
        }
    }
}

// eval-mode BN -> (a, b): y = a x + b, a = gamma / sqrt(var + eps), b = beta - mean a
__global__ void wrn_bn_fold_kernel(const float *__restrict__ gamma, const float *__restrict__ beta,
                                   const float *__restrict__ mean, const float *__restrict__ var, int c, float *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c; i += gridDim.x * blockDim.x) {
        const float av = gamma[i] / sqrtf(var[i] + 1e-5f);
        dst[i] = av;
        dst[c + i] = beta[i] - mean[i] * av;
    }
}

__global__ void wrn_bias_kernel(const float *__restrict__ b0, const float *__restrict__ b1, int c, float *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c; i += gridDim.x * blockDim.x)
        dst[i] = b1 ? b0[i] + b1[i] : b0[i];
}

// ---- stem: R0 = conv3x3(x) + bias (3 -> 16); A = split(relu(bn(R0))), X = split(R0), both padded to 32 channels ------
// Train mode (bn == null): no A planes; the raw 16-channel output goes to raw16 and its per-batch sums to stats.
template <bool F16>
__global__ void __launch_bounds__(256) wrn_stem_kernel(const float *__restrict__ x, const float *__restrict__ wsrc,
                                                       const float *__restrict__ bsrc, const float *__restrict__ bn,
                                                       float *__restrict__ a_hi, float *__restrict__ a_lo,
                                                       float *__restrict__ x_hi, float *__restrict__ x_lo,
                                                       float *__restrict__ raw16, double *__restrict__ stats, int batch) {
    // one CTA per image; thread = 4 consecutive pixels of a row x all 16 output channels
    __shared__ float xs[3][34][35];
    __shared__ __align__(16) float ws[27 * 16];
    __shared__ float bs[16], bns[32];
    const int n = blockIdx.x;
    for (int i = threadIdx.x; i < 27 * 16; i += 256) {                 // ws[(ci * 9 + tap) * 16 + co] <- w[co][ci][tap]
        const int co = i & 15, ct = i >> 4;
        ws[i] = __ldg(wsrc + co * 27 + ct);
    }
    if (threadIdx.x < 16) bs[threadIdx.x] = __ldg(bsrc + threadIdx.x);
    if (threadIdx.x < 32) bns[threadIdx.x] = bn ? __ldg(bn + threadIdx.x) : 0.f;
    for (int i = threadIdx.x; i < 3 * 34 * 34; i += 256) {
        const int ww = i % 34, hh = (i / 34) % 34, ci = i / (34 * 34);
        const int hi = hh - 1, wi = ww - 1;
        xs[ci][hh][ww] = (hi >= 0 && hi < 32 && wi >= 0 && wi < 32) ? __ldg(x + ((int64_t)n * 3 + ci) * 1024 + hi * 32 + wi) : 0.f;
    }
    __syncthreads();
    const int h = threadIdx.x >> 3, w0 = (threadIdx.x & 7) * 4;
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[p][c] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            float xv[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) xv[d] = xs[ci][h + kh][w0 + d];
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float4 *wp = reinterpret_cast<const float4 *>(ws + (ci * 9 + kh * 3 + kw) * 16);
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const float4 w4 = wp[qd];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        acc[p][4 * qd + 0] = fmaf(xv[p + kw], w4.x, acc[p][4 * qd + 0]);
                        acc[p][4 * qd + 1] = fmaf(xv[p + kw], w4.y, acc[p][4 * qd + 1]);
                        acc[p][4 * qd + 2] = fmaf(xv[p + kw], w4.z, acc[p][4 * qd + 2]);
                        acc[p][4 * qd + 3] = fmaf(xv[p + kw], w4.w, acc[p][4 * qd + 3]);
                    }
                }
            }
        }
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float st1[16], st2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) st1[c] = st2[c] = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int64_t off = ((int64_t)n * 1024 + h * 32 + w0 + p) * 32;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float r[4], y[4], rh[4], rl[4], yh[4], yl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = acc[p][i + k] + bs[i + k];
                y[k] = relu_nan(fmaf(bns[i + k], r[k], bns[16 + i + k]));
                rh[k] = rn_tf32(r[k]); rl[k] = rn_tf32(r[k] - rh[k]);
                yh[k] = rn_tf32(y[k]); yl[k] = rn_tf32(y[k] - yh[k]);
                st1[i + k] += r[k];
                st2[i + k] = fmaf(r[k], r[k], st2[i + k]);
            }
            if (F16) {                                     // A and X planes as halves
                uint2 hh, ll;
                const uint2 z2 = make_uint2(0u, 0u);
                __half *ah = reinterpret_cast<__half *>(a_hi), *al = reinterpret_cast<__half *>(a_lo);
                __half *xh = reinterpret_cast<__half *>(x_hi), *xl = reinterpret_cast<__half *>(x_lo);
                if (bn) {
                    wrn_split_h4(y[0], y[1], y[2], y[3], hh, ll);
                    *reinterpret_cast<uint2 *>(ah + off + i) = hh;
                    *reinterpret_cast<uint2 *>(al + off + i) = ll;
                    *reinterpret_cast<uint2 *>(ah + off + 16 + i) = z2;
                    *reinterpret_cast<uint2 *>(al + off + 16 + i) = z2;
                } else {
                    *reinterpret_cast<float4 *>(raw16 + (off >> 1) + i) = make_float4(r[0], r[1], r[2], r[3]);
                }
                wrn_split_h4(r[0], r[1], r[2], r[3], hh, ll);
                *reinterpret_cast<uint2 *>(xh + off + i) = hh;
                *reinterpret_cast<uint2 *>(xl + off + i) = ll;
                *reinterpret_cast<uint2 *>(xh + off + 16 + i) = z2;
                *reinterpret_cast<uint2 *>(xl + off + 16 + i) = z2;
                continue;
            }
            if (bn) {
                *reinterpret_cast<float4 *>(a_hi + off + i) = make_float4(yh[0], yh[1], yh[2], yh[3]);
                *reinterpret_cast<float4 *>(a_lo + off + i) = make_float4(yl[0], yl[1], yl[2], yl[3]);
                *reinterpret_cast<float4 *>(a_hi + off + 16 + i) = z4;
                *reinterpret_cast<float4 *>(a_lo + off + 16 + i) = z4;
            } else {
                *reinterpret_cast<float4 *>(raw16 + (off >> 1) + i) = make_float4(r[0], r[1], r[2], r[3]);
            }
            *reinterpret_cast<float4 *>(x_hi + off + i) = make_float4(rh[0], rh[1], rh[2], rh[3]);
            *reinterpret_cast<float4 *>(x_lo + off + i) = make_float4(rl[0], rl[1], rl[2], rl[3]);
            *reinterpret_cast<float4 *>(x_hi + off + 16 + i) = z4;
            *reinterpret_cast<float4 *>(x_lo + off + 16 + i) = z4;
        }
    }
    if (stats) {                                           // per-image sums -> one fp64 atomic per (channel, moment)
        __shared__ float red[8][32];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                st1[c] += __shfl_xor_sync(0xffffffffu, st1[c], o);
                st2[c] += __shfl_xor_sync(0xffffffffu, st2[c], o);
            }
            if (lane == 0) { red[warp][c] = st1[c]; red[warp][16 + c] = st2[c]; }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            float t = 0.f;
#pragma unroll
            for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
            atomicAdd(stats + (int64_t)(n / batch) * 32 + threadIdx.x, (double)t);      // [batch][2][16]
        }
    }
}

// ---- head: relu(bn(R)) -> 8x8 average pool -> linear; one CTA per image -------------------------------------------------
// proba_sum != nullptr: the BMA epilogue is fused -- the image's logits stay in shared memory and warp 0 adds softmax and the
// smoothed entropy to the image's row of the accumulators (softmax_accumulate_row, the arithmetic of ursa_bma_accumulate).
template <int PER_LANE>
__global__ void __launch_bounds__(256) wrn_head_kernel(const float *__restrict__ act, const float *__restrict__ bn, int cf,
                                                       const float *__restrict__ lw, const float *__restrict__ lb, int C,
                                                       float *__restrict__ logits, float *__restrict__ proba_sum,
                                                       float *__restrict__ entropy_sum, float one_minus_gamma,
                                                       float gamma_over_c) {
    extern __shared__ float feat[];                                    // [cf] then [C] logits
    const int n = blockIdx.x;
    const float *xp = act + (int64_t)n * 64 * cf;
    for (int c = threadIdx.x; c < cf; c += 256) {
        const float av = __ldg(bn + c), bv = __ldg(bn + cf + c);
        float f = 0.f;
        for (int px = 0; px < 64; ++px) f += relu_nan(fmaf(av, __ldg(xp + (int64_t)px * cf + c), bv));
        feat[c] = f * (1.f / 64.f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < C; c += 8) {
        float acc = 0.f;
        for (int k = lane; k < cf; k += 32) acc = fmaf(feat[k], __ldg(lw + (int64_t)c * cf + k), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float v = acc + __ldg(lb + c);
            if (logits != nullptr) logits[(int64_t)n * C + c] = v;
            feat[cf + c] = v;
        }
    }
    if (proba_sum == nullptr) return;
    __syncthreads();
    if (warp == 0) {
        float lg[PER_LANE], P[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int c = lane + 32 * j;
            lg[j] = c < C ? feat[cf + c] : 0.f;
            P[j] = c < C ? proba_sum[(int64_t)n * C + c] : 0.f;
        }
        float E = entropy_sum[n];
        softmax_accumulate_row<PER_LANE>(lg, C, lane, one_minus_gamma, gamma_over_c, P, E);
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j)
            if (lane + 32 * j < C) proba_sum[(int64_t)n * C + lane + 32 * j] = P[j];
        if (lane == 0) entropy_sum[n] = E;
    }
}

// ---- train-mode BatchNorm (re-estimation of the running statistics, reference util.py:212-247) ---------------------------
// stats: [nb][2][C] fp64 sums over the batch's pixels.  One thread per channel walks the chunk's batches in order: batch
// statistics -> (a, b) for the normalisation of THIS batch, running statistics with the cumulative momentum b / (n + b)
// (util.py:239-241) and PyTorch's unbiased running variance.  The sums are cleared for the next layer.
__global__ void wrn_bn_finalize_kernel(double *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                                       int C, int hw, int nc, int batch, int64_t n_before, float *__restrict__ run_mean,
                                       float *__restrict__ run_var, float *__restrict__ ab) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float rm = run_mean[c], rv = run_var[c];
    int64_t n = n_before;
    if (n == 0) { rm = 0.f; rv = 1.f; }                    // reset_bn (util.py:196-199)
    const int nb = (nc + batch - 1) / batch;
    for (int j = 0; j < nb; ++j) {
        const int bj = nc - j * batch < batch ? nc - j * batch : batch;
        const double cnt = (double)bj * hw;
        double *sp = stats + (int64_t)j * 2 * C;
        const double mean = sp[c] / cnt;
        double var = sp[C + c] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        sp[c] = 0.0;
        sp[C + c] = 0.0;
        const float av = gamma[c] / sqrtf((float)var + 1e-5f);
        ab[(int64_t)j * 2 * C + c] = av;
        ab[(int64_t)j * 2 * C + C + c] = beta[c] - (float)mean * av;
        const float mom = (float)((double)bj / (double)(n + bj));
        const float unbiased = (float)(cnt > 1.0 ? var * cnt / (cnt - 1.0) : var);
        rm = (1.f - mom) * rm + mom * (float)mean;
        rv = (1.f - mom) * rv + mom * unbiased;
        n += bj;
    }
    run_mean[c] = rm;
    run_var[c] = rv;
}

// raw [P][hw][C] -> A planes split(relu(a_j v + b_j)) [P][hw][c_pad] (channels >= C zero) and, optionally, X planes split(v)
template <bool F16>
__global__ void __launch_bounds__(256) wrn_bn_apply_kernel(const float *__restrict__ raw, const float *__restrict__ ab, int C,
                                                           int c_pad, int hw, int batch, int64_t total4, float *__restrict__ a_hi,
                                                           float *__restrict__ a_lo, float *__restrict__ x_hi,
                                                           float *__restrict__ x_lo) {
    const int cp4 = c_pad >> 2;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += (int64_t)gridDim.x * 256) {
        const int c = (int)(i % cp4) * 4;
        const int64_t px = i / cp4;
        const int j = (int)(px / hw) / batch;
        float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), lv = hv, xh = hv, xl = hv;
        if (F16) {                                             // the same planes as halves
            uint2 yh2 = make_uint2(0u, 0u), yl2 = yh2, xh2 = yh2, xl2 = yh2;
            if (c < C) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(raw + px * C + c));
                const float4 a4 = __ldg(reinterpret_cast<const float4 *>(ab + (int64_t)j * 2 * C + c));
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(ab + (int64_t)j * 2 * C + C + c));
                wrn_split_h4(relu_nan(fmaf(a4.x, v.x, b4.x)), relu_nan(fmaf(a4.y, v.y, b4.y)), relu_nan(fmaf(a4.z, v.z, b4.z)),
                             relu_nan(fmaf(a4.w, v.w, b4.w)), yh2, yl2);
                wrn_split_h4(v.x, v.y, v.z, v.w, xh2, xl2);
            }
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(a_hi) + px * c_pad + c) = yh2;
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(a_lo) + px * c_pad + c) = yl2;
            if (x_hi) {
                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(x_hi) + px * c_pad + c) = xh2;
                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(x_lo) + px * c_pad + c) = xl2;
            }
            continue;
        }
        if (c < C) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(raw + px * C + c));
            const float4 a4 = __ldg(reinterpret_cast<const float4 *>(ab + (int64_t)j * 2 * C + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(ab + (int64_t)j * 2 * C + C + c));
            const float y0 = relu_nan(fmaf(a4.x, v.x, b4.x)), y1 = relu_nan(fmaf(a4.y, v.y, b4.y));
            const float y2 = relu_nan(fmaf(a4.z, v.z, b4.z)), y3 = relu_nan(fmaf(a4.w, v.w, b4.w));
            hv.x = rn_tf32(y0); hv.y = rn_tf32(y1); hv.z = rn_tf32(y2); hv.w = rn_tf32(y3);
            lv.x = rn_tf32(y0 - hv.x); lv.y = rn_tf32(y1 - hv.y); lv.z = rn_tf32(y2 - hv.z); lv.w = rn_tf32(y3 - hv.w);
            xh.x = rn_tf32(v.x); xh.y = rn_tf32(v.y); xh.z = rn_tf32(v.z); xh.w = rn_tf32(v.w);
            xl.x = rn_tf32(v.x - xh.x); xl.y = rn_tf32(v.y - xh.y); xl.z = rn_tf32(v.z - xh.z); xl.w = rn_tf32(v.w - xh.w);
        }
        *reinterpret_cast<float4 *>(a_hi + px * c_pad + c) = hv;
        *reinterpret_cast<float4 *>(a_lo + px * c_pad + c) = lv;
        if (x_hi) {
            *reinterpret_cast<float4 *>(x_hi + px * c_pad + c) = xh;
            *reinterpret_cast<float4 *>(x_lo + px * c_pad + c) = xl;
        }
    }
}

// ---- plan ------------------------------------------------------------------------------------------------------------
struct WrnBlock {
    int cin, cin_p, cout, stride;
    bool transition;
    // bank-row offsets (model.parameters() order: bn1.w, bn1.b, conv1.w, conv1.b, bn2.w, bn2.b, conv2.w, conv2.b, [sc.w, sc.b])
    int64_t bn1_w, bn1_b, c1_w, c1_b, bn2_w, bn2_b, c2_w, c2_b, sc_w, sc_b;
    int64_t bn1_buf, bn2_buf;          // buffer-row offsets of running_mean (running_var follows at + channels)
    // packed offsets (floats)
    int64_t p_w1_hi, p_w1_lo, p_w2_hi, p_w2_lo, p_bias1, p_bias2, p_bn1, p_bn2;
    int k1, k2;                        // K extents of the packed filter rows
};

struct WrnPlan {
    int n, k, C, widths[4];
    WrnBlock blocks[3][WRN_MAX_BLOCKS];
    int64_t conv1_w, conv1_b, bnf_w, bnf_b, bnf_buf, lin_w, lin_b;
    int64_t p_bnf;
    int64_t packed_floats, D, NB;
};

static bool wrn_build_plan(int depth, int widen, int C, WrnPlan &pl) {
    if (depth < 10 || (depth - 4) % 6 != 0 || widen < 2 || widen % 2 != 0 || widen > 16 || C < 1) return false;
    const int n = (depth - 4) / 6;
    if (n > WRN_MAX_BLOCKS) return false;
    pl.n = n; pl.k = widen; pl.C = C;
    pl.widths[0] = 16; pl.widths[1] = 16 * widen; pl.widths[2] = 32 * widen; pl.widths[3] = 64 * widen;
    int64_t src = 0, buf = 0, dst = 0;
    auto take = [&](int64_t &cursor, int64_t cnt) { const int64_t o = cursor; cursor += cnt; return o; };
    auto take_dst = [&](int64_t cnt) { const int64_t o = dst; dst += (cnt + 255) & ~(int64_t)255; return o; };   // 1 KB aligned
    pl.conv1_w = take(src, 16 * 27);
    pl.conv1_b = take(src, 16);
    int inpl = 16;
    const int strides[3] = {1, 2, 2};
    for (int g = 0; g < 3; ++g)
        for (int b = 0; b < n; ++b) {
            WrnBlock &B = pl.blocks[g][b];
            B.cin = inpl; B.cin_p = (inpl + 31) & ~31; B.cout = pl.widths[g + 1];
            B.stride = b == 0 ? strides[g] : 1;
            B.transition = B.stride != 1 || B.cin != B.cout;
            B.bn1_w = take(src, B.cin); B.bn1_b = take(src, B.cin);
            B.c1_w = take(src, (int64_t)B.cout * B.cin * 9); B.c1_b = take(src, B.cout);
            B.bn2_w = take(src, B.cout); B.bn2_b = take(src, B.cout);
            B.c2_w = take(src, (int64_t)B.cout * B.cout * 9); B.c2_b = take(src, B.cout);
            B.sc_w = B.sc_b = -1;
            if (B.transition) { B.sc_w = take(src, (int64_t)B.cout * B.cin); B.sc_b = take(src, B.cout); }
            B.bn1_buf = take(buf, 2 * B.cin);
            B.bn2_buf = take(buf, 2 * B.cout);
            B.k1 = 9 * B.cin_p;
            B.k2 = 9 * B.cout + (B.transition ? B.cin_p : 0);
            B.p_w1_hi = take_dst((int64_t)B.cout * B.k1); B.p_w1_lo = take_dst((int64_t)B.cout * B.k1);
            B.p_w2_hi = take_dst((int64_t)B.cout * B.k2); B.p_w2_lo = take_dst((int64_t)B.cout * B.k2);
            B.p_bias1 = take_dst(B.cout); B.p_bias2 = take_dst(B.cout);
            B.p_bn1 = take_dst(2 * B.cin); B.p_bn2 = take_dst(2 * B.cout);
            inpl = B.cout;
        }
    pl.bnf_w = take(src, inpl); pl.bnf_b = take(src, inpl);
    pl.bnf_buf = take(buf, 2 * inpl);
    pl.lin_w = take(src, (int64_t)C * inpl); pl.lin_b = take(src, C);
    pl.p_bnf = take_dst(2 * inpl);
    pl.packed_floats = dst;
    pl.D = src; pl.NB = buf;
    return true;
}

struct WrnChunking {
    int nc;
    size_t unit_bytes, packed_bytes, logit_bytes, total;
};

static WrnChunking wrn_chunking(int64_t N, const WrnPlan &pl) {
    WrnChunking c;
    c.nc = (int)(N < WRN_CHUNK_IMAGES ? N : WRN_CHUNK_IMAGES);
    // unit = the largest activation plane: conv1 output of block (2, 0) at 32 x 32 x 32k channels
    c.unit_bytes = (((size_t)c.nc * 1024 * pl.widths[2] * sizeof(float)) + 1023) & ~(size_t)1023;
    c.packed_bytes = (((size_t)pl.packed_floats * sizeof(float)) + 1023) & ~(size_t)1023;
    c.logit_bytes = ((((size_t)c.nc * pl.C * sizeof(float)) + 1023) & ~(size_t)1023);
    // A1 hi/lo, A2 hi/lo (4 units), Xa hi/lo, Xb hi/lo, Ra, Rb (6 half units: <= 32 x 32 x 16k channels)
    c.total = 7 * c.unit_bytes + c.packed_bytes + c.logit_bytes + 2048;
    return c;
}

// plane [N][H][H][C]: stride-1 map, or the (hp, wp) parity sub-lattice for a stride-2 consumer; box = one 128-pixel tile
static int wrn_act_map(CUtensorMap *tm, const float *plane, int nc, int H, int C, int hout, int stride, int parity, bool f16) {
    const int WT = hout, HT = hout >= 16 ? 128 / hout : hout, NT = 128 / (WT * HT);
    const uint32_t box[4] = {32u, (uint32_t)WT, (uint32_t)HT, (uint32_t)NT};
    const uint64_t es = f16 ? 2 : 4;                           // FP16-split planes: halves, 64-byte swizzle rows
    const int swz = f16 ? 64 : 128;
    if (stride == 1) {
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)H, (uint64_t)H, (uint64_t)nc};
        const uint64_t st[3] = {(uint64_t)C * es, (uint64_t)H * C * es, (uint64_t)H * H * C * es};
        return make_tensor_map_t(tm, plane, 4, dims, st, box, swz, f16 ? 1 : 0);
    }
    const int hp = parity >> 1, wp = parity & 1;
    const char *base = reinterpret_cast<const char *>(plane) + ((int64_t)hp * H + wp) * C * es;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)(H / 2), (uint64_t)(H / 2), (uint64_t)nc};
    const uint64_t st[3] = {(uint64_t)2 * C * es, (uint64_t)2 * H * C * es, (uint64_t)H * H * C * es};
    return make_tensor_map_t(tm, base, 4, dims, st, box, swz, f16 ? 1 : 0);
}

static int wrn_pick_bn_tile(int cout, bool f16) {
    for (int nt = 1; nt <= 16; ++nt) {
        if (cout % nt != 0) continue;
        const int bn = cout / nt;
        (void)f16;                                             // both engines: <= 160 columns (FP16-split: 3 x 160 TMEM columns)
        if (bn <= 160 && bn % 32 == 0) return bn;              // two 16-column-granular halves
    }
    return 0;
}

// Can a CTA pair of this kernel be co-scheduled on this device / partition at all?  (asked once per thread and device)
static bool wrn_pairs_supported() {
    static thread_local int cached = -1, cached_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cached < 0 || dev != cached_dev) {
        cached_dev = dev;
        cached = 0;
        const size_t smem = 4 * (2 * (size_t)WRN_A_BYTES + 2 * (size_t)80 * 128) + 1024;        // 4 stages at N = 160
        if (cudaFuncSetAttribute(wrn_conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2);
            cfg.blockDim = dim3(WRN_THREADS);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, wrn_conv_tc_kernel<2>, &cfg) == cudaSuccess && n >= 1) cached = 1;
        }
        (void)cudaGetLastError();
    }
    return cached == 1;
}

// one conv launch: A planes [nc][hin][hin][cin_p] -> cout channels at hout = hin / stride; xin_p > 0 folds a 1x1 stride-`stride`
// conv over the X planes [nc][hin][hin][xin_p] into the accumulation
static int wrn_launch_conv(const float *a_hi, const float *a_lo, int hin, int cin_p, const float *x_hi, const float *x_lo,
                           int xin_p, int cout, int stride, int nc, const float *b_hi, const float *b_lo, int ktot,
                           WrnConvArgs g, cudaStream_t st, bool f16 = false) {
    const int hout = hin / stride;
    WrnMaps maps;
    const int nmaps = stride == 2 ? 4 : 1;
    for (int i = 0; i < nmaps; ++i) {
        if (int rc = wrn_act_map(&maps.a_hi[i], a_hi, nc, hin, cin_p, hout, stride, i, f16)) return rc;
        if (int rc = wrn_act_map(&maps.a_lo[i], a_lo, nc, hin, cin_p, hout, stride, i, f16)) return rc;
    }
    for (int i = nmaps; i < 4; ++i) { maps.a_hi[i] = maps.a_hi[0]; maps.a_lo[i] = maps.a_lo[0]; }
    if (xin_p > 0) {
        if (int rc = wrn_act_map(&maps.x_hi, x_hi, nc, hin, xin_p, hout, stride, 0, f16)) return rc;
        if (int rc = wrn_act_map(&maps.x_lo, x_lo, nc, hin, xin_p, hout, stride, 0, f16)) return rc;
    } else {
        maps.x_hi = maps.a_hi[0];
        maps.x_lo = maps.a_lo[0];
    }
    const int bn_tile = wrn_pick_bn_tile(cout, f16);
    URSA_REQUIRE(bn_tile > 0, "ursa_bma_wrn_forward: no output-channel tile for cout = %d", cout);
    int ncta = 2;                                                       // CTA pairs (cta_group::2); URSA_WRN_2CTA=0 selects single CTAs
    if (const char *e = getenv("URSA_WRN_2CTA")) ncta = atoi(e) == 0 ? 1 : 2;
    if (ncta == 2 && !wrn_pairs_supported()) ncta = 1;
    {
        const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)cout};
        const uint64_t sb[1] = {(uint64_t)ktot * (f16 ? 2 : 4)};
        const uint32_t box[2] = {32u, (uint32_t)(bn_tile / ncta)};
        if (int rc = make_tensor_map_t(&maps.b_hi, b_hi, 2, dims, sb, box, f16 ? 64 : 128, f16 ? 1 : 0)) return rc;
        if (int rc = make_tensor_map_t(&maps.b_lo, b_lo, 2, dims, sb, box, f16 ? 64 : 128, f16 ? 1 : 0)) return rc;
    }
    g.cout = cout; g.bn_tile = bn_tile; g.hout = hout; g.stride = stride; g.n_images = nc;
    g.kchunks = cin_p / 32; g.xchunks = xin_p / 32;
    const int tpi = (hout * hout) / 128;
    g.m_tiles = tpi > 0 ? nc * tpi : (nc + 1) / 2;
    g.n_tiles = cout / bn_tile;
    const size_t stage_bytes = f16 ? 2 * (size_t)WRN_A_BYTES_H + 2 * (size_t)(bn_tile / ncta) * 64
                                   : 2 * (size_t)WRN_A_BYTES + 2 * (size_t)(bn_tile / ncta) * 128;
    int stages = (int)(((size_t)(226 << 10) - 1024) / stage_bytes);
    if (stages > (f16 ? WRN_MAX_STAGES_H : WRN_MAX_STAGES)) stages = f16 ? WRN_MAX_STAGES_H : WRN_MAX_STAGES;
    if (const char *e = getenv("URSA_WRN_STAGES")) { const int v = atoi(e); if (v >= 1 && v < stages) stages = v; }
    g.stages = stages;
    // FP16-split: a K block is 6 MMAs of 80 clk instead of 12, and the ACC chain holds 2 MMAs per block: 24-block segments keep the
    // drains off the critical path.  WRN-28-10, 160-column tiles, SEG 8 / 16 / 24 / 32 / 48: 308 / 332 / 352 / 358 / 364 TFLOP/s at
    // 3.1 / 3.2 / 2.7 / 3.7 / 7.1 e-6 from an fp64 forward (PyTorch fp32: 4.0e-6)
    g.seg = f16 ? 6 * WRN_SEG : WRN_SEG;
    if (const char *e = getenv("URSA_WRN_SEG")) { const int v = atoi(e); if (v >= 1) g.seg = v; }      // accuracy / speed experiments
    g.seg0 = WRN_SEG0_FACTOR * g.seg;                        // ... but never more than a quarter of the K extent
    if (g.seg0 > (g.kchunks * 9 + g.xchunks) / 4) g.seg0 = (g.kchunks * 9 + g.xchunks) / 4;
    if (g.seg0 < g.seg) g.seg0 = g.seg;
    if (const char *e = getenv("URSA_WRN_SEG0")) { const int v = atoi(e); if (v >= 1) g.seg0 = v; }
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    if (ncta == 2) {
        const int units = ((g.m_tiles + 1) / 2) * g.n_tiles, pairs = sm_count() / 2;
        if (f16) URSA_CUDA(cudaFuncSetAttribute(wrn_conv_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else URSA_CUDA(cudaFuncSetAttribute(wrn_conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (units < pairs ? units : pairs));
        cfg.blockDim = dim3(WRN_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (f16) URSA_CUDA(cudaLaunchKernelEx(&cfg, wrn_conv_tc_kernel<2, true>, maps, g));
        else URSA_CUDA(cudaLaunchKernelEx(&cfg, wrn_conv_tc_kernel<2>, maps, g));
        URSA_LAUNCH_CHECK("wrn_conv_tc_kernel<2>");
        return URSA_OK;
    }
    const int tiles = g.m_tiles * g.n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    if (f16) {
        URSA_CUDA(cudaFuncSetAttribute(wrn_conv_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wrn_conv_tc_kernel<1, true><<<grid, WRN_THREADS, smem, st>>>(maps, g);
        URSA_LAUNCH_CHECK("wrn_conv_tc_kernel<1, f16>");
        return URSA_OK;
    }
    URSA_CUDA(cudaFuncSetAttribute(wrn_conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wrn_conv_tc_kernel<1><<<grid, WRN_THREADS, smem, st>>>(maps, g);
    URSA_LAUNCH_CHECK("wrn_conv_tc_kernel");
    return URSA_OK;
}

static int wrn_pack_sample(const WrnPlan &pl, const float *row, const float *brow, float *packed, cudaStream_t st,
                           bool fold_bn = true, bool f16 = false) {
    auto pack = [&](const float *src, float *dhi, float *dlo, int cin, int cin_p, int cout, int taps, int ktot, int koff) {
        const int64_t total = (int64_t)cout * taps * cin_p;
        int64_t blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (f16) wrn_pack_filter_kernel<true><<<(int)blocks, 256, 0, st>>>(src, dhi, dlo, cin, cin_p, cout, taps, ktot, koff);
        else wrn_pack_filter_kernel<false><<<(int)blocks, 256, 0, st>>>(src, dhi, dlo, cin, cin_p, cout, taps, ktot, koff);
    };
    for (int g = 0; g < 3; ++g)
        for (int b = 0; b < pl.n; ++b) {
            const WrnBlock &B = pl.blocks[g][b];
            pack(row + B.c1_w, packed + B.p_w1_hi, packed + B.p_w1_lo, B.cin, B.cin_p, B.cout, 9, B.k1, 0);
            pack(row + B.c2_w, packed + B.p_w2_hi, packed + B.p_w2_lo, B.cout, B.cout, B.cout, 9, B.k2, 0);
            if (B.transition)
                pack(row + B.sc_w, packed + B.p_w2_hi, packed + B.p_w2_lo, B.cin, B.cin_p, B.cout, 1, B.k2, 9 * B.cout);
            wrn_bias_kernel<<<1, 256, 0, st>>>(row + B.c1_b, nullptr, B.cout, packed + B.p_bias1);
            wrn_bias_kernel<<<1, 256, 0, st>>>(row + B.c2_b, B.transition ? row + B.sc_b : nullptr, B.cout, packed + B.p_bias2);
            if (!fold_bn) continue;
            wrn_bn_fold_kernel<<<1, 256, 0, st>>>(row + B.bn1_w, row + B.bn1_b, brow + B.bn1_buf, brow + B.bn1_buf + B.cin, B.cin,
                                                  packed + B.p_bn1);
            wrn_bn_fold_kernel<<<1, 256, 0, st>>>(row + B.bn2_w, row + B.bn2_b, brow + B.bn2_buf, brow + B.bn2_buf + B.cout,
                                                  B.cout, packed + B.p_bn2);
        }
    const int cf = pl.widths[3];
    if (fold_bn)
        wrn_bn_fold_kernel<<<1, 256, 0, st>>>(row + pl.bnf_w, row + pl.bnf_b, brow + pl.bnf_buf, brow + pl.bnf_buf + cf, cf,
                                              packed + pl.p_bnf);
    URSA_LAUNCH_CHECK("wrn pack kernels");
    return URSA_OK;
}

}  // namespace ursa

using namespace ursa;

extern "C" size_t ursa_bma_wrn_workspace(int S, int64_t N, int depth, int widen, int C, int algo) {
    WrnPlan pl;
    if (S < 1 || N < 1 || (algo != URSA_ALGO_TCGEN05 && algo != URSA_ALGO_TCGEN05_F16) || !wrn_build_plan(depth, widen, C, pl)) return 0;
    return wrn_chunking(N, pl).total;
}

extern "C" int ursa_bma_wrn_forward(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf, int S,
                                    const float *x, int64_t N, int depth, int widen, int C, float *proba_sum,
                                    float *entropy_sum, float *logits_out, double gamma, void *workspace,
                                    size_t workspace_bytes, int algo, void *stream) {
    URSA_REQUIRE(bank && bufbank && x && proba_sum && entropy_sum && workspace, "ursa_bma_wrn_forward: null pointer");
    URSA_REQUIRE(S >= 1 && N >= 1, "ursa_bma_wrn_forward: bad shape");
    if (algo != URSA_ALGO_TCGEN05 && algo != URSA_ALGO_TCGEN05_F16) {
        set_error("ursa_bma_wrn_forward: unknown algo %d (the WideResNet forward exists on the tcgen05 engines only)", algo);
        return URSA_ERR_UNSUPPORTED;
    }
    const bool f16 = algo == URSA_ALGO_TCGEN05_F16;
    static thread_local WrnPlan pl;
    if (!wrn_build_plan(depth, widen, C, pl)) {
        set_error("ursa_bma_wrn_forward: unsupported WRN-%d-%d (depth = 6n+4 with n <= %d, even widen factor 2..16)", depth, widen,
                  WRN_MAX_BLOCKS);
        return URSA_ERR_UNSUPPORTED;
    }
    URSA_REQUIRE(ld_bank >= pl.D, "ursa_bma_wrn_forward: ld_bank (%lld) < D (%lld)", (long long)ld_bank, (long long)pl.D);
    URSA_REQUIRE(ld_buf >= pl.NB, "ursa_bma_wrn_forward: ld_buf (%lld) < %lld", (long long)ld_buf, (long long)pl.NB);
    const WrnChunking ck = wrn_chunking(N, pl);
    URSA_REQUIRE(workspace_bytes >= ck.total, "ursa_bma_wrn_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char *wsb = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const size_t U = ck.unit_bytes, Hf = ((U / 2) + 1023) & ~(size_t)1023;
    float *A1h = reinterpret_cast<float *>(wsb), *A1l = reinterpret_cast<float *>(wsb + U);
    float *A2h = reinterpret_cast<float *>(wsb + 2 * U), *A2l = reinterpret_cast<float *>(wsb + 3 * U);
    char *hb = wsb + 4 * U;
    float *Xh[2] = {reinterpret_cast<float *>(hb), reinterpret_cast<float *>(hb + 2 * Hf)};
    float *Xl[2] = {reinterpret_cast<float *>(hb + Hf), reinterpret_cast<float *>(hb + 3 * Hf)};
    float *Ra = reinterpret_cast<float *>(hb + 4 * Hf), *Rb = reinterpret_cast<float *>(hb + 5 * Hf);
    float *packed = reinterpret_cast<float *>(wsb + 7 * U);
    float *logits = reinterpret_cast<float *>(wsb + 7 * U + ck.packed_bytes);
    const int n = pl.n, cf = pl.widths[3];

    for (int s = 0; s < S; ++s) {
        const float *row = bank + (int64_t)s * ld_bank, *brow = bufbank + (int64_t)s * ld_buf;
        if (int rc = wrn_pack_sample(pl, row, brow, packed, st, true, f16)) return rc;
        for (int64_t i0 = 0; i0 < N; i0 += ck.nc) {
            const int nc = (int)((N - i0 < ck.nc) ? (N - i0) : ck.nc);
            int xi = 0;                                   // X plane pair holding the current block's raw input
            if (f16)
                wrn_stem_kernel<true><<<nc, 256, 0, st>>>(x + i0 * 3 * 32 * 32, row + pl.conv1_w, row + pl.conv1_b,
                                                          packed + pl.blocks[0][0].p_bn1, A1h, A1l, Xh[xi], Xl[xi], nullptr, nullptr, 1);
            else
                wrn_stem_kernel<false><<<nc, 256, 0, st>>>(x + i0 * 3 * 32 * 32, row + pl.conv1_w, row + pl.conv1_b,
                                                           packed + pl.blocks[0][0].p_bn1, A1h, A1l, Xh[xi], Xl[xi], nullptr, nullptr, 1);
            URSA_LAUNCH_CHECK("wrn_stem_kernel");
            float *cur = Ra, *nxt = Rb;
            int hw = 32;
            for (int g = 0; g < 3; ++g)
                for (int b = 0; b < n; ++b) {
                    const WrnBlock &B = pl.blocks[g][b];
                    const bool last = g == 2 && b == n - 1;
                    const WrnBlock *NB = last ? nullptr : (b + 1 < n ? &pl.blocks[g][b + 1] : &pl.blocks[g + 1][0]);
                    WrnConvArgs c1 = {};
                    c1.bias = packed + B.p_bias1; c1.bn = packed + B.p_bn2; c1.out_hi = A2h; c1.out_lo = A2l;
                    if (int rc = wrn_launch_conv(A1h, A1l, hw, B.cin_p, nullptr, nullptr, 0, B.cout, 1, nc, packed + B.p_w1_hi,
                                                 packed + B.p_w1_lo, B.k1, c1, st, f16))
                        return rc;
                    WrnConvArgs c2 = {};
                    c2.bias = packed + B.p_bias2;
                    c2.res = B.transition ? nullptr : cur;
                    if (NB) {
                        c2.bn = packed + NB->p_bn1; c2.out_hi = A1h; c2.out_lo = A1l;
                        if (NB->transition) { c2.outx_hi = Xh[xi ^ 1]; c2.outx_lo = Xl[xi ^ 1]; }
                        else c2.out_raw = nxt;
                    } else {
                        c2.out_raw = nxt;
                    }
                    if (int rc = wrn_launch_conv(A2h, A2l, hw, B.cout, Xh[xi], Xl[xi], B.transition ? B.cin_p : 0, B.cout, B.stride,
                                                 nc, packed + B.p_w2_hi, packed + B.p_w2_lo, B.k2, c2, st, f16))
                        return rc;
                    if (c2.outx_hi) xi ^= 1;
                    if (c2.out_raw) { float *t = cur; cur = nxt; nxt = t; }
                    hw /= B.stride;
                }
            {   // head + softmax-average + entropy in one kernel (samples arrive one per launch: the order is the stream's)
                const float omg = (float)(1.0 - gamma), goc = (float)(gamma * 1.0 / (double)C);
                const size_t hsm = (size_t)(cf + C) * sizeof(float);
                float *lg = logits_out ? logits : nullptr;
                float *ps = proba_sum + i0 * C, *es = entropy_sum + i0;
                if (C <= 32)
                    wrn_head_kernel<1><<<nc, 256, hsm, st>>>(cur, packed + pl.p_bnf, cf, row + pl.lin_w, row + pl.lin_b, C, lg, ps, es, omg, goc);
                else if (C <= 128)
                    wrn_head_kernel<4><<<nc, 256, hsm, st>>>(cur, packed + pl.p_bnf, cf, row + pl.lin_w, row + pl.lin_b, C, lg, ps, es, omg, goc);
                else
                    wrn_head_kernel<32><<<nc, 256, hsm, st>>>(cur, packed + pl.p_bnf, cf, row + pl.lin_w, row + pl.lin_b, C, lg, ps, es, omg, goc);
                URSA_LAUNCH_CHECK("wrn_head_kernel");
            }
            if (logits_out)
                URSA_CUDA(cudaMemcpyAsync(logits_out + ((int64_t)s * N + i0) * C, logits, (size_t)nc * C * sizeof(float),
                                          cudaMemcpyDeviceToDevice, st));
        }
    }
    return URSA_OK;
}

// ---- BatchNorm re-estimation: one train-mode pass of ONE sample over the training images -----------------
namespace ursa {
struct WrnTrainLayout {
    int nc, nbmax;
    size_t unit_bytes, packed_bytes, stats_bytes, ab_bytes, total;
};
static bool wrn_train_layout(int64_t N, int batch, const WrnPlan &pl, WrnTrainLayout &L) {
    if (batch < 2 || (batch & 1) || batch > WRN_CHUNK_IMAGES || N < 1) return false;    // 8 x 8 tiles pair two images of a batch
    int64_t nc = (int64_t)batch * (WRN_CHUNK_IMAGES / batch);
    if (N < nc) nc = N;
    L.nc = (int)nc;
    L.nbmax = (int)((nc + batch - 1) / batch);
    L.unit_bytes = (((size_t)nc * 1024 * pl.widths[2] * sizeof(float)) + 1023) & ~(size_t)1023;
    L.packed_bytes = (((size_t)pl.packed_floats * sizeof(float)) + 1023) & ~(size_t)1023;
    L.stats_bytes = (((size_t)L.nbmax * 2 * pl.widths[3] * sizeof(double)) + 1023) & ~(size_t)1023;
    L.ab_bytes = (((size_t)L.nbmax * 2 * pl.widths[3] * sizeof(float)) + 1023) & ~(size_t)1023;
    L.total = 8 * L.unit_bytes + L.packed_bytes + L.stats_bytes + L.ab_bytes + 2048;
    return true;
}
}  // namespace ursa

extern "C" size_t ursa_wrn_bn_update_workspace(int64_t N, int batch, int depth, int widen, int C) {
    WrnPlan pl;
    WrnTrainLayout L;
    if (!wrn_build_plan(depth, widen, C, pl) || !wrn_train_layout(N, batch, pl, L)) return 0;
    return L.total;
}

static int wrn_bn_update_impl(const float *bank_row, float *buf_row, const float *x, int64_t N, int batch, int depth,
                              int widen, int C, void *workspace, size_t workspace_bytes, void *stream, bool f16) {
    URSA_REQUIRE(bank_row && buf_row && x && workspace, "ursa_wrn_bn_update: null pointer");
    static thread_local WrnPlan pl;
    WrnTrainLayout L;
    if (!wrn_build_plan(depth, widen, C, pl) || !wrn_train_layout(N, batch, pl, L)) {
        set_error("ursa_wrn_bn_update: unsupported WRN-%d-%d / batch %d (depth = 6n+4, even widen 2..16, even batch 2..%d)", depth,
                  widen, batch, WRN_CHUNK_IMAGES);
        return URSA_ERR_UNSUPPORTED;
    }
    URSA_REQUIRE(workspace_bytes >= L.total, "ursa_wrn_bn_update: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char *wsb = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const size_t U = L.unit_bytes, Hf = ((U / 2) + 1023) & ~(size_t)1023;
    float *A1h = reinterpret_cast<float *>(wsb), *A1l = reinterpret_cast<float *>(wsb + U);
    float *A2h = reinterpret_cast<float *>(wsb + 2 * U), *A2l = reinterpret_cast<float *>(wsb + 3 * U);
    float *T = reinterpret_cast<float *>(wsb + 4 * U);
    char *hb = wsb + 5 * U;
    float *Xh[2] = {reinterpret_cast<float *>(hb), reinterpret_cast<float *>(hb + 2 * Hf)};
    float *Xl[2] = {reinterpret_cast<float *>(hb + Hf), reinterpret_cast<float *>(hb + 3 * Hf)};
    float *Ra = reinterpret_cast<float *>(hb + 4 * Hf), *Rb = reinterpret_cast<float *>(hb + 5 * Hf);
    char *tail = wsb + 8 * U;
    float *packed = reinterpret_cast<float *>(tail);
    double *stats = reinterpret_cast<double *>(tail + L.packed_bytes);
    float *ab = reinterpret_cast<float *>(tail + L.packed_bytes + L.stats_bytes);
    const int n = pl.n;

    if (int rc = wrn_pack_sample(pl, bank_row, buf_row, packed, st, false, f16)) return rc;
    URSA_CUDA(cudaMemsetAsync(stats, 0, L.stats_bytes, st));
    auto finalize = [&](int64_t gw, int64_t gb, int64_t bufo, int Cc, int hw, int nc, int64_t n_before) {
        wrn_bn_finalize_kernel<<<(Cc + 127) / 128, 128, 0, st>>>(stats, bank_row + gw, bank_row + gb, Cc, hw, nc, batch, n_before,
                                                               buf_row + bufo, buf_row + bufo + Cc, ab);
    };
    auto apply = [&](const float *raw, int Cc, int c_pad, int hw, int nc, float *ah, float *al, float *xh, float *xl) {
        const int64_t total4 = (int64_t)nc * hw * (c_pad / 4);
        int64_t blocks = (total4 + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (f16) wrn_bn_apply_kernel<true><<<(int)blocks, 256, 0, st>>>(raw, ab, Cc, c_pad, hw, batch, total4, ah, al, xh, xl);
        else wrn_bn_apply_kernel<false><<<(int)blocks, 256, 0, st>>>(raw, ab, Cc, c_pad, hw, batch, total4, ah, al, xh, xl);
    };
    for (int64_t i0 = 0; i0 < N; i0 += L.nc) {
        const int nc = (int)((N - i0 < L.nc) ? (N - i0) : L.nc);
        int xi = 0;
        float *cur = Ra, *nxt = Rb;
        const WrnBlock &B0 = pl.blocks[0][0];
        if (f16)
            wrn_stem_kernel<true><<<nc, 256, 0, st>>>(x + i0 * 3 * 32 * 32, bank_row + pl.conv1_w, bank_row + pl.conv1_b, nullptr, nullptr,
                                                      nullptr, Xh[xi], Xl[xi], cur, stats, batch);
        else
            wrn_stem_kernel<false><<<nc, 256, 0, st>>>(x + i0 * 3 * 32 * 32, bank_row + pl.conv1_w, bank_row + pl.conv1_b, nullptr, nullptr,
                                                       nullptr, Xh[xi], Xl[xi], cur, stats, batch);
        URSA_LAUNCH_CHECK("wrn_stem_kernel");
        finalize(B0.bn1_w, B0.bn1_b, B0.bn1_buf, 16, 1024, nc, i0);
        apply(cur, 16, 32, 1024, nc, A1h, A1l, nullptr, nullptr);
        int hw = 32;
        for (int g = 0; g < 3; ++g)
            for (int b = 0; b < n; ++b) {
                const WrnBlock &B = pl.blocks[g][b];
                const bool last = g == 2 && b == n - 1;
                const WrnBlock *NB = last ? nullptr : (b + 1 < n ? &pl.blocks[g][b + 1] : &pl.blocks[g + 1][0]);
                WrnConvArgs c1 = {};
                c1.bias = packed + B.p_bias1; c1.out_raw = T; c1.stats = stats; c1.batch = batch;
                if (int rc = wrn_launch_conv(A1h, A1l, hw, B.cin_p, nullptr, nullptr, 0, B.cout, 1, nc, packed + B.p_w1_hi,
                                             packed + B.p_w1_lo, B.k1, c1, st, f16))
                    return rc;
                finalize(B.bn2_w, B.bn2_b, B.bn2_buf, B.cout, hw * hw, nc, i0);
                apply(T, B.cout, B.cout, hw * hw, nc, A2h, A2l, nullptr, nullptr);
                WrnConvArgs c2 = {};
                c2.bias = packed + B.p_bias2; c2.res = B.transition ? nullptr : cur; c2.out_raw = nxt; c2.stats = stats; c2.batch = batch;
                if (int rc = wrn_launch_conv(A2h, A2l, hw, B.cout, Xh[xi], Xl[xi], B.transition ? B.cin_p : 0, B.cout, B.stride, nc,
                                             packed + B.p_w2_hi, packed + B.p_w2_lo, B.k2, c2, st, f16))
                    return rc;
                hw /= B.stride;
                if (NB) {
                    finalize(NB->bn1_w, NB->bn1_b, NB->bn1_buf, B.cout, hw * hw, nc, i0);
                    const bool nx = NB->transition;
                    apply(nxt, B.cout, B.cout, hw * hw, nc, A1h, A1l, nx ? Xh[xi ^ 1] : nullptr, nx ? Xl[xi ^ 1] : nullptr);
                    if (nx) xi ^= 1;
                } else {
                    finalize(pl.bnf_w, pl.bnf_b, pl.bnf_buf, B.cout, hw * hw, nc, i0);
                }
                float *t = cur; cur = nxt; nxt = t;
            }
        URSA_LAUNCH_CHECK("wrn train-mode kernels");
    }
    return URSA_OK;
}

extern "C" int ursa_wrn_bn_update(const float *bank_row, float *buf_row, const float *x, int64_t N, int batch, int depth,
                                  int widen, int C, void *workspace, size_t workspace_bytes, void *stream) {
    return wrn_bn_update_impl(bank_row, buf_row, x, N, batch, depth, widen, C, workspace, workspace_bytes, stream, false);
}

extern "C" int ursa_wrn_bn_update_algo(const float *bank_row, float *buf_row, const float *x, int64_t N, int batch, int depth,
                                       int widen, int C, void *workspace, size_t workspace_bytes, int algo, void *stream) {
    URSA_REQUIRE(algo == URSA_ALGO_TCGEN05 || algo == URSA_ALGO_TCGEN05_F16, "ursa_wrn_bn_update_algo: algo must be URSA_ALGO_TCGEN05 or URSA_ALGO_TCGEN05_F16");
    return wrn_bn_update_impl(bank_row, buf_row, x, N, batch, depth, widen, C, workspace, workspace_bytes, stream,
                              algo == URSA_ALGO_TCGEN05_F16);
}
