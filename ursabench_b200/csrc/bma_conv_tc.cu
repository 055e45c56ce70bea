// K3 (PreResNet) on 5th-generation tensor cores: sample-batched implicit-GEMM 3x3 convolutions, 3xTF32,
// TMA-fed, accumulators in TMEM (see bma_mlp_tc.cu for the 3xTF32 rationale).
//
// Data layout (per chunk of S_c samples x N_c images, pair p = s*N_c + n), NHWC:
//   R   raw block outputs            fp32 [P][H][W][C]      (residual stream)
//   A   pre-activated conv inputs    tf32 hi / lo planes [P][H][W][C] = split(relu(bn(R)))  -- written by the
//       PRODUCER's epilogue, so the consumer's TMA can fetch the im2col slices directly and the zero padding
//       (TMA out-of-bounds fill) is applied after the activation, exactly like PyTorch pads relu(bn(x)).
// Implicit GEMM per CTA: M = 128 output pixels (4 rows x 32 | 8 rows x 16 | 2 images x 8 x 8), N = Cout,
// K = 9 taps x Cin.  For each tap the [128 pixels x CW channels] A slice is ONE 5-D tiled TMA box at shifted
// coordinates (c0, w0+kw-1, h0+kh-1, n0, s); stride-2 convs use four parity-split tensor maps (a space-to-depth
// view) so that every tap is still a dense box.  B = filters packed K-major [Cout][tap*Cin + ci], split hi / lo.
//   warp 0: TMA producer | warp 1: TMEM alloc + MMA issuer (3 tcgen05.mma per UMMA_K) | warps 2-5: epilogue
//   epilogue mode 0 (conv1 of a block): A' = split(relu(bn2(acc)))
//   epilogue mode 1 (conv2 of a block): R' = acc + shortcut ; A'' = split(relu(bn_next(R')))
// The 3->16 stem, the 1x1 stride-2 shortcuts and the head (1.6 % of the FLOPs) stay on CUDA cores.
#include <stdlib.h>

#include "preresnet_plan.cuh"
#include "tc_common.cuh"
#include "bma_conv_fused.cuh"
#include "bma_conv_fused16.cuh"
#include "bma_epilogue.cuh"

namespace ursa {

struct ConvTcMaps {
    CUtensorMap a_hi[4], a_lo[4];      // index = h-parity * 2 + w-parity for stride 2; [0] only for stride 1
    CUtensorMap b_hi, b_lo;
};

struct ConvTcArgs {
    int cin, cout, hout, stride, n_images;
    int kchunks;                       // cin / CW
    int stages;
    uint32_t tmem_cols;
    int mode;                          // 0 / 1, see above
    const float *packed;
    int64_t ld_packed, bn_off;         // epilogue BN (a[cout], b[cout]) in the packed row; < 0: none
    const float *res;                  // mode 1: shortcut [P][hout][hout][cout]
    float *out_raw;                    // mode 1
    float *out_hi, *out_lo;            // split outputs (may be null in mode 1 for the last block)
    // mode 0, FP16-split consumer: relu(bn(acc)) / 16 goes out as the next stage kernel's PLANE IMAGE (bma_conv_fused16.cuh)
    unsigned char *pi_out = nullptr;   // nullable
    int pi_G = 1, pi_pitch = 0, pi_img_pos = 0, pi_nplanes = 0, pi_ngroups = 0;
    int64_t pi_pass_bytes = 0;
    // mode 2 (train-mode BatchNorm pass, ursa_preresnet_bn_update): out_raw = acc (+ res) and the per-(sample, batch,
    // channel) sums of v and v^2 over the batch's pixels go to stats [S_c][n_batches][2][cout] (fp64 atomics)
    double *stats = nullptr;
    int batch = 0, n_batches = 0;
};

constexpr int CTC_THREADS = 192, CTC_MAX_STAGES = 8;

template <int CW>      // channels per shared-memory row: 16 (64-byte swizzle) or 32 (128-byte swizzle)
__global__ void __launch_bounds__(CTC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs a) {
    constexpr int ROW_BYTES = CW * 4;
    constexpr uint32_t A_BYTES = 128 * ROW_BYTES;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[CTC_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[CTC_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bn_s[128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.y;
    // tile geometry
    const int WT = a.hout, HT = a.hout >= 16 ? 128 / a.hout : a.hout, NT = 128 / (WT * HT);
    const int tpi = (a.hout * a.hout) / 128;                 // tiles per image (0 when one tile spans 2 images)
    int n0, h0;
    if (NT == 1) { n0 = blockIdx.x / tpi; h0 = (blockIdx.x % tpi) * HT; } else { n0 = blockIdx.x * NT; h0 = 0; }

    const uint32_t b_bytes = (uint32_t)a.cout * ROW_BYTES;
    const uint32_t stage_bytes = 2 * A_BYTES + 2 * b_bytes;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int k_blocks = 9 * a.kchunks;

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, a.tmem_cols);
    if (warp >= 2 && a.bn_off >= 0) {
        const float *bn = a.packed + (int64_t)s * a.ld_packed + a.bn_off;
        for (int i = threadIdx.x - 64; i < 2 * a.cout; i += 128) bn_s[i] = __ldg(bn + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            for (int kb = 0; kb < k_blocks; ++kb) {
                const int st = kb % a.stages;
                const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
                const int tap = kb / a.kchunks, cc = kb - tap * a.kchunks;
                const int kh = tap / 3, kw = tap - kh * 3;
                mbar_wait_a(smem_u32(&empty_bar[st]), ph ^ 1u);
                const uint32_t fb = smem_u32(&full_bar[st]);
                mbar_expect_tx_a(fb, stage_bytes);
                const uint32_t base = smem_base + (uint32_t)st * stage_bytes;
                int mi = 0, cw = kw - 1, ch = h0 + kh - 1;
                if (a.stride == 2) {
                    mi = ((kh + 1) & 1) * 2 + ((kw + 1) & 1);        // parity of (kh-1, kw-1)
                    cw = (kw - 1) >> 1;                                // floor((kw-1)/2)
                    ch = h0 + ((kh - 1) >> 1);
                }
                tma_load_5d_a(base, &maps.a_hi[mi], cc * CW, cw, ch, n0, s, fb);
                tma_load_5d_a(base + A_BYTES, &maps.a_lo[mi], cc * CW, cw, ch, n0, s, fb);
                tma_load_3d_a(base + 2 * A_BYTES, &maps.b_hi, tap * a.cin + cc * CW, 0, s, fb);
                tma_load_3d_a(base + 2 * A_BYTES + b_bytes, &maps.b_lo, tap * a.cin + cc * CW, 0, s, fb);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_tf32_idesc(128, a.cout);
            uint32_t acc = 0;
            for (int kb = 0; kb < k_blocks; ++kb) {
                const int st = kb % a.stages;
                const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
                mbar_wait_a(smem_u32(&full_bar[st]), ph);
                tc_fence_after();
                const uint32_t base = smem_base + (uint32_t)st * stage_bytes;
                const uint64_t d_ahi = make_kmajor_desc<ROW_BYTES>(base), d_alo = make_kmajor_desc<ROW_BYTES>(base + A_BYTES);
                const uint64_t d_bhi = make_kmajor_desc<ROW_BYTES>(base + 2 * A_BYTES);
                const uint64_t d_blo = make_kmajor_desc<ROW_BYTES>(base + 2 * A_BYTES + b_bytes);
#pragma unroll
                for (int k = 0; k < CW / 8; ++k) {
                    const uint64_t koff = (uint64_t)((k * 32) >> 4);
                    umma_tf32(tmem_base, d_alo + koff, d_bhi + koff, idesc, acc);
                    acc = 1;
                    umma_tf32(tmem_base, d_ahi + koff, d_blo + koff, idesc, 1);
                    umma_tf32(tmem_base, d_ahi + koff, d_bhi + koff, idesc, 1);
                }
                umma_commit(smem_u32(&empty_bar[st]));
            }
            umma_commit(smem_u32(&tmem_full_bar));
        }
    } else {
        // ===== epilogue: thread = output pixel =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int w = r % WT, h = (r / WT) % HT, nl = r / (WT * HT);
        const int n = n0 + nl;
        const bool valid = n < a.n_images;
        const int64_t off = ((((int64_t)s * a.n_images + n) * a.hout + (h0 + h)) * a.hout + w) * a.cout;
        const bool has_bn = a.bn_off >= 0;
        mbar_wait_a(smem_u32(&tmem_full_bar), 0);
        tc_fence_after();
        // plane-image address of this pixel: pass = (sample, image group), position = (g (H + 1) + h) pitch + w
        unsigned char *pi_px = nullptr;
        if (a.pi_out != nullptr && valid) {
            const int grp = n / a.pi_G, g = n - grp * a.pi_G;
            pi_px = a.pi_out + ((int64_t)s * a.pi_ngroups + grp) * a.pi_pass_bytes +
                    (int64_t)((g * (a.hout + 1) + (h0 + h)) * a.pi_pitch + w) * 16;
        }
        for (int c0 = 0; c0 < a.cout; c0 += 16) {
            uint32_t rr[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, rr);
            if (a.mode == 2) {
                // raw output (+ residual) and batch statistics.  The 32 pixels of a warp belong to one image (or to two images
                // of one batch): a transposing butterfly leaves lane l with the warp's sum of value l of
                // [v_0 .. v_15, v_0^2 .. v_15^2] after 31 shuffles, then one fp64 atomic per lane.
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = valid ? __uint_as_float(rr[i]) : 0.f;
                if (valid) {
                    if (a.res != nullptr) {
                        const float4 *rp = reinterpret_cast<const float4 *>(a.res + off + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 t = __ldg(rp + i);
                            v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                        }
                    }
                    float4 *op = reinterpret_cast<float4 *>(a.out_raw + off + c0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                float w32[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) { w32[i] = v[i]; w32[16 + i] = v[i] * v[i]; }
#pragma unroll
                for (int half = 16; half >= 1; half >>= 1) {
                    const bool up = (lane & half) != 0;
#pragma unroll
                    for (int i = 0; i < half; ++i) {
                        const float send = up ? w32[i] : w32[i + half];
                        const float got = __shfl_xor_sync(0xffffffffu, send, half);
                        w32[i] = (up ? w32[i + half] : w32[i]) + got;
                    }
                }
                // lane l now holds value index bit-reversed-free mapping: after the steps lane l owns index l (bits taken MSB first)
                const int n_first = n0 + (NT == 1 ? 0 : (q * 32) / (WT * HT));      // image of this warp's first pixel
                if (n_first < a.n_images) {
                    const int j = n_first / a.batch;
                    const int idx = lane;                                          // [0,16): sum v_c ; [16,32): sum v_c^2
                    double *sp = a.stats + (((int64_t)s * a.n_batches + j) * 2 + (idx >> 4)) * a.cout + c0 + (idx & 15);
                    atomicAdd(sp, (double)w32[0]);
                }
                continue;
            }
            if (!valid) continue;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
            if (pi_px != nullptr) {
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int c = c0 + i + 2 * jj;
                        const float y0 = relu_nan(fmaf(bn_s[c], v[i + 2 * jj], bn_s[a.cout + c])) * kActDown;
                        const float y1 = relu_nan(fmaf(bn_s[c + 1], v[i + 2 * jj + 1], bn_s[a.cout + c + 1])) * kActDown;
                        split_h2(y0, y1, hw[jj], lw[jj]);
                    }
                    const int64_t pl = (int64_t)((c0 + i) >> 3) * a.pi_img_pos * 16;
                    *reinterpret_cast<uint4 *>(pi_px + pl) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4 *>(pi_px + (int64_t)a.pi_nplanes * a.pi_img_pos * 16 + pl) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
                continue;
            }
            if (a.mode == 1) {
                const float4 *rp = reinterpret_cast<const float4 *>(a.res + off + c0);
                float4 *op = reinterpret_cast<float4 *>(a.out_raw + off + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 t = __ldg(rp + i);
                    v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                    op[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            if (a.out_hi) {
                float4 *hp = reinterpret_cast<float4 *>(a.out_hi + off + c0);
                float4 *lp = reinterpret_cast<float4 *>(a.out_lo + off + c0);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float y[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = v[i + j];
                        if (has_bn) t = relu_nan(fmaf(bn_s[c0 + i + j], t, bn_s[a.cout + c0 + i + j]));
                        y[j] = t;
                    }
                    if (a.out_lo == nullptr) {          // plain fp32 activations (consumer splits them itself)
                        hp[i >> 2] = make_float4(y[0], y[1], y[2], y[3]);
                        continue;
                    }
                    float4 hv, lv;
                    hv.x = rn_tf32(y[0]); hv.y = rn_tf32(y[1]); hv.z = rn_tf32(y[2]); hv.w = rn_tf32(y[3]);
                    lv.x = rn_tf32(y[0] - hv.x); lv.y = rn_tf32(y[1] - hv.y); lv.z = rn_tf32(y[2] - hv.z); lv.w = rn_tf32(y[3] - hv.w);
                    hp[i >> 2] = hv;
                    lp[i >> 2] = lv;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// ---- CUDA-core helpers in NHWC -----------------------------------------------------------------------------
// stem: R0 = conv3x3(x) (3 -> 16), A0 = split(relu(bn1_of_block0(R0)));  x is NCHW and shared by all samples
__global__ void __launch_bounds__(256) stem_nhwc_kernel(const float *__restrict__ x, const float *__restrict__ packed,
                                                         int64_t ld_packed, int64_t w_off, int64_t bn_off, int n_images,
                                                         float *__restrict__ out_raw, float *__restrict__ out_hi,
                                                         float *__restrict__ out_lo, unsigned char *__restrict__ pi_out) {
    // one CTA per (image, sample); thread = 4 consecutive pixels of a row x all 16 output channels (64 accumulators), so a
    // filter tap's 16 weights (4 broadcast LDS.128) feed 64 FMAs and an input row segment (6 LDS) feeds 3 taps
    __shared__ float xs[3][34][35];
    __shared__ __align__(16) float ws[27 * 16];
    __shared__ float bn[32];
    const int n = blockIdx.x, s = blockIdx.y;
    const float *pk = packed + (int64_t)s * ld_packed;
    for (int i = threadIdx.x; i < 27 * 16; i += 256) ws[i] = __ldg(pk + w_off + i);     // [ci][tap][co]
    for (int i = threadIdx.x; i < 32; i += 256) bn[i] = __ldg(pk + bn_off + i);
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
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int64_t off = (((int64_t)s * n_images + n) * 1024 + h * 32 + w0 + p) * 16;
        if (pi_out != nullptr) {
            // stage-1 plane image (F16Cfg<16>: one image per pass, pitch 33, 2 planes of 8 channels, hi then lo')
            using P16 = F16Cfg<16>;
            unsigned char *px = pi_out + ((int64_t)s * n_images + n) * P16::PASS_BYTES + (int64_t)(h * P16::PITCH + w0 + p) * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = i + 2 * jj;
                    split_h2(relu_nan(fmaf(bn[c], acc[p][c], bn[16 + c])) * kActDown,
                             relu_nan(fmaf(bn[c + 1], acc[p][c + 1], bn[16 + c + 1])) * kActDown, hw[jj], lw[jj]);
                }
                *reinterpret_cast<uint4 *>(px + (int64_t)(i >> 3) * P16::IMG_POS * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(px + (int64_t)(2 + (i >> 3)) * P16::IMG_POS * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            *reinterpret_cast<float4 *>(out_raw + off + i) = make_float4(acc[p][i], acc[p][i + 1], acc[p][i + 2], acc[p][i + 3]);
            if (out_hi == nullptr) continue;
            float y[4], hv[4], lv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                y[k] = relu_nan(fmaf(bn[i + k], acc[p][i + k], bn[16 + i + k]));
                hv[k] = rn_tf32(y[k]);
                lv[k] = rn_tf32(y[k] - hv[k]);
            }
            *reinterpret_cast<float4 *>(out_hi + off + i) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4 *>(out_lo + off + i) = make_float4(lv[0], lv[1], lv[2], lv[3]);
        }
    }
}

// test images (NCHW fp32, shared by all samples) -> stage-1 plane images of the FP16-split path: position (h, w) of plane 0
// holds [x0, x1, x2, 0 x 5] / 16 as FP16 hi and lo'; plane 1 (channels 8..15) stays at the zeros of the workspace memset.
// With these as its input planes the network's conv1 is simply the first conv of the stage-1 kernel's chain.
__global__ void __launch_bounds__(256) image_planes_kernel(const float *__restrict__ x, int n_images, unsigned char *__restrict__ pi) {
    using P16 = F16Cfg<16>;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= n_images * 1024) return;
    const int n = idx >> 10, px = idx & 1023, h = px >> 5, w = px & 31;
    const float *xp = x + (int64_t)n * 3072 + px;
    uint32_t h01, l01, h2, l2;
    split_h2(__ldg(xp) * kActDown, __ldg(xp + 1024) * kActDown, h01, l01);
    split_h2(__ldg(xp + 2048) * kActDown, 0.f, h2, l2);
    unsigned char *dst = pi + (int64_t)n * P16::PASS_BYTES + (int64_t)(h * P16::PITCH + w) * 16;
    *reinterpret_cast<uint4 *>(dst) = make_uint4(h01, h2, 0u, 0u);
    *reinterpret_cast<uint4 *>(dst + (int64_t)2 * P16::IMG_POS * 16) = make_uint4(l01, l2, 0u, 0u);
}

// 1x1 stride-2 shortcut on the raw NHWC block input
__global__ void __launch_bounds__(256) shortcut_nhwc_kernel(const float *__restrict__ in, const float *__restrict__ packed,
                                                             int64_t ld_packed, int64_t w_off, int cin, int cout, int hout,
                                                             int n_images, float *__restrict__ out, int compact) {
    // one CTA per (image, sample); work item = (output pixel, group of 16 output channels): the pixel's cin inputs are
    // read as float4s, the weights [ci][co] come from shared memory as LDS.128
    __shared__ __align__(16) float ws[32 * 64];
    const int n = blockIdx.x, s = blockIdx.y;
    const float *pk = packed + (int64_t)s * ld_packed + w_off;                    // [ci][co]
    for (int i = threadIdx.x; i < cin * cout; i += 256) ws[i] = __ldg(pk + i);
    __syncthreads();
    // compact != 0: `in` holds only the even-row / even-column pixels, [P][hout][hout][cin] (FP16-split stage kernels)
    const int hin = compact ? hout : hout * 2, ngrp = cout >> 4, step = compact ? 1 : 2;
    const float *ib = in + ((int64_t)s * n_images + n) * hin * hin * cin;
    float *ob = out + ((int64_t)s * n_images + n) * hout * hout * cout;
    for (int item = threadIdx.x; item < hout * hout * ngrp; item += 256) {
        const int g = item % ngrp, px = item / ngrp;
        const int h = px / hout, w = px - h * hout;
        const float4 *ip = reinterpret_cast<const float4 *>(ib + ((int64_t)(step * h) * hin + step * w) * cin);
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        for (int c4 = 0; c4 < cin / 4; ++c4) {
            const float4 xv = __ldg(ip + c4);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 *wp = reinterpret_cast<const float4 *>(ws + (c4 * 4 + j) * cout + g * 16);
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const float4 w4 = wp[qd];
                    acc[4 * qd + 0] = fmaf(xa[j], w4.x, acc[4 * qd + 0]);
                    acc[4 * qd + 1] = fmaf(xa[j], w4.y, acc[4 * qd + 1]);
                    acc[4 * qd + 2] = fmaf(xa[j], w4.z, acc[4 * qd + 2]);
                    acc[4 * qd + 3] = fmaf(xa[j], w4.w, acc[4 * qd + 3]);
                }
            }
        }
        float4 *op = reinterpret_cast<float4 *>(ob + (int64_t)px * cout + g * 16);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) op[qd] = make_float4(acc[4 * qd], acc[4 * qd + 1], acc[4 * qd + 2], acc[4 * qd + 3]);
    }
}

// final BN + ReLU + AvgPool2d(8) + fc on NHWC raw [P][8][8][64]; one warp per pair
__global__ void __launch_bounds__(256) head_nhwc_kernel(const float *__restrict__ act, const float *__restrict__ packed,
                                                         int64_t ld_packed, int64_t bn_off, int64_t fc_off, int n_images,
                                                         int n_pairs, int C, float *__restrict__ logits) {
    __shared__ float feat_s[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 8 + warp;
    if (pair >= n_pairs) return;
    const int s = pair / n_images;
    const float *pk = packed + (int64_t)s * ld_packed;
    const float *x = act + (int64_t)pair * 64 * 64;
    const float a0 = __ldg(pk + bn_off + lane), b0 = __ldg(pk + bn_off + 64 + lane);
    const float a1 = __ldg(pk + bn_off + 32 + lane), b1 = __ldg(pk + bn_off + 96 + lane);
    float f0 = 0.f, f1 = 0.f;
    for (int px = 0; px < 64; ++px) {
        f0 += relu_nan(fmaf(a0, __ldg(x + px * 64 + lane), b0));
        f1 += relu_nan(fmaf(a1, __ldg(x + px * 64 + 32 + lane), b1));
    }
    feat_s[warp][lane] = f0 * (1.f / 64.f);
    feat_s[warp][32 + lane] = f1 * (1.f / 64.f);
    __syncwarp();
    const float *fw = pk + fc_off, *fb = pk + fc_off + (int64_t)C * 64;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) acc = fmaf(feat_s[warp][k], __ldg(fw + c * 64 + k), acc);
        logits[(int64_t)pair * C + c] = acc + __ldg(fb + c);
    }
}

// The same head FUSED with the BMA epilogue: one warp per IMAGE walks the chunk's samples in order -- final BN + ReLU +
// AvgPool2d(8) + fc leave the logits in registers (class c in lane c % 32), softmax_accumulate_row adds softmax_s and the
// smoothed entropy to the warp's running sums, and the image's row of proba_sum / entropy_sum is read and written once.
// No logits round trip, no separate accumulate launch; the per-row sample order (= the reference's list order) is kept.
template <int PER_LANE>
__global__ void __launch_bounds__(128) head_bma_kernel(const float *__restrict__ act, const float *__restrict__ packed,
                                                        int64_t ld_packed, int64_t bn_off, int64_t fc_off, int n_images,
                                                        int n_samples, int C, float *__restrict__ proba_sum,
                                                        float *__restrict__ entropy_sum, float one_minus_gamma,
                                                        float gamma_over_c) {
    // one CTA per image; warp w computes the logits of samples s = w, w + 4, .. into shared memory, then warp 0 folds
    // them into the image's row in sample order
    extern __shared__ float lg_s[];                       // [n_samples][C]
    __shared__ float feat_s[4][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x;
    for (int s = warp; s < n_samples; s += 4) {
        const float *pk = packed + (int64_t)s * ld_packed;
        const float *x = act + ((int64_t)s * n_images + n) * 64 * 64;
        const float a0 = __ldg(pk + bn_off + lane), b0 = __ldg(pk + bn_off + 64 + lane);
        const float a1 = __ldg(pk + bn_off + 32 + lane), b1 = __ldg(pk + bn_off + 96 + lane);
        float f0 = 0.f, f1 = 0.f;
#pragma unroll 8
        for (int px = 0; px < 64; ++px) {
            f0 += relu_nan(fmaf(a0, __ldg(x + px * 64 + lane), b0));
            f1 += relu_nan(fmaf(a1, __ldg(x + px * 64 + 32 + lane), b1));
        }
        __syncwarp();
        feat_s[warp][lane] = f0 * (1.f / 64.f);
        feat_s[warp][32 + lane] = f1 * (1.f / 64.f);
        __syncwarp();
        const float *fw = pk + fc_off, *fb = pk + fc_off + (int64_t)C * 64;
        for (int c = lane; c < C; c += 32) {
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < 64; ++k) acc = fmaf(feat_s[warp][k], __ldg(fw + c * 64 + k), acc);
            lg_s[s * C + c] = acc + __ldg(fb + c);
        }
    }
    __syncthreads();
    if (warp != 0) return;
    float P[PER_LANE];
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) P[j] = (lane + 32 * j < C) ? proba_sum[(int64_t)n * C + lane + 32 * j] : 0.f;
    float E = entropy_sum[n];
    for (int s = 0; s < n_samples; ++s) {
        float lg[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) lg[j] = (lane + 32 * j < C) ? lg_s[s * C + lane + 32 * j] : 0.f;
        softmax_accumulate_row<PER_LANE>(lg, C, lane, one_minus_gamma, gamma_over_c, P, E);
    }
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j)
        if (lane + 32 * j < C) proba_sum[(int64_t)n * C + lane + 32 * j] = P[j];
    if (lane == 0) entropy_sum[n] = E;
}

static int launch_head_bma(const float *act, const float *packed, int64_t ld_packed, int64_t bn_off, int64_t fc_off, int nc,
                           int sc, int C, float *proba_sum, float *entropy_sum, double gamma, cudaStream_t st) {
    const float omg = (float)(1.0 - gamma), goc = (float)(gamma * 1.0 / C);      // util.py:134: gamma * 1 / C
    const size_t sm = (size_t)sc * C * sizeof(float);
    if (C <= 32)
        head_bma_kernel<1><<<nc, 128, sm, st>>>(act, packed, ld_packed, bn_off, fc_off, nc, sc, C, proba_sum, entropy_sum, omg, goc);
    else if (C <= 128)
        head_bma_kernel<4><<<nc, 128, sm, st>>>(act, packed, ld_packed, bn_off, fc_off, nc, sc, C, proba_sum, entropy_sum, omg, goc);
    else
        head_bma_kernel<32><<<nc, 128, sm, st>>>(act, packed, ld_packed, bn_off, fc_off, nc, sc, C, proba_sum, entropy_sum, omg, goc);
    URSA_LAUNCH_CHECK("head_bma_kernel");
    return URSA_OK;
}

// ---- host ---------------------------------------------------------------------------------------------------------
constexpr int kTcChunkSamples = 8, kTcChunkImages = 512;      // images per launch: layer-by-layer engine and bn_update batches
// images per launch of the fused stage engines: a launch is 8 samples x 2 500 images = 20 000 passes over 148 persistent CTAs, so the
// ragged last round and the launch gaps of the 7-kernel chain are 1/5 of what 512-image launches pay (490 -> 465 ms at
// S = 100 x N = 10 000); 9.2 GB of workspace on a 180 GB part
constexpr int kFusedChunkImages = 2500;

struct TcChunking {
    int sc, nc;
    size_t raw_bytes, packed_bytes, logit_bytes, total;
};

static TcChunking tc_chunking(int S, int64_t N, const NetPlan &pl, bool fused = false) {
    TcChunking c;
    c.sc = S < kTcChunkSamples ? S : kTcChunkSamples;
    static const int chunk_images = [] {            // URSA_CHUNK_IMAGES: images per launch of the conv forwards (experiments)
        const char *e = getenv("URSA_CHUNK_IMAGES");
        const int v = e ? atoi(e) : 0;
        return v >= 64 && v <= 16384 ? v : 0;
    }();
    // fused engines: a launch holds up to 8 x 2 500 (sample, image) pairs whatever the split -- fewer samples, more images
    // (S = 1: the whole 10 000-image test set in one launch chain; matters for the small shares of an 8-rank evaluation)
    const int ci = chunk_images ? chunk_images : (fused ? kFusedChunkImages * (kTcChunkSamples / c.sc) : kTcChunkImages);
    c.nc = (int)(N < ci ? N : ci);
    const size_t pairs = (size_t)c.sc * c.nc;
    c.raw_bytes = pairs * 16 * 32 * 32 * sizeof(float);               // largest activation: 64 KB per pair
    c.packed_bytes = (((size_t)c.sc * pl.packed_floats * sizeof(float)) + 1023) & ~(size_t)1023;
    c.logit_bytes = (((pairs * pl.C * sizeof(float)) + 1023) & ~(size_t)1023);
    // R_a, R_b, R_shortcut, A1 hi/lo, A2 hi/lo = 7 planes
    c.total = 7 * c.raw_bytes + c.packed_bytes + c.logit_bytes + 2048;
    return c;
}

// activations plane [S_c][N_c][H][W][C]; parity >= 0 selects the (h, w) parity sub-lattice of a stride-2 conv input
static int make_act_map(CUtensorMap *tm, const float *plane, int sc, int nc, int H, int C, int cw, int hout, int stride,
                        int parity) {
    const int WT = hout, HT = hout >= 16 ? 128 / hout : hout, NT = 128 / (WT * HT);
    const uint32_t box[5] = {(uint32_t)cw, (uint32_t)WT, (uint32_t)HT, (uint32_t)NT, 1};
    if (stride == 1) {
        const uint64_t dims[5] = {(uint64_t)C, (uint64_t)H, (uint64_t)H, (uint64_t)nc, (uint64_t)sc};
        const uint64_t st[4] = {(uint64_t)C * 4, (uint64_t)H * C * 4, (uint64_t)H * H * C * 4, (uint64_t)nc * H * H * C * 4};
        return make_tensor_map(tm, plane, 5, dims, st, box, cw * 4);
    }
    const int hp = parity >> 1, wp = parity & 1;
    const float *base = plane + ((int64_t)hp * H + wp) * C;
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)(H / 2), (uint64_t)(H / 2), (uint64_t)nc, (uint64_t)sc};
    const uint64_t st[4] = {(uint64_t)2 * C * 4, (uint64_t)2 * H * C * 4, (uint64_t)H * H * C * 4, (uint64_t)nc * H * H * C * 4};
    return make_tensor_map(tm, base, 5, dims, st, box, cw * 4);
}

static int launch_conv_tc(const float *a_hi, const float *a_lo, int hin, int cin, int cout, int stride, int sc, int nc,
                          const float *packed, int64_t ld_packed, int64_t w_hi_off, int64_t w_lo_off, ConvTcArgs g,
                          cudaStream_t st) {
    const int hout = hin / stride;
    const int cw = cin >= 32 ? 32 : 16;
    ConvTcMaps maps;
    const int nmaps = stride == 2 ? 4 : 1;
    for (int i = 0; i < nmaps; ++i) {
        if (int rc = make_act_map(&maps.a_hi[i], a_hi, sc, nc, hin, cin, cw, hout, stride, i)) return rc;
        if (int rc = make_act_map(&maps.a_lo[i], a_lo, sc, nc, hin, cin, cw, hout, stride, i)) return rc;
    }
    for (int i = nmaps; i < 4; ++i) { maps.a_hi[i] = maps.a_hi[0]; maps.a_lo[i] = maps.a_lo[0]; }
    {
        const uint64_t dims[3] = {(uint64_t)9 * cin, (uint64_t)cout, (uint64_t)sc};
        const uint64_t sb[2] = {(uint64_t)9 * cin * 4, (uint64_t)ld_packed * 4};
        const uint32_t box[3] = {(uint32_t)cw, (uint32_t)cout, 1};
        if (int rc = make_tensor_map(&maps.b_hi, packed + w_hi_off, 3, dims, sb, box, cw * 4)) return rc;
        if (int rc = make_tensor_map(&maps.b_lo, packed + w_lo_off, 3, dims, sb, box, cw * 4)) return rc;
    }
    g.cin = cin; g.cout = cout; g.hout = hout; g.stride = stride; g.n_images = nc;
    g.kchunks = cin / cw;
    g.packed = packed; g.ld_packed = ld_packed;
    uint32_t cols = 32;
    while (cols < (uint32_t)cout) cols <<= 1;
    g.tmem_cols = cols;
    const size_t stage_bytes = 2 * (size_t)128 * cw * 4 + 2 * (size_t)cout * cw * 4;
    // one tile per CTA has only 9 * kchunks k-blocks, so the fixed per-tile latency (barrier init, TMEM alloc, first
    // TMA round trip, epilogue) is hidden by co-resident CTAs rather than by a deep pipeline: 3 stages of 18 KB
    // (Cin = 16) -> 4 CTAs/SM, 2-3 stages of 40-48 KB (Cin >= 32) -> 2 CTAs/SM
    int stages = cw == 16 ? 3 : (stage_bytes * 3 + 1024 <= (size_t)(113 << 10) ? 3 : 2);
    if (const char *e = getenv("URSA_CONV_TC_STAGES")) { const int v = atoi(e); if (v >= 1) stages = v; }
    if ((size_t)stages * stage_bytes + 1024 > (size_t)(226 << 10)) stages = (int)(((size_t)(226 << 10) - 1024) / stage_bytes);
    if (stages > CTC_MAX_STAGES) stages = CTC_MAX_STAGES;
    if (stages > 9 * g.kchunks) stages = 9 * g.kchunks;
    g.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    const int WT = hout, HT = hout >= 16 ? 128 / hout : hout, NT = 128 / (WT * HT);
    const int tiles = NT == 1 ? nc * ((hout * hout) / 128) : (nc + NT - 1) / NT;
    dim3 grid(tiles, sc);
    if (cw == 16) {
        URSA_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv3x3_tc_kernel<16><<<grid, CTC_THREADS, smem, st>>>(maps, g);
    } else {
        URSA_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv3x3_tc_kernel<32><<<grid, CTC_THREADS, smem, st>>>(maps, g);
    }
    URSA_LAUNCH_CHECK("conv3x3_tc_kernel");
    return URSA_OK;
}


// ---- stride-2 transition conv of the FP16-split path ------------------------------------------------------------------
// conv1 of layer2.0 / layer3.0 (3x3, stride 2, Cin -> 2 Cin) as an implicit GEMM like conv3x3_tc_kernel, but on the FP16-split
// operands the stage kernels use: the input is what the previous stage kernel's last epilogue stored by TMA -- rows
// [pass][position][hi(16 ch) | lo'(16 ch)] x (Cin / 16) halves of relu(bn1(R)) / 16 in the padded-pitch position space of that stage -- so a
// tap's A tile is ONE 5-D TMA box (64- / 128-byte rows: hi and lo' of a pixel travel together; parity-split maps make the
// stride-2 taps dense; out-of-bounds fill = zero padding) and the filters are K-major rows [co][tap][hi | lo'] (prep type 7).
// Three kind::f16 MMAs per 16 channels: hi x hi -> ACC, hi x lo' -> LO, lo' x hi -> LO (the operands of the second and third
// are the same tiles at a 2 Cin-byte K offset).  The epilogue applies bn2 + ReLU and writes the next stage kernel's plane image.
struct ConvS2Maps {
    CUtensorMap a[4];                  // index = h-parity * 2 + w-parity of (kh - 1, kw - 1)
    CUtensorMap b;
    CUtensorMap r, bsc;                // shortcut: even-pixel view of the raw residual rows, 1x1 filters (has_sc)
};
struct ConvS2Args {
    int cout, hout, n_images, stages;
    int g_in, n_groups_in;             // images per pass / passes per sample of the PRODUCER stage
    const float *packed;
    int64_t ld_packed, bn_off;
    unsigned char *pi_out;             // plane image of the consumer stage (bma_conv_fused16.cuh)
    int pi_G, pi_pitch, pi_img_pos, pi_nplanes, pi_ngroups;
    int64_t pi_pass_bytes;
    // the block's 1x1 stride-2 shortcut (reference models/preresnet.py:122-128, applied to the RAW block input) as a tenth
    // K block into its own accumulator columns: Rs = W_ds * R, written NHWC fp32 for the stage kernel's residual registers
    int has_sc;
    float *rs_out;
};

template <int CIN>
__global__ void __launch_bounds__(CTC_THREADS, 1)
conv3x3s2_f16_kernel(const __grid_constant__ ConvS2Maps maps, const ConvS2Args a) {
    constexpr int ROW_BYTES = CIN * 4;                          // [hi(CIN) | lo'(CIN)] halves
    constexpr uint32_t A_BYTES = 128 * ROW_BYTES;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[CTC_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[CTC_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar, sc_full_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bn_s[128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.y;
    const int WT = a.hout, HT = a.hout >= 16 ? 128 / a.hout : a.hout, NT = 128 / (WT * HT);
    const int tpi = (a.hout * a.hout) / 128;
    int n0, h0;
    if (NT == 1) { n0 = blockIdx.x / tpi; h0 = (blockIdx.x % tpi) * HT; } else { n0 = blockIdx.x * NT; h0 = 0; }

    const uint32_t b_bytes = (uint32_t)a.cout * ROW_BYTES;
    const uint32_t stage_bytes = A_BYTES + b_bytes;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t tmem_cols = (a.has_sc ? 4 : 2) * a.cout;     // [ACC | LO | SC_ACC | SC_LO]: 64 .. 256 columns
    const int n_kb = a.has_sc ? 10 : 9;

    // K blocks in ISSUE order: the shortcut block first (its accumulators are drained by the epilogue warps while the nine
    // taps are still streaming in), then taps 0 .. 8
    const int pass_in = s * a.n_groups_in + n0 / a.g_in, g0 = n0 % a.g_in;
    auto issue_loads = [&](int k) {
        const int st = k % a.stages;
        const uint32_t fb = smem_u32(&full_bar[st]);
        mbar_expect_tx_a(fb, stage_bytes);
        const uint32_t base = smem_base + (uint32_t)st * stage_bytes;
        const int tap = k - a.has_sc;
        if (tap < 0) {                                               // shortcut: pixels (2 oh, 2 ow) of the raw rows
            tma_load_5d_a(base, &maps.r, 0, 0, h0, g0, pass_in, fb);
            tma_load_3d_a(base + A_BYTES, &maps.bsc, 0, 0, s, fb);
            return;
        }
        const int kh = tap / 3, kw = tap - kh * 3;
        const int mi = ((kh + 1) & 1) * 2 + ((kw + 1) & 1);          // parity of (kh - 1, kw - 1)
        tma_load_5d_a(base, &maps.a[mi], 0, (kw - 1) >> 1, h0 + ((kh - 1) >> 1), g0, pass_in, fb);
        tma_load_3d_a(base + A_BYTES, &maps.b, tap * 2 * CIN, 0, s, fb);
    };
    const int n_early = a.stages < n_kb ? a.stages : n_kb;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        mbar_init(&sc_full_bar, 1);
        fence_barrier_init();
        // a CTA lives for one tile: the first round of loads leaves before the CTA-wide sync (TMEM allocation, BatchNorm
        // constants) instead of after it -- the stages are empty by construction
        for (int k = 0; k < n_early; ++k) issue_loads(k);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, tmem_cols);
    if (warp >= 2) {
        const float *bn = a.packed + (int64_t)s * a.ld_packed + a.bn_off;
        for (int i = threadIdx.x - 64; i < 2 * a.cout; i += 128) bn_s[i] = __ldg(bn + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            for (int k = n_early; k < n_kb; ++k) {
                const uint32_t ph = (uint32_t)(k / a.stages) & 1u;
                mbar_wait_a(smem_u32(&empty_bar[k % a.stages]), ph ^ 1u);
                issue_loads(k);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_f16_idesc(128, a.cout);
            // a row = [hi(16 ch) | lo'(16 ch)] per 16-channel group: hi of K step ks at 64 ks bytes, lo' 32 bytes further
            for (int k = 0; k < n_kb; ++k) {
                const int st = k % a.stages;
                const uint32_t ph = (uint32_t)(k / a.stages) & 1u;
                const int tap = k - a.has_sc;
                mbar_wait_a(smem_u32(&full_bar[st]), ph);
                tc_fence_after();
                const uint32_t base = smem_base + (uint32_t)st * stage_bytes;
                const uint64_t d_a = make_kmajor_desc<ROW_BYTES>(base), d_b = make_kmajor_desc<ROW_BYTES>(base + A_BYTES);
                const uint32_t d0 = tmem_base + (tap < 0 ? 2 * a.cout : 0);       // shortcut block: its own accumulators
#pragma unroll
                for (int ks = 0; ks < CIN / 16; ++ks) {
                    const uint64_t kh = (uint64_t)(4 * ks), kl = kh + 2;          // 16-byte descriptor units
                    const uint32_t acc = (tap > 0 || ks != 0) ? 1u : 0u;          // first MMA of the shortcut block / of tap 0
                    umma_f16(d0, d_a + kh, d_b + kh, idesc, acc);
                    umma_f16(d0 + a.cout, d_a + kh, d_b + kl, idesc, acc);
                    umma_f16(d0 + a.cout, d_a + kl, d_b + kh, idesc, 1);
                }
                umma_commit(smem_u32(&empty_bar[st]));
                if (tap < 0) umma_commit(smem_u32(&sc_full_bar));
            }
            umma_commit(smem_u32(&tmem_full_bar));
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int w = r % WT, h = (r / WT) % HT, nl = r / (WT * HT);
        const int n = n0 + nl;
        const bool valid = n < a.n_images;
        unsigned char *pi_px = nullptr;
        if (valid) {
            const int grp = n / a.pi_G, g = n - grp * a.pi_G;
            pi_px = a.pi_out + ((int64_t)s * a.pi_ngroups + grp) * a.pi_pass_bytes +
                    (int64_t)((g * (a.hout + 1) + (h0 + h)) * a.pi_pitch + w) * 16;
        }
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        if (a.has_sc) {                                              // drained while the nine taps are still in flight
            mbar_wait_a(smem_u32(&sc_full_bar), 0);
            tc_fence_after();
            float *rp = a.rs_out + ((((int64_t)s * a.n_images + n) * a.hout + (h0 + h)) * a.hout + w) * a.cout;
            for (int c0 = 0; c0 < a.cout; c0 += 16) {
                uint32_t ra[16], rl[16];
                tmem_ld<16>(tl + (uint32_t)(2 * a.cout + c0), ra);
                tmem_ld<16>(tl + (uint32_t)(3 * a.cout + c0), rl);
                tmem_ld_wait();
                if (!valid) continue;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)                   // operand rows were R / 16
                        v[e] = fmaf(__uint_as_float(rl[i + e]), kLoUnscale, __uint_as_float(ra[i + e])) * kActUp;
                    *reinterpret_cast<float4 *>(rp + c0 + i) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        mbar_wait_a(smem_u32(&tmem_full_bar), 0);
        tc_fence_after();
        for (int c0 = 0; c0 < a.cout; c0 += 16) {
            uint32_t ra[16], rl[16];
            tmem_ld<16>(tl + (uint32_t)c0, ra);
            tmem_ld<16>(tl + (uint32_t)(a.cout + c0), rl);
            tmem_ld_wait();
            if (!valid) continue;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = c0 + i + 2 * jj;
                    // operands were y / 16: conv = 16 (acc + lo 2^-11); the next planes hold relu(bn2(conv)) / 16
                    const float v0 = fmaf(__uint_as_float(rl[i + 2 * jj]), kLoUnscale, __uint_as_float(ra[i + 2 * jj])) * kActUp;
                    const float v1 = fmaf(__uint_as_float(rl[i + 2 * jj + 1]), kLoUnscale, __uint_as_float(ra[i + 2 * jj + 1])) * kActUp;
                    split_h2(relu_nan(fmaf(bn_s[c], v0, bn_s[a.cout + c])) * kActDown,
                             relu_nan(fmaf(bn_s[c + 1], v1, bn_s[a.cout + c + 1])) * kActDown, hw[jj], lw[jj]);
                }
                const int64_t pl = (int64_t)((c0 + i) >> 3) * a.pi_img_pos * 16;
                *reinterpret_cast<uint4 *>(pi_px + pl) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(pi_px + (int64_t)a.pi_nplanes * a.pi_img_pos * 16 + pl) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Negative result (round 2): a persistent form of this kernel (one CTA per SM walking the (sample, tile) grid, the TMA producer
// 16 / 8 K blocks ahead across tile boundaries, 4 / 2 tiles' accumulators in the 512 TMEM columns, epilogue of tile i under the
// loads of tiles i + 1 ..) was parity-green and SLOWER: 612 vs 595 us per launch.  The one-tile CTAs' idle phases are therefore
// not what holds the kernel at ~4.3 TB/s of DRAM traffic; four co-resident CTAs already overlap them.

// a_rows: the producer stage's TMA-stored rows (stage geometry CIN: H = hin, pitch, positions per pass, images per pass)
template <int CIN>
static int launch_conv_s2_f16(const void *a_rows, const void *r_rows, int hin, int pitch_in, int img_pos_in, int g_in, int sc,
                              int nc, const float *packed, int64_t ld_packed, int64_t w_off, int64_t ds_off, ConvS2Args g,
                              cudaStream_t st) {
    constexpr int ROWB = CIN * 4;
    const int hout = hin / 2, cout = 2 * CIN;
    const int n_groups_in = (nc + g_in - 1) / g_in;
    const int WT = hout, HT = hout >= 16 ? 128 / hout : hout, NT = 128 / (WT * HT);
    ConvS2Maps maps;
    for (int par = 0; par < 4; ++par) {
        const int hp = par >> 1, wp = par & 1;
        const unsigned char *base = reinterpret_cast<const unsigned char *>(a_rows) + (size_t)(hp * pitch_in + wp) * ROWB;
        const uint64_t dims[5] = {(uint64_t)2 * CIN, (uint64_t)(hin / 2), (uint64_t)(hin / 2), (uint64_t)g_in,
                                  (uint64_t)sc * n_groups_in};
        const uint64_t sb[4] = {(uint64_t)2 * ROWB, (uint64_t)2 * pitch_in * ROWB, (uint64_t)(hin + 1) * pitch_in * ROWB,
                                (uint64_t)img_pos_in * ROWB};
        const uint32_t box[5] = {(uint32_t)2 * CIN, (uint32_t)WT, (uint32_t)HT, (uint32_t)NT, 1};
        if (int rc = make_tensor_map_t(&maps.a[par], base, 5, dims, sb, box, ROWB, 1)) return rc;
        if (par == 0 && r_rows != nullptr)
            if (int rc = make_tensor_map_t(&maps.r, r_rows, 5, dims, sb, box, ROWB, 1)) return rc;
    }
    g.has_sc = r_rows != nullptr ? 1 : 0;
    if (g.has_sc) {
        const uint64_t dims[3] = {(uint64_t)2 * CIN, (uint64_t)cout, (uint64_t)sc};
        const uint64_t sb[2] = {(uint64_t)2 * CIN * 2, (uint64_t)ld_packed * 4};
        const uint32_t box[3] = {(uint32_t)2 * CIN, (uint32_t)cout, 1};
        if (int rc = make_tensor_map_t(&maps.bsc, packed + ds_off, 3, dims, sb, box, ROWB, 1)) return rc;
    } else {
        maps.r = maps.a[0];
        maps.bsc = maps.a[0];
    }
    {
        const uint64_t dims[3] = {(uint64_t)18 * CIN, (uint64_t)cout, (uint64_t)sc};
        const uint64_t sb[2] = {(uint64_t)18 * CIN * 2, (uint64_t)ld_packed * 4};
        const uint32_t box[3] = {(uint32_t)2 * CIN, (uint32_t)cout, 1};
        if (int rc = make_tensor_map_t(&maps.b, packed + w_off, 3, dims, sb, box, ROWB, 1)) return rc;
    }
    g.cout = cout; g.hout = hout; g.n_images = nc; g.g_in = g_in; g.n_groups_in = n_groups_in;
    g.packed = packed; g.ld_packed = ld_packed;
    const size_t stage_bytes = (size_t)128 * ROWB + (size_t)cout * ROWB;
    g.stages = 4;                                   // 10-24 KB per stage: several CTAs per SM hide the per-tile latencies
    const size_t smem = (size_t)g.stages * stage_bytes + 1024;
    const int tiles = NT == 1 ? nc * ((hout * hout) / 128) : (nc + NT - 1) / NT;
    URSA_CUDA(cudaFuncSetAttribute(conv3x3s2_f16_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3s2_f16_kernel<CIN><<<dim3(tiles, sc), CTC_THREADS, smem, st>>>(maps, g);
    URSA_LAUNCH_CHECK("conv3x3s2_f16_kernel");
    return URSA_OK;
}

size_t preresnet_workspace_tcgen05(int S, int64_t N, int depth, int C) {
    NetPlan pl;
    if (!build_plan(depth, C, pl, 1)) return 0;
    return tc_chunking(S, N, pl).total;
}

int preresnet_forward_tcgen05(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf, int S, const float *x,
                              int64_t N, int depth, int C, float *proba_sum, float *entropy_sum, float *logits_out,
                              double gamma, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    static thread_local NetPlan pl;
    if (!build_plan(depth, C, pl, 1)) {
        set_error("ursa_bma_preresnet_forward: unsupported depth %d (BasicBlock PreResNet: depth = 6n+2, 8..38)", depth);
        return URSA_ERR_UNSUPPORTED;
    }
    URSA_REQUIRE(ld_bank >= pl.D, "ursa_bma_preresnet_forward: ld_bank (%lld) < D (%lld)", (long long)ld_bank, (long long)pl.D);
    URSA_REQUIRE(ld_buf >= pl.NB, "ursa_bma_preresnet_forward: ld_buf (%lld) < %lld", (long long)ld_buf, (long long)pl.NB);
    const TcChunking ck = tc_chunking(S, N, pl);
    URSA_REQUIRE(workspace_bytes >= ck.total, "ursa_bma_preresnet_forward: workspace too small");
    char *wsb = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    float *Ra = reinterpret_cast<float *>(wsb);
    float *Rb = reinterpret_cast<float *>(wsb + ck.raw_bytes);
    float *Rs = reinterpret_cast<float *>(wsb + 2 * ck.raw_bytes);
    float *A1h = reinterpret_cast<float *>(wsb + 3 * ck.raw_bytes), *A1l = reinterpret_cast<float *>(wsb + 4 * ck.raw_bytes);
    float *A2h = reinterpret_cast<float *>(wsb + 5 * ck.raw_bytes), *A2l = reinterpret_cast<float *>(wsb + 6 * ck.raw_bytes);
    float *packed = reinterpret_cast<float *>(wsb + 7 * ck.raw_bytes);
    float *logits = reinterpret_cast<float *>(wsb + 7 * ck.raw_bytes + ck.packed_bytes);
    const int n = pl.n_blocks;

    for (int s0 = 0; s0 < S; s0 += ck.sc) {
        const int sc = (S - s0 < ck.sc) ? (S - s0) : ck.sc;
        preresnet_prep_kernel<<<dim3(pl.table.n, sc), 256, 0, st>>>(pl.table, bank + (int64_t)s0 * ld_bank, ld_bank,
                                                                    bufbank + (int64_t)s0 * ld_buf, ld_buf, packed,
                                                                    pl.packed_floats);
        URSA_LAUNCH_CHECK("preresnet_prep_kernel");
        for (int64_t i0 = 0; i0 < N; i0 += ck.nc) {
            const int nc = (int)((N - i0 < ck.nc) ? (N - i0) : ck.nc);
            // stem: R = conv(x), A1 = split(relu(bn1(R)))
            stem_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(x + i0 * 3 * 32 * 32, packed, pl.packed_floats, pl.conv1_w,
                                                           pl.blocks[0][0].bn1, nc, Ra, A1h, A1l, nullptr);
            URSA_LAUNCH_CHECK("stem_nhwc_kernel");
            float *cur = Ra, *nxt = Rb;
            int ch = 16, hw = 32;
            for (int stg = 0; stg < 3; ++stg) {
                for (int b = 0; b < n; ++b) {
                    const NetPlan::Block &B = pl.blocks[stg][b];
                    const bool down = B.ds >= 0;
                    const int cout = down ? ch * 2 : ch, stride = down ? 2 : 1, hout = hw / stride;
                    const float *res = cur;
                    if (down) {
                        shortcut_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(cur, packed, pl.packed_floats, B.ds, ch, cout, hout, nc, Rs, 0);
                        URSA_LAUNCH_CHECK("shortcut_nhwc_kernel");
                        res = Rs;
                    }
                    ConvTcArgs g;
                    // conv1: A1 -> A2 = split(relu(bn2(acc)))
                    g.mode = 0; g.bn_off = B.bn2; g.res = nullptr; g.out_raw = nullptr; g.out_hi = A2h; g.out_lo = A2l;
                    if (int rc = launch_conv_tc(A1h, A1l, hw, ch, cout, stride, sc, nc, packed, pl.packed_floats, B.w1, B.w1_lo, g, st))
                        return rc;
                    // conv2: A2 -> R' = acc + shortcut ; A1 = split(relu(bn_next(R')))
                    int64_t bn_next = -1;
                    if (b + 1 < n) bn_next = pl.blocks[stg][b + 1].bn1;
                    else if (stg + 1 < 3) bn_next = pl.blocks[stg + 1][0].bn1;
                    g.mode = 1; g.bn_off = bn_next; g.res = res; g.out_raw = nxt;
                    g.out_hi = bn_next >= 0 ? A1h : nullptr; g.out_lo = bn_next >= 0 ? A1l : nullptr;
                    if (int rc = launch_conv_tc(A2h, A2l, hout, cout, cout, 1, sc, nc, packed, pl.packed_floats, B.w2, B.w2_lo, g, st))
                        return rc;
                    float *t = cur; cur = nxt; nxt = t;
                    ch = cout; hw = hout;
                }
            }
            const int pairs = sc * nc;
            head_nhwc_kernel<<<(pairs + 7) / 8, 256, 0, st>>>(cur, packed, pl.packed_floats, pl.bn_final, pl.fc, nc, pairs, C, logits);
            URSA_LAUNCH_CHECK("head_nhwc_kernel");
            if (int rc = ursa_bma_accumulate(logits, sc, nc, C, (int64_t)nc * C, proba_sum + i0 * C, entropy_sum + i0, gamma, (void *)st))
                return rc;
            if (logits_out)
                URSA_CUDA(cudaMemcpy2DAsync(logits_out + ((int64_t)s0 * N + i0) * C, (size_t)N * C * sizeof(float), logits,
                                            (size_t)nc * C * sizeof(float), (size_t)nc * C * sizeof(float), sc,
                                            cudaMemcpyDeviceToDevice, st));
        }
    }
    return URSA_OK;
}

}  // namespace ursa

// ---- train-mode BatchNorm pass: re-estimation of the running statistics ---------------------------------
// util.bn_update (reference util.py:212-247, called once per SWAG / ESS sample at inference/swag.py:123-124 and
// pca_subspace.py:137) = ONE train-mode forward of the sample over the training set: every BatchNorm normalises with the
// statistics of the current batch and folds them into its running statistics with the cumulative momentum b / (n + b).
// Batch statistics need the whole batch's conv output before the next conv can start, so the fused stage kernels cannot be
// used; the pass runs layer by layer on conv3x3_tc_kernel (3xTF32 tcgen05, mode 2: raw output + a statistics epilogue),
// SAMPLE-BATCHED (grid.y = sample): per conv  conv -> finalize (sums -> this batch's (a, b), running statistics) -> apply
// (split(relu(a v + b)) = the next conv's TMA planes).
namespace ursa {

// per-(sample, batch, channel) sums of v and v^2 of a raw NHWC tensor [S_c][nc][hw][C] (the stem output)
__global__ void __launch_bounds__(256) bn_stats_nhwc_kernel(const float *__restrict__ raw, int C, int hw, int nc, int batch,
                                                            int n_batches, double *__restrict__ stats) {
    // one CTA per (image, sample); thread = (pixel lane, channel): 256 threads = 256 / C pixel lanes
    const int n = blockIdx.x, s = blockIdx.y;
    const int c = threadIdx.x % C, pl = threadIdx.x / C, npl = 256 / C;
    const float *p = raw + ((int64_t)s * nc + n) * hw * C;
    double s1 = 0.0, s2 = 0.0;
    for (int px = pl; px < hw; px += npl) {
        const float v = __ldg(p + (int64_t)px * C + c);
        s1 += v;
        s2 += (double)v * v;
    }
    __shared__ double r1[256], r2[256];
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    if (pl == 0) {
        for (int k = 1; k < npl; ++k) { s1 += r1[k * C + c]; s2 += r2[k * C + c]; }
        double *sp = stats + (((int64_t)s * n_batches + n / batch) * 2) * C + c;
        atomicAdd(sp, s1);
        atomicAdd(sp + C, s2);
    }
}

// sums -> (a, b) of every batch of the chunk and the running statistics (in the sample's buffer row); clears the sums
__global__ void bn_finalize_kernel(double *__restrict__ stats, const float *__restrict__ bank, int64_t ld_bank, int64_t w_off,
                                   int64_t b_off, float *__restrict__ bufbank, int64_t ld_buf, int64_t buf_off, int C, int hw, int nc,
                                   int batch, int n_batches, int64_t n_before, float *__restrict__ ab) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (c >= C) return;
    const float gamma = bank[(int64_t)s * ld_bank + w_off + c], beta = bank[(int64_t)s * ld_bank + b_off + c];
    float *rmp = bufbank + (int64_t)s * ld_buf + buf_off + c, *rvp = rmp + C;
    float rm = *rmp, rv = *rvp;
    int64_t n = n_before;
    if (n == 0) { rm = 0.f; rv = 1.f; }                    // reset_bn (util.py:196-199)
    for (int j = 0; j < n_batches; ++j) {
        const int bj = nc - j * batch < batch ? nc - j * batch : batch;
        if (bj <= 0) break;
        const double cnt = (double)bj * hw;
        double *sp = stats + (((int64_t)s * n_batches + j) * 2) * C;
        const double mean = sp[c] / cnt;
        double var = sp[C + c] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        sp[c] = 0.0;
        sp[C + c] = 0.0;
        const float av = gamma / sqrtf((float)var + 1e-5f);
        float *abp = ab + (((int64_t)s * n_batches + j) * 2) * C;
        abp[c] = av;
        abp[C + c] = beta - (float)mean * av;
        const float mom = (float)((double)bj / (double)(n + bj));          // util.py:239-241
        const float unbiased = (float)(cnt > 1.0 ? var * cnt / (cnt - 1.0) : var);
        rm = (1.f - mom) * rm + mom * (float)mean;
        rv = (1.f - mom) * rv + mom * unbiased;
        n += bj;
    }
    *rmp = rm;
    *rvp = rv;
}

// raw [S_c][nc][hw][C] -> TF32 planes split(relu(a v + b)) with the (a, b) of the pixel's (sample, batch)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float *__restrict__ raw, const float *__restrict__ ab, int C, int hw, int nc,
                                                       int batch, int n_batches, int64_t total4, float *__restrict__ a_hi,
                                                       float *__restrict__ a_lo) {
    const int c4n = C >> 2;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += (int64_t)gridDim.x * 256) {
        const int c = (int)(i % c4n) * 4;
        const int64_t px = i / c4n;
        const int64_t img = px / hw;                       // s * nc + n
        const int s = (int)(img / nc), n = (int)(img - (int64_t)s * nc);
        const float *abp = ab + (((int64_t)s * n_batches + n / batch) * 2) * C;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(raw + px * C + c));
        const float4 a4 = __ldg(reinterpret_cast<const float4 *>(abp + c)), b4 = __ldg(reinterpret_cast<const float4 *>(abp + C + c));
        const float y0 = relu_nan(fmaf(a4.x, v.x, b4.x)), y1 = relu_nan(fmaf(a4.y, v.y, b4.y));
        const float y2 = relu_nan(fmaf(a4.z, v.z, b4.z)), y3 = relu_nan(fmaf(a4.w, v.w, b4.w));
        float4 hv, lv;
        hv.x = rn_tf32(y0); hv.y = rn_tf32(y1); hv.z = rn_tf32(y2); hv.w = rn_tf32(y3);
        lv.x = rn_tf32(y0 - hv.x); lv.y = rn_tf32(y1 - hv.y); lv.z = rn_tf32(y2 - hv.z); lv.w = rn_tf32(y3 - hv.w);
        *reinterpret_cast<float4 *>(a_hi + px * C + c) = hv;
        *reinterpret_cast<float4 *>(a_lo + px * C + c) = lv;
    }
}

struct BnTrainLayout {
    int sc, nc, nb;
    size_t raw_bytes, packed_bytes, stats_bytes, ab_bytes, total;
};
static bool bn_train_layout(int S, int64_t N, int batch, const NetPlan &pl, BnTrainLayout &L) {
    if (batch < 2 || (batch & 1) || batch > kTcChunkImages || N < 1 || S < 1) return false;      // 8 x 8 tiles pair two images of a batch
    L.sc = S < kTcChunkSamples ? S : kTcChunkSamples;
    int64_t nc = (int64_t)batch * (kTcChunkImages / batch);
    if (N < nc) nc = N;
    L.nc = (int)nc;
    L.nb = (int)((nc + batch - 1) / batch);
    L.raw_bytes = (size_t)L.sc * L.nc * 16 * 32 * 32 * sizeof(float);
    L.packed_bytes = (((size_t)L.sc * pl.packed_floats * sizeof(float)) + 1023) & ~(size_t)1023;
    L.stats_bytes = (((size_t)L.sc * L.nb * 2 * 64 * sizeof(double)) + 1023) & ~(size_t)1023;
    L.ab_bytes = (((size_t)L.sc * L.nb * 2 * 64 * sizeof(float)) + 1023) & ~(size_t)1023;
    L.total = 6 * L.raw_bytes + L.packed_bytes + L.stats_bytes + L.ab_bytes + 2048;     // Ra, Rb, Rs, Craw, A hi / lo
    return true;
}

}  // namespace ursa

extern "C" size_t ursa_preresnet_bn_update_workspace(int S, int64_t N, int batch, int depth, int C) {
    using namespace ursa;
    NetPlan pl;
    BnTrainLayout L;
    if (!build_plan(depth, C, pl, 1) || !bn_train_layout(S, N, batch, pl, L)) return 0;
    return L.total;
}

extern "C" int ursa_preresnet_bn_update(const float *bank, int64_t ld_bank, float *bufbank, int64_t ld_buf, int S, const float *x,
                                        int64_t N, int batch, int depth, int C, void *workspace, size_t workspace_bytes,
                                        void *stream) {
    using namespace ursa;
    URSA_REQUIRE(bank && bufbank && x && workspace, "ursa_preresnet_bn_update: null pointer");
    static thread_local NetPlan pl;
    BnTrainLayout L;
    if (!build_plan(depth, C, pl, 1) || !bn_train_layout(S, N, batch, pl, L)) {
        set_error("ursa_preresnet_bn_update: unsupported depth %d / batch %d (depth = 6n+2 in 8..38, even batch 2..%d)", depth, batch,
                  kTcChunkImages);
        return URSA_ERR_UNSUPPORTED;
    }
    URSA_REQUIRE(ld_bank >= pl.D && ld_buf >= pl.NB, "ursa_preresnet_bn_update: ld_bank / ld_buf too small");
    URSA_REQUIRE(workspace_bytes >= L.total, "ursa_preresnet_bn_update: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char *wsb = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    float *Ra = reinterpret_cast<float *>(wsb), *Rb = reinterpret_cast<float *>(wsb + L.raw_bytes);
    float *Rs = reinterpret_cast<float *>(wsb + 2 * L.raw_bytes), *Cr = reinterpret_cast<float *>(wsb + 3 * L.raw_bytes);
    float *Ah = reinterpret_cast<float *>(wsb + 4 * L.raw_bytes), *Al = reinterpret_cast<float *>(wsb + 5 * L.raw_bytes);
    float *packed = reinterpret_cast<float *>(wsb + 6 * L.raw_bytes);
    double *stats = reinterpret_cast<double *>(wsb + 6 * L.raw_bytes + L.packed_bytes);
    float *ab = reinterpret_cast<float *>(wsb + 6 * L.raw_bytes + L.packed_bytes + L.stats_bytes);
    // BatchNorm layers in forward order = the type-2 entries of the prep table (bank offsets of gamma / beta, buffer offset)
    int bn_idx[kMaxLayers], n_bn = 0;
    for (int i = 0; i < pl.table.n; ++i)
        if (pl.table.e[i].type == 2) bn_idx[n_bn++] = i;
    const int n = pl.n_blocks;
    URSA_REQUIRE(n_bn == 6 * n + 1, "ursa_preresnet_bn_update: plan mismatch");

    for (int s0 = 0; s0 < S; s0 += L.sc) {
        const int sc = (S - s0 < L.sc) ? (S - s0) : L.sc;
        const float *bk = bank + (int64_t)s0 * ld_bank;
        float *bf = bufbank + (int64_t)s0 * ld_buf;
        preresnet_prep_kernel<<<dim3(pl.table.n, sc), 256, 0, st>>>(pl.table, bk, ld_bank, bf, ld_buf, packed, pl.packed_floats);
        URSA_LAUNCH_CHECK("preresnet_prep_kernel");
        URSA_CUDA(cudaMemsetAsync(stats, 0, L.stats_bytes, st));
        for (int64_t i0 = 0; i0 < N; i0 += L.nc) {
            const int nc = (int)((N - i0 < L.nc) ? (N - i0) : L.nc);
            const int nb = (nc + batch - 1) / batch;
            int bn_pos = 0;
            auto finalize = [&](int Cc, int hw) -> int {
                const PrepEntry &e = pl.table.e[bn_idx[bn_pos++]];
                bn_finalize_kernel<<<dim3((Cc + 63) / 64, sc), 64, 0, st>>>(stats, bk, ld_bank, e.src, e.src2, bf, ld_buf, e.buf, Cc, hw, nc,
                                                                            batch, L.nb, i0, ab);
                URSA_LAUNCH_CHECK("bn_finalize_kernel");
                return URSA_OK;
            };
            auto apply = [&](const float *raw, int Cc, int hw) -> int {
                const int64_t total4 = (int64_t)sc * nc * hw * Cc / 4;
                int gx = (int)((total4 + 255) / 256);
                if (gx > 148 * 16) gx = 148 * 16;
                bn_apply_kernel<<<gx, 256, 0, st>>>(raw, ab, Cc, hw, nc, batch, L.nb, total4, Ah, Al);
                URSA_LAUNCH_CHECK("bn_apply_kernel");
                return URSA_OK;
            };
            stem_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(x + i0 * 3 * 32 * 32, packed, pl.packed_floats, pl.conv1_w, pl.blocks[0][0].bn1, nc,
                                                           Ra, nullptr, nullptr, nullptr);
            URSA_LAUNCH_CHECK("stem_nhwc_kernel");
            bn_stats_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(Ra, 16, 1024, nc, batch, L.nb, stats);
            URSA_LAUNCH_CHECK("bn_stats_nhwc_kernel");
            float *cur = Ra, *nxt = Rb;
            int ch = 16, hw = 32;
            for (int stg = 0; stg < 3; ++stg)
                for (int b = 0; b < n; ++b) {
                    const NetPlan::Block &B = pl.blocks[stg][b];
                    const bool down = B.ds >= 0;
                    const int cout = down ? ch * 2 : ch, stride = down ? 2 : 1, hout = hw / stride;
                    if (int rc = finalize(ch, hw * hw)) return rc;                      // bn1: statistics of the block input
                    if (int rc = apply(cur, ch, hw * hw)) return rc;
                    const float *res = cur;
                    if (down) {
                        shortcut_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(cur, packed, pl.packed_floats, B.ds, ch, cout, hout, nc, Rs, 0);
                        URSA_LAUNCH_CHECK("shortcut_nhwc_kernel");
                        res = Rs;
                    }
                    ConvTcArgs g;
                    g.mode = 2; g.bn_off = -1; g.res = nullptr; g.out_raw = Cr; g.out_hi = nullptr; g.out_lo = nullptr;
                    g.stats = stats; g.batch = batch; g.n_batches = L.nb;
                    if (int rc = launch_conv_tc(Ah, Al, hw, ch, cout, stride, sc, nc, packed, pl.packed_floats, B.w1, B.w1_lo, g, st)) return rc;
                    if (int rc = finalize(cout, hout * hout)) return rc;                // bn2: statistics of conv1's output
                    if (int rc = apply(Cr, cout, hout * hout)) return rc;
                    g.res = res; g.out_raw = nxt;                                       // conv2 + shortcut -> next block input (+ its statistics)
                    if (int rc = launch_conv_tc(Ah, Al, hout, cout, cout, 1, sc, nc, packed, pl.packed_floats, B.w2, B.w2_lo, g, st)) return rc;
                    float *t = cur; cur = nxt; nxt = t;
                    ch = cout; hw = hout;
                }
            if (int rc = finalize(64, 64)) return rc;                                   // the final bn (no apply: only its statistics matter)
            (void)nb;
        }
    }
    return URSA_OK;
}

namespace ursa {
// ---- fused-stage path (URSA_ALGO_TCGEN05_FUSED): stem -> [stage kernel] -> (shortcut + stride-2 conv) -> [stage kernel] ...
// f16 != 0 (URSA_ALGO_TCGEN05_FUSED_F16): FP16-split stage kernels (bma_conv_fused16.cuh), type-6 filters, and the stride-2
// convs hand their activations over as one plain fp32 plane
// plane images of the three stages for one chunk (FP16-split path): passes = samples x image groups
struct PiLayout {
    size_t off[3], bytes[3], total;
};
static PiLayout pi_layout(int sc, int nc) {
    PiLayout L;
    const size_t pass_bytes[3] = {(size_t)F16Cfg<16>::PASS_BYTES, (size_t)F16Cfg<32>::PASS_BYTES, (size_t)F16Cfg<64>::PASS_BYTES};
    const int G[3] = {F16Cfg<16>::G, F16Cfg<32>::G, F16Cfg<64>::G};
    size_t o = 0;
    for (int i = 0; i < 3; ++i) {
        L.off[i] = o;
        L.bytes[i] = (size_t)sc * ((nc + G[i] - 1) / G[i]) * pass_bytes[i];
        o += (L.bytes[i] + 1023) & ~(size_t)1023;
    }
    L.total = o;
    return L;
}

size_t preresnet_workspace_fused(int S, int64_t N, int depth, int C, int f16) {
    NetPlan pl;
    if (!build_plan(depth, C, pl, f16 ? 3 : 2)) return 0;
    const TcChunking ck = tc_chunking(S, N, pl, true);
    return ck.total + (f16 ? pi_layout(ck.sc, ck.nc).total + 1024 : 0);
}

int preresnet_forward_fused(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf, int S, const float *x,
                            int64_t N, int depth, int C, float *proba_sum, float *entropy_sum, float *logits_out,
                            double gamma, void *workspace, size_t workspace_bytes, int f16, int ws_kept, cudaStream_t st) {
    static thread_local NetPlan pl;
    if (!build_plan(depth, C, pl, f16 ? 3 : 2)) {
        set_error("ursa_bma_preresnet_forward: unsupported depth %d (BasicBlock PreResNet: depth = 6n+2, 8..38)", depth);
        return URSA_ERR_UNSUPPORTED;
    }
    URSA_REQUIRE(ld_bank >= pl.D, "ursa_bma_preresnet_forward: ld_bank (%lld) < D (%lld)", (long long)ld_bank, (long long)pl.D);
    URSA_REQUIRE(ld_buf >= pl.NB, "ursa_bma_preresnet_forward: ld_buf (%lld) < %lld", (long long)ld_buf, (long long)pl.NB);
    const TcChunking ck = tc_chunking(S, N, pl, true);
    const PiLayout pil = pi_layout(ck.sc, ck.nc);
    URSA_REQUIRE(workspace_bytes >= ck.total + (f16 ? pil.total + 1024 : 0), "ursa_bma_preresnet_forward: workspace too small");
    char *wsb = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    unsigned char *pi_base = reinterpret_cast<unsigned char *>(wsb) + ((ck.total - 2048 + 1023) & ~(size_t)1023);
    // The producers only write pixel positions: the pad positions of the plane images must read as zero.  With
    // URSA_ALGO_FLAG_WS_KEPT the caller vouches that they still do (same workspace, same S-chunk and N as its previous call,
    // nothing else wrote to it): 2.7 GB of memset per call at the default chunking.
    if (f16 && !ws_kept) URSA_CUDA(cudaMemsetAsync(pi_base, 0, pil.total, st));
    float *Ra = reinterpret_cast<float *>(wsb);
    float *Rb = reinterpret_cast<float *>(wsb + ck.raw_bytes);
    float *Rs = reinterpret_cast<float *>(wsb + 2 * ck.raw_bytes);
    float *A1h = reinterpret_cast<float *>(wsb + 3 * ck.raw_bytes), *A1l = reinterpret_cast<float *>(wsb + 4 * ck.raw_bytes);
    float *A2h = reinterpret_cast<float *>(wsb + 5 * ck.raw_bytes), *A2l = reinterpret_cast<float *>(wsb + 6 * ck.raw_bytes);
    float *packed = reinterpret_cast<float *>(wsb + 7 * ck.raw_bytes);
    float *logits = reinterpret_cast<float *>(wsb + 7 * ck.raw_bytes + ck.packed_bytes);
    const int n = pl.n_blocks;
    URSA_REQUIRE(2 * n + 1 <= kFusedMaxConvs, "ursa_bma_preresnet_forward: depth %d exceeds the fused-stage chain length", depth);

    for (int s0 = 0; s0 < S; s0 += ck.sc) {
        const int sc = (S - s0 < ck.sc) ? (S - s0) : ck.sc;
        {
            ProfScope ps(URSA_PROF_PREP, st);
            preresnet_prep_kernel<<<dim3(pl.table.n, sc), 256, 0, st>>>(pl.table, bank + (int64_t)s0 * ld_bank, ld_bank,
                                                                        bufbank + (int64_t)s0 * ld_buf, ld_buf, packed,
                                                                        pl.packed_floats);
            URSA_LAUNCH_CHECK("preresnet_prep_kernel");
        }
        for (int64_t i0 = 0; i0 < N; i0 += ck.nc) {
            const int nc = (int)((N - i0 < ck.nc) ? (N - i0) : ck.nc);
            {
                ProfScope ps(URSA_PROF_STEM, st);
                if (f16) {      // conv1 runs inside the stage-1 kernel: only the images' plane form is prepared here
                    image_planes_kernel<<<(nc * 1024 + 255) / 256, 256, 0, st>>>(x + i0 * 3 * 32 * 32, nc, pi_base + pil.off[0]);
                    URSA_LAUNCH_CHECK("image_planes_kernel");
                } else {
                    stem_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(x + i0 * 3 * 32 * 32, packed, pl.packed_floats, pl.conv1_w,
                                                                   pl.blocks[0][0].bn1, nc, Ra, nullptr, nullptr, nullptr);
                    URSA_LAUNCH_CHECK("stem_nhwc_kernel");
                }
            }
            float *cur = Ra, *nxt = Rb;          // residual stream in / out of the current stage
            int ch = 16, hw = 32;
            for (int stg = 0; stg < 3; ++stg) {
                FusedStageArgs g;
                g.packed = packed; g.ld_packed = pl.packed_floats; g.n_images = nc; g.n_samples = sc;
                g.n_convs = 0;
                g.pi_in = f16 ? pi_base + pil.off[stg] : nullptr;
                if (stg == 0) {
                    g.bn_in_off = pl.blocks[0][0].bn1;
                    g.r_in = cur; g.a_in_hi = g.a_in_lo = nullptr;
                    if (f16) {  // conv1 (3 -> 16) as conv 0 of the chain: residual update from zero, then bn1 of block 0
                        g.r_in = nullptr;
                        g.pi_per_image = 1;
                        g.w_off[0] = pl.conv1_w16; g.mode[0] = 1; g.bn_off[0] = pl.blocks[0][0].bn1; g.n_convs = 1;
                    }
                } else {
                    // transition block: 1x1 stride-2 shortcut on the raw stream, stride-2 conv1 on the layer-wise kernel
                    const NetPlan::Block &B0 = pl.blocks[stg][0];
                    if (!f16) {     // FP16-split path: the shortcut is a tenth K block of conv3x3s2_f16_kernel (below)
                        ProfScope ps(URSA_PROF_SHORTCUT, st);
                        shortcut_nhwc_kernel<<<dim3(nc, sc), 256, 0, st>>>(cur, packed, pl.packed_floats, B0.ds, ch, 2 * ch, hw / 2, nc, Rs, 0);
                        URSA_LAUNCH_CHECK("shortcut_nhwc_kernel");
                    }
                    if (f16) {
                        // conv1 of the transition block: FP16-split implicit GEMM on the rows the previous stage kernel
                        // stored by TMA (in the A1h / A1l region), writes this stage kernel's plane image
                        ConvS2Args c2;
                        c2.bn_off = B0.bn2;
                        const int Gn = stg == 1 ? F16Cfg<32>::G : F16Cfg<64>::G;
                        c2.pi_out = pi_base + pil.off[stg];
                        c2.pi_G = Gn;
                        c2.pi_pitch = stg == 1 ? F16Cfg<32>::PITCH : F16Cfg<64>::PITCH;
                        c2.pi_img_pos = stg == 1 ? F16Cfg<32>::IMG_POS : F16Cfg<64>::IMG_POS;
                        c2.pi_nplanes = stg == 1 ? F16Cfg<32>::NPLANES : F16Cfg<64>::NPLANES;
                        c2.pi_ngroups = (nc + Gn - 1) / Gn;
                        c2.pi_pass_bytes = stg == 1 ? F16Cfg<32>::PASS_BYTES : F16Cfg<64>::PASS_BYTES;
                        c2.rs_out = Rs;                     // + the block's 1x1 shortcut on the raw rows (A2h region)
                        ProfScope ps(URSA_PROF_CONV_S2, st);
                        int rc;
                        if (stg == 1)
                            rc = launch_conv_s2_f16<16>(A1h, A2h, 32, F16Cfg<16>::PITCH, F16Cfg<16>::IMG_POS, F16Cfg<16>::G, sc, nc,
                                                        packed, pl.packed_floats, B0.w1, B0.ds16, c2, st);
                        else
                            rc = launch_conv_s2_f16<32>(A1h, A2h, 16, F16Cfg<32>::PITCH, F16Cfg<32>::IMG_POS, F16Cfg<32>::G, sc, nc,
                                                        packed, pl.packed_floats, B0.w1, B0.ds16, c2, st);
                        if (rc) return rc;
                    } else {
                    ConvTcArgs c1;
                    c1.mode = 0; c1.bn_off = B0.bn2; c1.res = nullptr; c1.out_raw = nullptr; c1.out_hi = A2h;
                    c1.out_lo = A2l;
                    {
                        ProfScope ps(URSA_PROF_CONV_S2, st);
                        if (int rc = launch_conv_tc(A1h, A1l, hw, ch, 2 * ch, 2, sc, nc, packed, pl.packed_floats, B0.w1, B0.w1_lo, c1, st))
                            return rc;
                    }
                    }
                    ch *= 2; hw /= 2;
                    g.bn_in_off = -1;
                    g.r_in = Rs; g.a_in_hi = A2h; g.a_in_lo = A2l;
                }
                for (int b = 0; b < n; ++b) {
                    const NetPlan::Block &B = pl.blocks[stg][b];
                    if (!(stg > 0 && b == 0)) {                 // conv1 of the transition block ran above
                        g.w_off[g.n_convs] = B.w1; g.mode[g.n_convs] = 0; g.bn_off[g.n_convs] = B.bn2; ++g.n_convs;
                    }
                    int64_t bn_next = -1;
                    if (b + 1 < n) bn_next = pl.blocks[stg][b + 1].bn1;
                    else if (stg + 1 < 3) bn_next = pl.blocks[stg + 1][0].bn1;
                    g.w_off[g.n_convs] = B.w2; g.mode[g.n_convs] = 1; g.bn_off[g.n_convs] = bn_next; ++g.n_convs;
                }
                g.r_out = nxt;
                g.r_out_compact = (f16 && stg < 2) ? 1 : 0;     // stages 1 / 2: only the shortcut conv reads the raw stream
                g.a_out_hi = stg < 2 ? A1h : nullptr;
                g.a_out_lo = stg < 2 ? A1l : nullptr;
                if (f16 && stg < 2) {
                    // rows [pass][position][hi(C) | lo'(C)] halves for the next stride-2 conv, as a 2-D fp32 view C x positions
                    const int Cs = stg == 0 ? 16 : 32, Gs = stg == 0 ? F16Cfg<16>::G : F16Cfg<32>::G;
                    const int img_pos = stg == 0 ? F16Cfg<16>::IMG_POS : F16Cfg<32>::IMG_POS;
                    const uint64_t dims[2] = {(uint64_t)Cs, (uint64_t)sc * ((nc + Gs - 1) / Gs) * img_pos};
                    const uint64_t sb[1] = {(uint64_t)Cs * 4};
                    const uint32_t box[2] = {16, 32};               // one warp's rows x one 16-channel group's [hi | lo'] = 64 B
                    if (int rc = make_tensor_map(&g.out_map, A1h, 2, dims, sb, box, 64)) return rc;
                    g.has_out_map = 1;
                    if (int rc = make_tensor_map(&g.rrow_map, A2h, 2, dims, sb, box, 64)) return rc;   // raw residual rows
                    g.has_rrow_map = 1;
                }
                int rc;
                {
                    ProfScope ps(URSA_PROF_STAGE_C16 + stg, st);
                    if (f16) rc = stg == 0 ? launch_stage16<16>(g, st) : (stg == 1 ? launch_stage16<32>(g, st) : launch_stage16<64>(g, st));
                    else rc = stg == 0 ? launch_stage<16>(g, st) : (stg == 1 ? launch_stage<32>(g, st) : launch_stage<64>(g, st));
                }
                if (rc) return rc;
                float *t = cur; cur = nxt; nxt = t;
            }
            const int pairs = sc * nc;
            if (logits_out == nullptr) {
                ProfScope ps(URSA_PROF_HEAD, st);        // head + softmax-average + entropy in ONE kernel, logits stay in registers
                if (int rc = launch_head_bma(cur, packed, pl.packed_floats, pl.bn_final, pl.fc, nc, sc, C, proba_sum + i0 * C,
                                             entropy_sum + i0, gamma, st))
                    return rc;
            } else {
                ProfScope ps(URSA_PROF_HEAD, st);
                head_nhwc_kernel<<<(pairs + 7) / 8, 256, 0, st>>>(cur, packed, pl.packed_floats, pl.bn_final, pl.fc, nc, pairs, C, logits);
                URSA_LAUNCH_CHECK("head_nhwc_kernel");
                if (int rc = ursa_bma_accumulate(logits, sc, nc, C, (int64_t)nc * C, proba_sum + i0 * C, entropy_sum + i0, gamma, (void *)st))
                    return rc;
            }
            if (logits_out)
                URSA_CUDA(cudaMemcpy2DAsync(logits_out + ((int64_t)s0 * N + i0) * C, (size_t)N * C * sizeof(float), logits,
                                            (size_t)nc * C * sizeof(float), (size_t)nc * C * sizeof(float), sc,
                                            cudaMemcpyDeviceToDevice, st));
        }
    }
    return URSA_OK;
}

}  // namespace ursa
