// K3 (MLP) on 5th-generation tensor cores: sample-batched 3xTF32 GEMM, TMA-fed, accumulators in TMEM.
//
// Why 3xTF32: BMA probabilities must match the reference's fp32 forward to 1e-5; one TF32 pass has ~1e-3 relative
// logit error.  Every operand is split a = hi + lo with hi = rn_tf32(a), lo = rn_tf32(a - hi) (exact residual), and
//     D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          (fp32 accumulation in TMEM; the lo*lo term ~2^-22 is dropped)
// which restores ~fp32 accuracy for 3x the MMA work -- still far above the CUDA-core roofline.
//
// Kernel anatomy (one 128 x BN output tile of one sample per CTA, 192 threads):
//   warp 0      TMA producer: per k-block (32 fp32 = one 128-byte swizzle row) four 3-D tiled loads
//               (A_hi, A_lo: 128 x 32; B_hi, B_lo: BN x 32) into a 4-stage shared-memory ring, mbarrier complete_tx
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues 4 (UMMA_K = 8) x 3 tcgen05.mma.kind::tf32 per
//               stage; tcgen05.commit frees the stage / signals the epilogue
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 16 columns) -> + bias -> ReLU -> split hi/lo -> global (the next
//               layer's TMA source), or plain fp32 logits for the last layer
// TWO-LEVEL ACCUMULATION (as in bma_wrn_tc.cu): the tensor core adds into TMEM with truncated alignment, a bias that grows
// with the length of an MMA chain -- one chain over K = 784 (300 MMAs) put the probabilities 13x further from an fp64
// forward than PyTorch's fp32 is (7.1e-6 vs 5.4e-7 at |logit| <= 2.6).  A chain now covers TC_SEG K blocks (48 MMAs);
// up to 4 TMEM accumulators rotate per segment and the epilogue warps drain each finished segment into fp32 registers
// (round-to-nearest) while the next segments' MMAs run.
// Operand layout: K-major rows of 128 bytes, SWIZZLE_128B in both the tensor maps and the UMMA descriptors.
#include <stdlib.h>

#include "tc_common.cuh"

namespace ursa {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 4, TC_THREADS = 192;
constexpr int TC_SEG = 4, TC_MAX_TBUF = 4, TC_EPI_CHUNKS = 8;          // BN <= 128 = 8 chunks of 16 columns
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;     // 16 KB per A tile

struct TcGemmArgs {
    const float *bias;            // per-sample bias vector: bias + s * bias_stride
    int64_t bias_stride;
    float *out_hi, *out_lo;       // split outputs [S][M][ld_out]; out_lo == nullptr: plain fp32 output in out_hi
    int64_t out_batch_stride;
    int ld_out;
    int64_t M;
    int n_valid;                  // real output features (bias guard / plain-store guard)
    int BN;                       // N tile (multiple of 16, <= 256)
    int k_blocks;                 // Kp / 32
    int a_batched;                // 0: A shared by all samples (layer 1 input)
    int b_shared = 0;             // 1: B shared by all samples (the data matrix of the MLP weight-gradient GEMM)
    int relu;
    int stages;                   // smem ring depth (<= TC_STAGES)
    uint32_t tmem_cols;
    int seg, ntbuf;               // K blocks per TMEM accumulation segment; rotating accumulators (column pitch BN)
    // chain-batched MLP gradient (ursa_hmc_mlp_grad): optional extras of the epilogue
    const float *mask = nullptr;  // [S][M][ld_mask]: v *= (mask > 0)  (ReLU derivative of the layer's forward activation)
    int ld_mask = 0;
    int64_t mask_batch_stride = 0;
    float *outT_hi = nullptr, *outT_lo = nullptr;   // TRANSPOSED copy [S][features][ld_t] (lo null: plain fp32): the operand of
    int ld_t = 0;                                   // the weight-gradient GEMMs, which contract over the rows of this one;
    int64_t outT_batch_stride = 0;                  // rows M .. ld_t of it are written as zeros
    int t_valid = 0;                                // plain transposed store: features [0, t_valid) only
};

__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const TcGemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[TC_MAX_TBUF];
    __shared__ __align__(8) uint64_t tempty_bar[TC_MAX_TBUF];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, n_blk = blockIdx.y, s = blockIdx.z;
    const uint32_t ntbuf = (uint32_t)a.ntbuf;
    const uint32_t b_bytes = (uint32_t)a.BN * TC_BK * 4;
    const uint32_t stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzle-128B tiles need 1024-byte alignment

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < TC_MAX_TBUF; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);                                    // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_a_lo);
        tma_prefetch_desc(&tm_b_hi);
        tma_prefetch_desc(&tm_b_lo);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer (one elected lane) =====
        if (elect_one()) {
            const int a_b = a.a_batched ? s : 0;
            uint32_t st = 0, ph = 0;                                         // ring position / phase (no division on the issue path)
            for (int kb = 0; kb < a.k_blocks; ++kb) {
                mbar_wait_a(smem_u32(&empty_bar[st]), ph ^ 1u);              // slot free (first pass: passes at once)
                const uint32_t fb = smem_u32(&full_bar[st]);
                mbar_expect_tx_a(fb, stage_bytes);
                const uint32_t base = smem_base + st * stage_bytes;
                const int k0 = kb * TC_BK;
                tma_load_3d_a(base, &tm_a_hi, k0, m_blk * TC_BM, a_b, fb);
                tma_load_3d_a(base + TC_A_BYTES, &tm_a_lo, k0, m_blk * TC_BM, a_b, fb);
                tma_load_3d_a(base + 2 * TC_A_BYTES, &tm_b_hi, k0, n_blk * a.BN, a.b_shared ? 0 : s, fb);
                tma_load_3d_a(base + 2 * TC_A_BYTES + b_bytes, &tm_b_lo, k0, n_blk * a.BN, a.b_shared ? 0 : s, fb);
                if (++st == (uint32_t)a.stages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (elect_one()) {
            const uint32_t idesc = make_tf32_idesc(TC_BM, a.BN);
            uint32_t buf = 0, bph = 0, st = 0, ph = 0;
            for (int kb0 = 0; kb0 < a.k_blocks; kb0 += a.seg) {
                mbar_wait_a(smem_u32(&tempty_bar[buf]), bph ^ 1u);           // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)a.BN;
                const int kb1 = kb0 + a.seg < a.k_blocks ? kb0 + a.seg : a.k_blocks;
                uint32_t acc = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait_a(smem_u32(&full_bar[st]), ph);                // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t base = smem_base + st * stage_bytes;
                    const uint64_t d_ahi = make_kmajor_desc<128>(base), d_alo = make_kmajor_desc<128>(base + TC_A_BYTES);
                    const uint64_t d_bhi = make_kmajor_desc<128>(base + 2 * TC_A_BYTES);
                    const uint64_t d_blo = make_kmajor_desc<128>(base + 2 * TC_A_BYTES + b_bytes);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)((k * 32) >> 4);     // advance 32 bytes inside the swizzle row
                        umma_tf32(d_tmem, d_alo + koff, d_bhi + koff, idesc, acc);
                        acc = 1;
                        umma_tf32(d_tmem, d_ahi + koff, d_blo + koff, idesc, 1);
                        umma_tf32(d_tmem, d_ahi + koff, d_bhi + koff, idesc, 1);
                    }
                    umma_commit(smem_u32(&empty_bar[st]));                   // frees the smem slot when the MMAs retire
                    if (++st == (uint32_t)a.stages) { st = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(&tfull_bar[buf]));                      // segment complete -> epilogue drains it
                if (++buf == ntbuf) { buf = 0; bph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        float accr[TC_EPI_CHUNKS][16];
#pragma unroll
        for (int j = 0; j < TC_EPI_CHUNKS; ++j)
#pragma unroll
            for (int e = 0; e < 16; ++e) accr[j][e] = 0.f;
        {
            uint32_t buf = 0, bph = 0;
            for (int kb0 = 0; kb0 < a.k_blocks; kb0 += a.seg) {
                mbar_wait_a(smem_u32(&tfull_bar[buf]), bph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)a.BN;
#pragma unroll
                for (int j = 0; j < TC_EPI_CHUNKS; ++j) {
                    if (j * 16 < a.BN) {
                        uint32_t rr[16];
                        tmem_ld16(taddr + (uint32_t)(j * 16), rr);
#pragma unroll
                        for (int e = 0; e < 16; ++e) accr[j][e] += __uint_as_float(rr[e]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                if (++buf == ntbuf) { buf = 0; bph ^= 1u; }
            }
        }
        const int64_t row = (int64_t)m_blk * TC_BM + q * 32 + lane;
        const float *bias = a.bias + (int64_t)s * a.bias_stride;
        const bool split = a.out_lo != nullptr;
        float *ohi = a.out_hi ? a.out_hi + (int64_t)s * a.out_batch_stride + row * a.ld_out : nullptr;
        float *olo = split ? a.out_lo + (int64_t)s * a.out_batch_stride + row * a.ld_out : nullptr;
        const float *mrow = a.mask ? a.mask + (int64_t)s * a.mask_batch_stride + row * a.ld_mask : nullptr;
        float *thi = a.outT_hi ? a.outT_hi + (int64_t)s * a.outT_batch_stride + row : nullptr;
        float *tlo = a.outT_lo ? a.outT_lo + (int64_t)s * a.outT_batch_stride + row : nullptr;
#pragma unroll
        for (int j = 0; j < TC_EPI_CHUNKS; ++j) {
            const int c0 = j * 16;
            if (c0 >= a.BN) continue;
            const float (&r)[16] = accr[j];
            const int col0 = n_blk * a.BN + c0;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int col = col0 + i;
                float x = r[i] + (col < a.n_valid ? __ldg(bias + col) : 0.f);
                if (a.relu) x = fmaxf(x, 0.f);
                v[i] = x;
            }
            if (mrow != nullptr && row < a.M) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (col0 + i < a.n_valid && !(__ldg(mrow + col0 + i) > 0.f)) v[i] = 0.f;
            }
            if (row < a.M && ohi != nullptr) {
                if (split) {
                    if (col0 + 16 <= a.ld_out) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            float4 h, l;
                            h.x = rn_tf32(v[i]); h.y = rn_tf32(v[i + 1]); h.z = rn_tf32(v[i + 2]); h.w = rn_tf32(v[i + 3]);
                            l.x = rn_tf32(v[i] - h.x); l.y = rn_tf32(v[i + 1] - h.y);
                            l.z = rn_tf32(v[i + 2] - h.z); l.w = rn_tf32(v[i + 3] - h.w);
                            *reinterpret_cast<float4 *>(ohi + col0 + i) = h;
                            *reinterpret_cast<float4 *>(olo + col0 + i) = l;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (col0 + i < a.n_valid) ohi[col0 + i] = v[i];
                }
            }
            if (thi != nullptr && row < a.ld_t) {
                // transposed copy: lanes are consecutive rows, so every store of a warp is one contiguous 128-byte segment
                const bool live = row < a.M;
                if (tlo != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float x = live ? v[i] : 0.f, h = rn_tf32(x);
                        thi[(int64_t)(col0 + i) * a.ld_t] = h;
                        tlo[(int64_t)(col0 + i) * a.ld_t] = rn_tf32(x - h);
                    }
                } else if (live) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (col0 + i < a.t_valid) thi[(int64_t)(col0 + i) * a.ld_t] = v[i];
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

// fp32 [rows, cols] (ld, per-batch stride) -> zero-padded hi / lo planes [batch][rows_p][cols_p]
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ src, int64_t ld_src, int64_t src_batch_stride,
                                                         int rows, int cols, float *__restrict__ hi, float *__restrict__ lo,
                                                         int rows_p, int cols_p) {
    const int b = blockIdx.y;
    const int64_t total = (int64_t)rows_p * cols_p;
    const float *sb = src + (int64_t)b * src_batch_stride;
    float *hb = hi + (int64_t)b * total, *lb = lo + (int64_t)b * total;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c = (int)(i % cols_p);
        const int64_t r = i / cols_p;
        float x = 0.f;
        if (r < rows && c < cols) x = __ldg(sb + r * ld_src + c);
        const float h = rn_tf32(x);
        hb[i] = h;
        lb[i] = rn_tf32(x - h);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
// 3-D fp32 tensor [batch][rows][cols_p] (cols innermost), box = [1][box_rows][32], 128-byte swizzle
static int make_tmap(CUtensorMap *tm, const float *base, int64_t cols_p, int64_t rows, int64_t batch, int box_rows) {
    const uint64_t dims[3] = {(uint64_t)cols_p, (uint64_t)rows, (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)cols_p * 4, (uint64_t)cols_p * 4 * (uint64_t)rows};
    const uint32_t box[3] = {TC_BK, (uint32_t)box_rows, 1};
    return make_tensor_map(tm, base, 3, dims, strides, box, 128);
}

static int round_up_i(int v, int m) { return (v + m - 1) / m * m; }

// N tile: multiple of 16, <= 128, minimal padding (ties -> larger tile)
static int pick_bn(int n_out) {
    const int nr = round_up_i(n_out, 16);
    if (nr <= 128) return nr;
    int best = 128, best_waste = round_up_i(nr, 128) - nr;
    for (int bn = 112; bn >= 64; bn -= 16) {
        const int w = round_up_i(nr, bn) - nr;
        if (w < best_waste) { best = bn; best_waste = w; }
    }
    return best;
}

struct TcPlan {
    int K1p, K2p, BN1, BN3, Np1, Np3, sc;
    size_t x_plane, w1_plane, w2_plane, w3_plane, h_plane, logit_plane;   // floats per (sample) plane
    size_t total_bytes;
};

static TcPlan tc_plan(int S, int64_t N, int in_dim, int hidden, int C) {
    TcPlan p;
    p.K1p = round_up_i(in_dim, TC_BK);
    p.K2p = round_up_i(hidden, TC_BK);
    p.BN1 = pick_bn(hidden);
    p.BN3 = pick_bn(C);
    p.Np1 = round_up_i(round_up_i(hidden, 16), p.BN1);
    p.Np3 = round_up_i(round_up_i(C, 16), p.BN3);
    p.x_plane = (size_t)N * p.K1p;
    p.w1_plane = (size_t)p.Np1 * p.K1p;
    p.w2_plane = (size_t)p.Np1 * p.K2p;
    p.w3_plane = (size_t)p.Np3 * p.K2p;
    p.h_plane = (size_t)N * p.K2p;
    p.logit_plane = ((size_t)N * C + 3) & ~(size_t)3;
    const size_t per_sample = 2 * (p.w1_plane + p.w2_plane + p.w3_plane) + 4 * p.h_plane + p.logit_plane;
    size_t sc = (size_t)(768ull << 20) / (per_sample * 4 + 1);
    if (sc < 1) sc = 1;
    if (sc > (size_t)S) sc = S;
    if (sc > 16) sc = 16;
    p.sc = (int)sc;
    p.total_bytes = (2 * p.x_plane + (size_t)p.sc * per_sample) * sizeof(float) + 1024;
    return p;
}

static int launch_tc_gemm(const float *a_hi, const float *a_lo, int64_t a_rows, int64_t a_batch, int Kp, const float *b_hi,
                          const float *b_lo, int Np, int BN, int batch, TcGemmArgs g, cudaStream_t st) {
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (int rc = make_tmap(&ta_hi, a_hi, Kp, a_rows, a_batch, TC_BM)) return rc;
    if (int rc = make_tmap(&ta_lo, a_lo, Kp, a_rows, a_batch, TC_BM)) return rc;
    if (int rc = make_tmap(&tb_hi, b_hi, Kp, Np, g.b_shared ? 1 : batch, BN)) return rc;
    if (int rc = make_tmap(&tb_lo, b_lo, Kp, Np, g.b_shared ? 1 : batch, BN)) return rc;
    g.BN = BN;
    g.k_blocks = Kp / TC_BK;
    g.a_batched = a_batch > 1 ? 1 : 0;
    g.seg = TC_SEG;
    if (const char *e = getenv("URSA_MLP_SEG")) { const int v = atoi(e); if (v >= 1) g.seg = v; }   // accuracy / speed experiments
    g.ntbuf = 512 / BN < TC_MAX_TBUF ? 512 / BN : TC_MAX_TBUF;
    uint32_t cols = 32;
    while (cols < (uint32_t)(g.ntbuf * BN)) cols <<= 1;
    g.tmem_cols = cols;
    const size_t stage_bytes = 2 * TC_A_BYTES + 2 * (size_t)BN * TC_BK * 4;
    int stages = (int)((size_t)(226 << 10) / stage_bytes);
    if (stages > TC_STAGES) stages = TC_STAGES;
    if (stages > g.k_blocks) stages = g.k_blocks;
    URSA_REQUIRE(stages >= 1, "mlp_tc_gemm: tile does not fit in shared memory");
    g.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    URSA_CUDA(cudaFuncSetAttribute(mlp_tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((g.M + TC_BM - 1) / TC_BM), (unsigned)(Np / BN), (unsigned)batch);
    mlp_tc_gemm_kernel<<<grid, TC_THREADS, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, g);
    URSA_LAUNCH_CHECK("mlp_tc_gemm_kernel");
    return URSA_OK;
}

size_t mlp_workspace_tcgen05(int S, int64_t N, int in_dim, int hidden, int C) {
    if (in_dim % 4 != 0 || hidden % 4 != 0) return 0;      // TMA needs 16-byte row pitches (covers MLP200/400/600)
    return tc_plan(S, N, in_dim, hidden, C).total_bytes;
}

int mlp_forward_tcgen05(const float *bank, int64_t ld_bank, int S, const float *x, int64_t N, int in_dim, int hidden,
                        int C, float *proba_sum, float *entropy_sum, float *logits_out, double gamma, void *workspace,
                        size_t workspace_bytes, cudaStream_t st) {
    if (in_dim % 4 != 0 || hidden % 4 != 0) {
        set_error("URSA_ALGO_TCGEN05 needs in_dim %% 4 == 0 and hidden %% 4 == 0 (TMA 16-byte pitch); use URSA_ALGO_FFMA");
        return URSA_ERR_UNSUPPORTED;
    }
    const TcPlan p = tc_plan(S, N, in_dim, hidden, C);
    URSA_REQUIRE(workspace_bytes >= p.total_bytes, "ursa_bma_mlp_forward: workspace too small");
    float *ws = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    float *x_hi = ws, *x_lo = x_hi + p.x_plane;
    float *w1_hi = x_lo + p.x_plane, *w1_lo = w1_hi + p.sc * p.w1_plane;
    float *w2_hi = w1_lo + p.sc * p.w1_plane, *w2_lo = w2_hi + p.sc * p.w2_plane;
    float *w3_hi = w2_lo + p.sc * p.w2_plane, *w3_lo = w3_hi + p.sc * p.w3_plane;
    float *h1_hi = w3_lo + p.sc * p.w3_plane, *h1_lo = h1_hi + p.sc * p.h_plane;
    float *h2_hi = h1_lo + p.sc * p.h_plane, *h2_lo = h2_hi + p.sc * p.h_plane;
    float *lg = h2_lo + p.sc * p.h_plane;

    const int64_t oW1 = 0, ob1 = oW1 + (int64_t)hidden * in_dim, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)hidden * hidden,
                  oW3 = ob2 + hidden, ob3 = oW3 + (int64_t)C * hidden;
    auto split = [&](const float *src, int64_t ld, int64_t bstride, int rows, int cols, float *hi, float *lo, int rows_p,
                     int cols_p, int batch) -> int {
        const int64_t total = (int64_t)rows_p * cols_p;
        int gx = (int)((total + 255) / 256);
        if (gx > 148 * 8) gx = 148 * 8;
        split_tf32_kernel<<<dim3(gx, batch), 256, 0, st>>>(src, ld, bstride, rows, cols, hi, lo, rows_p, cols_p);
        URSA_LAUNCH_CHECK("split_tf32_kernel");
        return URSA_OK;
    };
    if (int rc = split(x, in_dim, 0, (int)N, in_dim, x_hi, x_lo, (int)N, p.K1p, 1)) return rc;
    // hidden-activation padding columns (hidden..K2p) are never written by a tile: zero them once
    URSA_CUDA(cudaMemsetAsync(h1_hi, 0, 4 * (size_t)p.sc * p.h_plane * sizeof(float), st));

    for (int s0 = 0; s0 < S; s0 += p.sc) {
        const int nb = (S - s0 < p.sc) ? (S - s0) : p.sc;
        const float *bk = bank + (int64_t)s0 * ld_bank;
        if (int rc = split(bk + oW1, in_dim, ld_bank, hidden, in_dim, w1_hi, w1_lo, p.Np1, p.K1p, nb)) return rc;
        if (int rc = split(bk + oW2, hidden, ld_bank, hidden, hidden, w2_hi, w2_lo, p.Np1, p.K2p, nb)) return rc;
        if (int rc = split(bk + oW3, hidden, ld_bank, C, hidden, w3_hi, w3_lo, p.Np3, p.K2p, nb)) return rc;
        TcGemmArgs g;
        g.M = N; g.bias_stride = ld_bank;
        // layer 1: relu(x W1^T + b1) -> h1 (split)
        g.bias = bk + ob1; g.out_hi = h1_hi; g.out_lo = h1_lo; g.out_batch_stride = (int64_t)p.h_plane; g.ld_out = p.K2p;
        g.n_valid = hidden; g.relu = 1;
        if (int rc = launch_tc_gemm(x_hi, x_lo, N, 1, p.K1p, w1_hi, w1_lo, p.Np1, p.BN1, nb, g, st)) return rc;
        // layer 2: relu(h1 W2^T + b2) -> h2 (split)
        g.bias = bk + ob2; g.out_hi = h2_hi; g.out_lo = h2_lo;
        if (int rc = launch_tc_gemm(h1_hi, h1_lo, N, nb, p.K2p, w2_hi, w2_lo, p.Np1, p.BN1, nb, g, st)) return rc;
        // layer 3: logits = h2 W3^T + b3 (plain fp32)
        g.bias = bk + ob3; g.out_hi = lg; g.out_lo = nullptr; g.out_batch_stride = (int64_t)p.logit_plane; g.ld_out = C;
        g.n_valid = C; g.relu = 0;
        if (int rc = launch_tc_gemm(h2_hi, h2_lo, N, nb, p.K2p, w3_hi, w3_lo, p.Np3, p.BN3, nb, g, st)) return rc;
        if (int rc = ursa_bma_accumulate(lg, nb, N, C, (int64_t)p.logit_plane, proba_sum, entropy_sum, gamma, (void *)st))
            return rc;
        if (logits_out)
            URSA_CUDA(cudaMemcpy2DAsync(logits_out + (size_t)s0 * N * C, (size_t)N * C * sizeof(float), lg,
                                        p.logit_plane * sizeof(float), (size_t)N * C * sizeof(float), nb,
                                        cudaMemcpyDeviceToDevice, st));
    }
    return URSA_OK;
}

}  // namespace ursa


// ---- chain-batched likelihood gradient of the 3-layer MLP (HMC, BASELINE.json configs[3]) -------------------------------------
// What hamiltorch does per leapfrog step and chain -- one autograd pass of the module over the full batch
// (reference inference/hmc.py:71-75) -- for ALL chains as eight tcgen05 GEMMs on the 3xTF32 kernel above, with every operand
// produced in split (hi / lo) form by the epilogue of the GEMM before it:
//   forward   a1 = relu(X W1^T + b1), a2 = relu(a1 W2^T + b2), logits = a2 W3^T + b3       (a1, a2 also stored TRANSPOSED)
//   loss      ce[c] = sum_n -log softmax(logits)[y_n];  dlo = softmax - onehot              (point- and feature-major)
//   backward  da2 = (dlo W3) . [a2 > 0]   dW3 = dlo^T a2   dW2 = da2^T a1   da1 = (da2 W2) . [a1 > 0]   dW1 = da1^T X
// The weight-gradient GEMMs contract over the data points, i.e. over the ROWS of the activations: their operands are the
// transposed copies the producing epilogues wrote next to the point-major ones (a warp's lanes are consecutive points, so
// those stores are contiguous) -- no transpose kernels, no per-call operand split passes over the activations.
namespace ursa {

struct HmcGradPlan {
    int K1p, K2p, K3p, Kq, BNh, Nph, BNc, Npc, BNi, Npi;
    size_t x, xt, w1, w2, w3, w3t, act, actT, dlo, dloT, logits;     // floats per plane
    size_t total_bytes;
};

static HmcGradPlan hmc_grad_plan(int C, int64_t Npts, int in_dim, int hid, int ncls) {
    HmcGradPlan p;
    p.K1p = round_up_i(in_dim, TC_BK);
    p.K2p = round_up_i(hid, TC_BK);
    p.K3p = round_up_i(ncls, TC_BK);
    p.Kq = round_up_i((int)Npts, TC_BK);
    p.BNh = pick_bn(hid);   p.Nph = round_up_i(round_up_i(hid, 16), p.BNh);
    p.BNc = pick_bn(ncls);  p.Npc = round_up_i(round_up_i(ncls, 16), p.BNc);
    p.BNi = pick_bn(in_dim); p.Npi = round_up_i(round_up_i(in_dim, 16), p.BNi);
    p.x = (size_t)Npts * p.K1p;
    p.xt = (size_t)p.Npi * p.Kq;
    p.w1 = (size_t)C * p.Nph * p.K1p;
    p.w2 = (size_t)C * p.Nph * p.K2p;
    p.w3 = (size_t)C * p.Npc * p.K2p;
    p.w3t = (size_t)C * p.Nph * p.K3p;
    p.act = (size_t)C * Npts * p.K2p;
    p.actT = (size_t)C * p.Nph * p.Kq;
    p.dlo = (size_t)C * Npts * p.K3p;
    p.dloT = (size_t)C * p.Npc * p.Kq;
    p.logits = ((size_t)C * Npts * ncls + 3) & ~(size_t)3;
    // hi / lo pairs: X, XT, W1, W2, W2T, W3, W3T, a1, a2, da2, a1T, a2T, da2T, da1T, dlo, dloT ; plain: logits, zero bias
    const size_t fl = 2 * (p.x + p.xt + p.w1 + 2 * p.w2 + p.w3 + p.w3t + 3 * p.act + 4 * p.actT + p.dlo + p.dloT) + p.logits +
                      (size_t)(p.Nph > p.Npi ? p.Nph : p.Npi);
    p.total_bytes = fl * sizeof(float) + 4096;
    return p;
}

// theta rows -> split filter planes: W1 [h][in], W2 [h][h], W2^T, W3 [C][h], W3^T  (zero padded to the plane shapes)
__global__ void __launch_bounds__(256) hmc_mlp_prep_kernel(const float *__restrict__ theta, int64_t ld, int in_dim, int hid, int ncls,
                                                           HmcGradPlan p, float *w1h, float *w1l, float *w2h, float *w2l, float *w2th,
                                                           float *w2tl, float *w3h, float *w3l, float *w3th, float *w3tl) {
    const int c = blockIdx.y;
    const float *row = theta + (int64_t)c * ld;
    const int64_t oW1 = 0, oW2 = (int64_t)hid * in_dim + hid, oW3 = oW2 + (int64_t)hid * hid + hid;
    const int64_t n1 = (int64_t)p.Nph * p.K1p, n2 = (int64_t)p.Nph * p.K2p, n3 = (int64_t)p.Npc * p.K2p, n3t = (int64_t)p.Nph * p.K3p;
    const int64_t total = n1 + 2 * n2 + n3 + n3t;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        float x = 0.f, *dh, *dl;
        int64_t j = i;
        if (j < n1) {
            const int r = (int)(j / p.K1p), k = (int)(j % p.K1p);
            if (r < hid && k < in_dim) x = __ldg(row + oW1 + (int64_t)r * in_dim + k);
            dh = w1h + c * n1 + j; dl = w1l + c * n1 + j;
        } else if ((j -= n1) < n2) {
            const int r = (int)(j / p.K2p), k = (int)(j % p.K2p);
            if (r < hid && k < hid) x = __ldg(row + oW2 + (int64_t)r * hid + k);
            dh = w2h + c * n2 + j; dl = w2l + c * n2 + j;
        } else if ((j -= n2) < n2) {
            const int r = (int)(j / p.K2p), k = (int)(j % p.K2p);                      // W2^T[r = in-feature][k = out-feature]
            if (r < hid && k < hid) x = __ldg(row + oW2 + (int64_t)k * hid + r);
            dh = w2th + c * n2 + j; dl = w2tl + c * n2 + j;
        } else if ((j -= n2) < n3) {
            const int r = (int)(j / p.K2p), k = (int)(j % p.K2p);
            if (r < ncls && k < hid) x = __ldg(row + oW3 + (int64_t)r * hid + k);
            dh = w3h + c * n3 + j; dl = w3l + c * n3 + j;
        } else {
            j -= n3;
            const int r = (int)(j / p.K3p), k = (int)(j % p.K3p);                      // W3^T[r = hidden][k = class]
            if (r < hid && k < ncls) x = __ldg(row + oW3 + (int64_t)k * hid + r);
            dh = w3th + c * n3t + j; dl = w3tl + c * n3t + j;
        }
        const float h = rn_tf32(x);
        *dh = h;
        *dl = rn_tf32(x - h);
    }
}

// X [N][in] -> X planes [N][K1p] and X^T planes [Npi][Kq]
__global__ void __launch_bounds__(256) hmc_mlp_xprep_kernel(const float *__restrict__ x, int64_t Npts, int in_dim, HmcGradPlan p,
                                                            float *xh, float *xl, float *xth, float *xtl) {
    const int64_t n1 = (int64_t)Npts * p.K1p, n2 = (int64_t)p.Npi * p.Kq;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n1 + n2; i += (int64_t)gridDim.x * 256) {
        float v = 0.f, *dh, *dl;
        if (i < n1) {
            const int64_t r = i / p.K1p; const int k = (int)(i % p.K1p);
            if (k < in_dim) v = __ldg(x + r * in_dim + k);
            dh = xh + i; dl = xl + i;
        } else {
            const int64_t j = i - n1;
            const int r = (int)(j / p.Kq); const int64_t k = j % p.Kq;
            if (r < in_dim && k < Npts) v = __ldg(x + k * in_dim + r);
            dh = xth + j; dl = xtl + j;
        }
        const float h = rn_tf32(v);
        *dh = h;
        *dl = rn_tf32(v - h);
    }
}

// one CTA per chain: ce[c] (fixed reduction order) and dlo = softmax - onehot in both layouts (pads stay at the memset zeros)
__global__ void __launch_bounds__(256) hmc_mlp_loss_kernel(const float *__restrict__ logits, const int64_t *__restrict__ y, int64_t Npts,
                                                           int ncls, HmcGradPlan p, float *dh, float *dl, float *dth, float *dtl,
                                                           float *__restrict__ ce) {
    __shared__ float red[256];
    const int c = blockIdx.x;
    const float *lg = logits + (int64_t)c * Npts * ncls;
    float acc = 0.f;
    for (int64_t n = threadIdx.x; n < Npts; n += 256) {
        const float *l = lg + n * ncls;
        float m = -INFINITY;
        for (int k = 0; k < ncls; ++k) m = fmaxf(m, l[k]);
        float sum = 0.f;
        for (int k = 0; k < ncls; ++k) sum += expf(l[k] - m);
        const float lse = logf(sum);
        const int yy = (int)y[n];
        acc += -((l[yy] - m) - lse);
        for (int k = 0; k < ncls; ++k) {
            const float d = expf((l[k] - m) - lse) - (k == yy ? 1.f : 0.f);
            const float h = rn_tf32(d), lo = rn_tf32(d - h);
            const int64_t a = ((int64_t)c * Npts + n) * p.K3p + k, b = ((int64_t)c * p.Npc + k) * p.Kq + n;
            dh[a] = h; dl[a] = lo; dth[b] = h; dtl[b] = lo;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) ce[c] = red[0];
}

// bias gradients: db[c][f] = sum_n (hi + lo)[c][f][n] over the transposed planes; one warp per (chain, feature)
__global__ void __launch_bounds__(256) hmc_mlp_bias_kernel(const float *__restrict__ th, const float *__restrict__ tl, int rows_p, int Kq,
                                                           int n_feat, int C, float *__restrict__ grad, int64_t ld, int64_t off) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= C * n_feat) return;
    const int c = w / n_feat, f = w - c * n_feat;
    const float *ph = th + ((int64_t)c * rows_p + f) * Kq, *pl = tl + ((int64_t)c * rows_p + f) * Kq;
    float s = 0.f;
    for (int k = lane; k < Kq; k += 32) s += ph[k] + pl[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) grad[(int64_t)c * ld + off + f] = s;
}

}  // namespace ursa

extern "C" size_t ursa_hmc_mlp_grad_workspace(int C, int64_t Npts, int in_dim, int hidden, int ncls) {
    if (C < 1 || Npts < 1 || in_dim < 1 || hidden < 1 || ncls < 1 || in_dim % 4 != 0 || hidden % 4 != 0 || ncls > 128) return 0;
    return ursa::hmc_grad_plan(C, Npts, in_dim, hidden, ncls).total_bytes;
}

extern "C" int ursa_hmc_mlp_grad(const float *theta, int64_t ld, int C, const float *x, const int64_t *y, int64_t Npts, int in_dim,
                                 int hidden, int ncls, float *grad, float *ce, void *workspace, size_t workspace_bytes,
                                 void *stream) {
    using namespace ursa;
    URSA_REQUIRE(theta && x && y && grad && ce && workspace, "ursa_hmc_mlp_grad: null pointer");
    URSA_REQUIRE(ursa_hmc_mlp_grad_workspace(C, Npts, in_dim, hidden, ncls) != 0,
                 "ursa_hmc_mlp_grad: unsupported shape (in_dim %% 4, hidden %% 4, classes <= 128)");
    const int64_t D = (int64_t)hidden * in_dim + hidden + (int64_t)hidden * hidden + hidden + (int64_t)ncls * hidden + ncls;
    URSA_REQUIRE(ld >= D, "ursa_hmc_mlp_grad: ld (%lld) < D (%lld)", (long long)ld, (long long)D);
    const HmcGradPlan p = hmc_grad_plan(C, Npts, in_dim, hidden, ncls);
    URSA_REQUIRE(workspace_bytes >= p.total_bytes, "ursa_hmc_mlp_grad: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float *w = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    auto take = [&](size_t n) { float *r = w; w += n; return r; };
    float *xh = take(p.x), *xl = take(p.x), *xth = take(p.xt), *xtl = take(p.xt);
    float *w1h = take(p.w1), *w1l = take(p.w1), *w2h = take(p.w2), *w2l = take(p.w2), *w2th = take(p.w2), *w2tl = take(p.w2);
    float *w3h = take(p.w3), *w3l = take(p.w3), *w3th = take(p.w3t), *w3tl = take(p.w3t);
    float *a1h = take(p.act), *a1l = take(p.act), *a2h = take(p.act), *a2l = take(p.act), *d2h = take(p.act), *d2l = take(p.act);
    float *a1th = take(p.actT), *a1tl = take(p.actT), *a2th = take(p.actT), *a2tl = take(p.actT);
    float *d2th = take(p.actT), *d2tl = take(p.actT), *d1th = take(p.actT), *d1tl = take(p.actT);
    float *dlh = take(p.dlo), *dll = take(p.dlo), *dlth = take(p.dloT), *dltl = take(p.dloT);
    float *logits = take(p.logits);
    float *zero_bias = take((size_t)(p.Nph > p.Npi ? p.Nph : p.Npi));
    const int64_t oW1 = 0, ob1 = (int64_t)hidden * in_dim, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)hidden * hidden, oW3 = ob2 + hidden,
                  ob3 = oW3 + (int64_t)ncls * hidden;

    // pads: K2p - hidden columns of the point-major activations, the class pads of dlo / dloT, zero bias
    if (p.Nph < p.K2p)      // otherwise the N tiles of the producing GEMMs cover every column (zeros beyond `hidden`)
        URSA_CUDA(cudaMemsetAsync(a1h, 0, 6 * p.act * sizeof(float), st));
    URSA_CUDA(cudaMemsetAsync(dlh, 0, (2 * p.dlo + 2 * p.dloT) * sizeof(float), st));
    URSA_CUDA(cudaMemsetAsync(zero_bias, 0, (size_t)(p.Nph > p.Npi ? p.Nph : p.Npi) * sizeof(float), st));
    hmc_mlp_xprep_kernel<<<148 * 4, 256, 0, st>>>(x, Npts, in_dim, p, xh, xl, xth, xtl);
    URSA_LAUNCH_CHECK("hmc_mlp_xprep_kernel");
    hmc_mlp_prep_kernel<<<dim3(148, C), 256, 0, st>>>(theta, ld, in_dim, hidden, ncls, p, w1h, w1l, w2h, w2l, w2th, w2tl, w3h, w3l, w3th,
                                                      w3tl);
    URSA_LAUNCH_CHECK("hmc_mlp_prep_kernel");

    const int64_t act_bs = (int64_t)Npts * p.K2p, actT_bs = (int64_t)p.Nph * p.Kq;
    TcGemmArgs g;
    // ---- forward
    g = TcGemmArgs(); g.M = Npts; g.bias = theta + ob1; g.bias_stride = ld; g.relu = 1; g.n_valid = hidden;
    g.out_hi = a1h; g.out_lo = a1l; g.out_batch_stride = act_bs; g.ld_out = p.K2p;
    g.outT_hi = a1th; g.outT_lo = a1tl; g.outT_batch_stride = actT_bs; g.ld_t = p.Kq;
    if (int rc = launch_tc_gemm(xh, xl, Npts, 1, p.K1p, w1h, w1l, p.Nph, p.BNh, C, g, st)) return rc;
    g.bias = theta + ob2; g.out_hi = a2h; g.out_lo = a2l; g.outT_hi = a2th; g.outT_lo = a2tl;
    if (int rc = launch_tc_gemm(a1h, a1l, Npts, C, p.K2p, w2h, w2l, p.Nph, p.BNh, C, g, st)) return rc;
    g = TcGemmArgs(); g.M = Npts; g.bias = theta + ob3; g.bias_stride = ld; g.relu = 0; g.n_valid = ncls;
    g.out_hi = logits; g.out_lo = nullptr; g.out_batch_stride = (int64_t)Npts * ncls; g.ld_out = ncls;
    if (int rc = launch_tc_gemm(a2h, a2l, Npts, C, p.K2p, w3h, w3l, p.Npc, p.BNc, C, g, st)) return rc;
    // ---- loss
    hmc_mlp_loss_kernel<<<C, 256, 0, st>>>(logits, y, Npts, ncls, p, dlh, dll, dlth, dltl, ce);
    URSA_LAUNCH_CHECK("hmc_mlp_loss_kernel");
    // ---- backward
    // da2 = (dlo W3) . [a2 > 0]  -> point-major (A of the da1 GEMM) and transposed (A of the dW2 GEMM)
    g = TcGemmArgs(); g.M = Npts; g.bias = zero_bias; g.bias_stride = 0; g.relu = 0; g.n_valid = hidden;
    g.out_hi = d2h; g.out_lo = d2l; g.out_batch_stride = act_bs; g.ld_out = p.K2p;
    g.outT_hi = d2th; g.outT_lo = d2tl; g.outT_batch_stride = actT_bs; g.ld_t = p.Kq;
    g.mask = a2h; g.ld_mask = p.K2p; g.mask_batch_stride = act_bs;
    if (int rc = launch_tc_gemm(dlh, dll, Npts, C, p.K3p, w3th, w3tl, p.Nph, p.BNh, C, g, st)) return rc;
    // dW3^T[h][k] = sum_n a2T[h][n] dloT[k][n]  -> stored transposed = W3's own [k][h] layout
    g = TcGemmArgs(); g.M = hidden; g.bias = zero_bias; g.bias_stride = 0; g.relu = 0; g.n_valid = ncls;
    g.out_hi = nullptr; g.outT_hi = grad + oW3; g.outT_lo = nullptr; g.outT_batch_stride = ld; g.ld_t = hidden; g.t_valid = ncls;
    if (int rc = launch_tc_gemm(a2th, a2tl, p.Nph, C, p.Kq, dlth, dltl, p.Npc, p.BNc, C, g, st)) return rc;
    // dW2[o][i] = sum_n da2T[o][n] a1T[i][n]
    g = TcGemmArgs(); g.M = hidden; g.bias = zero_bias; g.bias_stride = 0; g.relu = 0; g.n_valid = hidden;
    g.out_hi = grad + oW2; g.out_lo = nullptr; g.out_batch_stride = ld; g.ld_out = hidden;
    if (int rc = launch_tc_gemm(d2th, d2tl, p.Nph, C, p.Kq, a1th, a1tl, p.Nph, p.BNh, C, g, st)) return rc;
    // da1 = (da2 W2) . [a1 > 0]  -> only its transposed form is needed
    g = TcGemmArgs(); g.M = Npts; g.bias = zero_bias; g.bias_stride = 0; g.relu = 0; g.n_valid = hidden;
    g.out_hi = nullptr; g.outT_hi = d1th; g.outT_lo = d1tl; g.outT_batch_stride = actT_bs; g.ld_t = p.Kq;
    g.mask = a1h; g.ld_mask = p.K2p; g.mask_batch_stride = act_bs;
    if (int rc = launch_tc_gemm(d2h, d2l, Npts, C, p.K2p, w2th, w2tl, p.Nph, p.BNh, C, g, st)) return rc;
    // dW1[o][k] = sum_n da1T[o][n] XT[k][n]   (XT shared by all chains)
    g = TcGemmArgs(); g.M = hidden; g.bias = zero_bias; g.bias_stride = 0; g.relu = 0; g.n_valid = in_dim; g.b_shared = 1;
    g.out_hi = grad + oW1; g.out_lo = nullptr; g.out_batch_stride = ld; g.ld_out = in_dim;
    if (int rc = launch_tc_gemm(d1th, d1tl, p.Nph, C, p.Kq, xth, xtl, p.Npi, p.BNi, C, g, st)) return rc;
    // bias gradients = row sums of the transposed planes
    const int gb1 = (C * hidden + 7) / 8, gb3 = (C * ncls + 7) / 8;
    hmc_mlp_bias_kernel<<<gb1, 256, 0, st>>>(d1th, d1tl, p.Nph, p.Kq, hidden, C, grad, ld, ob1);
    hmc_mlp_bias_kernel<<<gb1, 256, 0, st>>>(d2th, d2tl, p.Nph, p.Kq, hidden, C, grad, ld, ob2);
    hmc_mlp_bias_kernel<<<gb3, 256, 0, st>>>(dlth, dltl, p.Npc, p.Kq, ncls, C, grad, ld, ob3);
    URSA_LAUNCH_CHECK("hmc_mlp_bias_kernel");
    return URSA_OK;
}

// ---- generic entry: batched  out[b] = A[b or shared] B[b]^T (+ bias[b]) (ReLU)  on the same 3xTF32 kernel ----------------------
// (the chain-batched HMC likelihood gradient of the MLPs runs its nine GEMMs through this; SURVEY 8(d) cfg 4)
using namespace ursa;

static void gemm_plan(int batch, int64_t M, int N, int K, int a_batched, int &Kp, int &BN, int &Np, size_t &a_plane, size_t &b_plane,
                      size_t &total) {
    Kp = round_up_i(K, TC_BK);
    BN = pick_bn(N);
    Np = round_up_i(round_up_i(N, 16), BN);
    a_plane = (size_t)(a_batched ? batch : 1) * (size_t)M * Kp;
    b_plane = (size_t)batch * (size_t)Np * Kp;
    total = (2 * a_plane + 2 * b_plane + (size_t)Np) * sizeof(float) + 2048;
}

extern "C" size_t ursa_gemm_nt_3xtf32_workspace(int batch, int64_t M, int N, int K, int a_batched) {
    if (batch < 1 || M < 1 || N < 1 || K < 1) return 0;
    int Kp, BN, Np;
    size_t ap, bp, total;
    gemm_plan(batch, M, N, K, a_batched, Kp, BN, Np, ap, bp, total);
    return total;
}

extern "C" int ursa_gemm_nt_3xtf32(const float *A, int64_t lda, int64_t a_batch_stride, const float *B, int64_t ldb,
                                   int64_t b_batch_stride, const float *bias, int64_t bias_stride, int relu, float *out,
                                   int64_t ldo, int64_t out_batch_stride, int batch, int64_t M, int N, int K, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    URSA_REQUIRE(A && B && out && workspace, "ursa_gemm_nt_3xtf32: null pointer");
    URSA_REQUIRE(batch >= 1 && M >= 1 && N >= 1 && K >= 1 && lda >= K && ldb >= K && ldo >= N, "ursa_gemm_nt_3xtf32: bad shape");
    URSA_REQUIRE(M < (int64_t)1 << 31, "ursa_gemm_nt_3xtf32: M too large");
    const int a_batched = a_batch_stride != 0;
    int Kp, BN, Np;
    size_t ap, bp, total;
    gemm_plan(batch, M, N, K, a_batched, Kp, BN, Np, ap, bp, total);
    URSA_REQUIRE(workspace_bytes >= total, "ursa_gemm_nt_3xtf32: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float *ws = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    float *a_hi = ws, *a_lo = a_hi + ap, *b_hi = a_lo + ap, *b_lo = b_hi + bp, *zero_bias = b_lo + bp;
    auto split = [&](const float *src, int64_t ld, int64_t bstride, int rows, int cols, float *hi, float *lo, int rows_p, int cols_p,
                     int nb) -> int {
        const int64_t tot = (int64_t)rows_p * cols_p;
        int gx = (int)((tot + 255) / 256);
        if (gx > 148 * 8) gx = 148 * 8;
        split_tf32_kernel<<<dim3(gx, nb), 256, 0, st>>>(src, ld, bstride, rows, cols, hi, lo, rows_p, cols_p);
        URSA_LAUNCH_CHECK("split_tf32_kernel");
        return URSA_OK;
    };
    if (int rc = split(A, lda, a_batch_stride, (int)M, K, a_hi, a_lo, (int)M, Kp, a_batched ? batch : 1)) return rc;
    if (int rc = split(B, ldb, b_batch_stride, N, K, b_hi, b_lo, Np, Kp, batch)) return rc;
    TcGemmArgs g;
    g.M = M;
    if (bias) { g.bias = bias; g.bias_stride = bias_stride; }
    else {
        URSA_CUDA(cudaMemsetAsync(zero_bias, 0, (size_t)Np * sizeof(float), st));
        g.bias = zero_bias; g.bias_stride = 0;
    }
    g.out_hi = out; g.out_lo = nullptr; g.out_batch_stride = out_batch_stride; g.ld_out = (int)ldo;
    g.n_valid = N; g.relu = relu ? 1 : 0;
    return launch_tc_gemm(a_hi, a_lo, M, a_batched ? batch : 1, Kp, b_hi, b_lo, Np, BN, batch, g, st);
}

