// K3 (MLP), product path: sample-batched GEMM on 2xFP16-split operands, PERSISTENT CTAs, TMA-fed, accumulators in TMEM.
//
// Why a second tensor-core engine.  The 3xTF32 kernel (bma_mlp_tc.cu) streams fp32 hi / lo planes: 8 bytes per operand element
// and K = 8 per MMA.  At S = 100 x N = 10 000 its CTAs pull 80 GB through the L2 -> SM path per evaluation -- 7 TB/s of the
// ~12 TB/s the L2 slices deliver -- and every CTA ran prologue, main loop and epilogue back to back (one tile per CTA).
// Here an operand element is 4 bytes and one MMA covers K = 16:
//     x = hi + lo' 2^-11,   hi = rn_f16(x),   lo' = rn_f16((x - hi) 2^11)              (22 significant bits, as 3xTF32)
//     ACC += A_hi B_hi                      LO += A_hi B_lo' + A_lo' B_hi               (lo' lo' ~ 2^-22 dropped)
//     result = ACC + LO 2^-11                                                           (fp32, in the epilogue registers)
// as TWO tcgen05.mma.kind::f16 per K = 16: A_hi x [B_hi ; B_lo'] (N = 2 BN: the two B tiles are adjacent in shared memory, so
// one descriptor covers both and the accumulator columns come out as [ACC | LO]) and A_lo' x B_hi (N = BN) into the LO columns.
// Range: fp16 overflows beyond 65 504 -> inf -> NaN logits (the ReLU here propagates NaN): loud by construction; the class
// layer (tasks/_engine.py) redoes such an evaluation on the 3xTF32 engine, which has fp32's range.
//
// Kernel anatomy (320 threads, one CTA per SM, static tile schedule t = blockIdx.x + i gridDim.x over (sample, m, n), n fastest):
//   warp 0      TMA producer: per k-block (64 halves = one 128-byte swizzle row) four 3-D tiled loads into a 4-stage ring
//   warp 1      MMA issuer: 4 (K = 16) x 2 MMAs per stage; a chain covers F_SEG k-blocks (two-level accumulation: the tensor
//               core adds into TMEM with truncation, a bias that grows with the chain), then moves to the next of the rotating
//               TMEM accumulators -- across tile boundaries too, so the MMAs of tile i + 1 run under the epilogue of tile i
//   warps 2-9   epilogue (two warps per TMEM lane quarter, alternating 16-column chunks; with four warps layer 2, K = 448, was
//               epilogue bound): drain each finished segment (tcgen05.ld ACC and LO chunks -> acc + lo 2^-11 -> fp32 registers),
//               after a tile's last segment + bias -> ReLU -> split -> global (the next layer's TMA source), or fp32 logits
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace ursa {

constexpr int F_BM = 128, F_BK = 64, F_MAX_STAGES = 6, F_EPI_WARPS = 8, F_THREADS = 64 + 32 * F_EPI_WARPS;
constexpr int F_SEG = 4, F_MAX_TBUF = 4, F_EPI_CHUNKS = 4;             // BN <= 128 = 8 chunks of 16 columns, 4 per epilogue warp
constexpr uint32_t F_A_BYTES = F_BM * F_BK * 2;                          // 16 KB per A tile
constexpr float kF16LoScale = 2048.f, kF16LoUnscale = 1.f / 2048.f;    // 2^11

struct F16GemmArgs {
    const float *bias;            // per-sample bias vector: bias + s * bias_stride
    int64_t bias_stride;
    __half *out_hi, *out_lo;      // split outputs [S][M][ld_out] (next layer's operand planes), or null
    float *out_f32;               // plain fp32 output [S][M][ld_out] (the logits), or null
    int64_t out_batch_stride;
    int ld_out;
    int64_t M;
    int n_valid;                  // real output features
    int BN, k_blocks, a_batched, relu, stages;
    uint32_t tmem_cols;
    int seg, ntbuf;
    int m_blocks, n_blocks, tiles;
    // chain-batched MLP gradient (ursa_hmc_mlp_grad_f16): optional extras of the epilogue
    int b_shared = 0;                               // B shared by all samples (the data matrix of the dW1 GEMM)
    const __half *mask_hi = nullptr, *mask_lo = nullptr;   // [S][M][ld_mask] planes of a forward activation: v *= (act > 0)
    int ld_mask = 0;
    int64_t mask_batch_stride = 0;
    __half *outT_hi = nullptr, *outT_lo = nullptr;  // TRANSPOSED split copy [S][features][ld_t]: the operand of the weight-gradient
    float *outT_f32 = nullptr;                      // GEMMs (they contract over the rows of this one); or a plain fp32 transposed
    int ld_t = 0;                                   // store of features [0, t_valid).  Rows M .. ld_t of the split copy are written
    int64_t outT_batch_stride = 0;                  // as zeros.
    int t_valid = 0;
};

__device__ __forceinline__ uint32_t make_f16_idesc_mlp(int m, int n) {     // D = F32, A = B = F16, K-major both
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_mlp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(F_THREADS, 1)
mlp_f16_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                    const F16GemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[F_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F_MAX_STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[F_MAX_TBUF];
    __shared__ __align__(8) uint64_t tempty_bar[F_MAX_TBUF];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ntbuf = (uint32_t)a.ntbuf;
    const uint32_t b_bytes = (uint32_t)a.BN * F_BK * 2;
    const uint32_t stage_bytes = 2 * F_A_BYTES + 2 * b_bytes;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzle-128B tiles need 1024-byte alignment
    const int tiles_per_sample = a.m_blocks * a.n_blocks;

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < F_MAX_TBUF; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], F_EPI_WARPS);                          // one arrival per epilogue warp
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
            uint32_t st = 0, ph = 0;                                         // ring position / phase, carried across tiles
            for (int t = blockIdx.x; t < a.tiles; t += gridDim.x) {
                const int s = t / tiles_per_sample, r = t - s * tiles_per_sample;
                const int m_blk = r / a.n_blocks, n_blk = r - m_blk * a.n_blocks;
                const int a_b = a.a_batched ? s : 0;
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait_a(smem_u32(&empty_bar[st]), ph ^ 1u);          // slot free (first pass: passes at once)
                    const uint32_t fb = smem_u32(&full_bar[st]);
                    mbar_expect_tx_a(fb, stage_bytes);
                    const uint32_t base = smem_base + st * stage_bytes;
                    const int k0 = kb * F_BK;
                    tma_load_3d_a(base, &tm_a_hi, k0, m_blk * F_BM, a_b, fb);
                    tma_load_3d_a(base + F_A_BYTES, &tm_a_lo, k0, m_blk * F_BM, a_b, fb);
                    tma_load_3d_a(base + 2 * F_A_BYTES, &tm_b_hi, k0, n_blk * a.BN, a.b_shared ? 0 : s, fb);
                    tma_load_3d_a(base + 2 * F_A_BYTES + b_bytes, &tm_b_lo, k0, n_blk * a.BN, a.b_shared ? 0 : s, fb);
                    if (++st == (uint32_t)a.stages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (elect_one()) {
            const uint32_t idesc_cat = make_f16_idesc_mlp(F_BM, 2 * a.BN), idesc_lo = make_f16_idesc_mlp(F_BM, a.BN);
            uint32_t buf = 0, bph = 0, st = 0, ph = 0;
            for (int t = blockIdx.x; t < a.tiles; t += gridDim.x) {
                for (int kb0 = 0; kb0 < a.k_blocks; kb0 += a.seg) {
                    mbar_wait_a(smem_u32(&tempty_bar[buf]), bph ^ 1u);       // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * (uint32_t)(2 * a.BN);
                    const int kb1 = kb0 + a.seg < a.k_blocks ? kb0 + a.seg : a.k_blocks;
                    uint32_t acc = 0;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_a(smem_u32(&full_bar[st]), ph);            // TMA bytes have landed
                        tc_fence_after();
                        const uint32_t base = smem_base + st * stage_bytes;
                        const uint64_t d_ahi = make_kmajor_desc<128>(base), d_alo = make_kmajor_desc<128>(base + F_A_BYTES);
                        const uint64_t d_b = make_kmajor_desc<128>(base + 2 * F_A_BYTES);     // [B_hi ; B_lo'] rows
#pragma unroll
                        for (int k = 0; k < F_BK / 16; ++k) {
                            const uint64_t koff = (uint64_t)((k * 32) >> 4);  // advance 32 bytes inside the swizzle row
                            umma_f16_mlp(d_tmem, d_ahi + koff, d_b + koff, idesc_cat, acc);
                            acc = 1;
                            umma_f16_mlp(d_tmem + (uint32_t)a.BN, d_alo + koff, d_b + koff, idesc_lo, 1);
                        }
                        umma_commit(smem_u32(&empty_bar[st]));               // frees the smem slot when the MMAs retire
                        if (++st == (uint32_t)a.stages) { st = 0; ph ^= 1u; }
                    }
                    umma_commit(smem_u32(&tfull_bar[buf]));                  // segment complete -> epilogue drains it
                    if (++buf == ntbuf) { buf = 0; bph ^= 1u; }
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4; warps 2-5 take the even 16-column chunks, 6-9 the odd ones =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint32_t buf = 0, bph = 0;
        for (int t = blockIdx.x; t < a.tiles; t += gridDim.x) {
            const int s = t / tiles_per_sample, r = t - s * tiles_per_sample;
            const int m_blk = r / a.n_blocks, n_blk = r - m_blk * a.n_blocks;
            float accr[F_EPI_CHUNKS][16];
#pragma unroll
            for (int j = 0; j < F_EPI_CHUNKS; ++j)
#pragma unroll
                for (int e = 0; e < 16; ++e) accr[j][e] = 0.f;
            for (int kb0 = 0; kb0 < a.k_blocks; kb0 += a.seg) {
                mbar_wait_a(smem_u32(&tfull_bar[buf]), bph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(2 * a.BN);
#pragma unroll
                for (int j = 0; j < F_EPI_CHUNKS; ++j) {
                    const int c0 = (2 * j + half) * 16;
                    if (c0 < a.BN) {
                        uint32_t ra[16], rl[16];
                        tmem_ld16_nowait(taddr + (uint32_t)c0, ra);
                        tmem_ld16_nowait(taddr + (uint32_t)(a.BN + c0), rl);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            accr[j][e] += fmaf(__uint_as_float(rl[e]), kF16LoUnscale, __uint_as_float(ra[e]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                if (++buf == ntbuf) { buf = 0; bph ^= 1u; }
            }
            const int64_t row = (int64_t)m_blk * F_BM + q * 32 + lane;
            const bool live = row < a.M;
            const bool tlive = (a.outT_hi != nullptr || a.outT_f32 != nullptr) && row < a.ld_t;
            if (!live && !tlive) continue;
            const float *bias = a.bias ? a.bias + (int64_t)s * a.bias_stride : nullptr;
            __half *ohi = a.out_hi ? a.out_hi + (int64_t)s * a.out_batch_stride + row * a.ld_out : nullptr;
            __half *olo = a.out_hi ? a.out_lo + (int64_t)s * a.out_batch_stride + row * a.ld_out : nullptr;
            float *of = a.out_f32 ? a.out_f32 + (int64_t)s * a.out_batch_stride + row * a.ld_out : nullptr;
            const __half *mh = a.mask_hi ? a.mask_hi + (int64_t)s * a.mask_batch_stride + row * a.ld_mask : nullptr;
            const __half *ml = a.mask_hi ? a.mask_lo + (int64_t)s * a.mask_batch_stride + row * a.ld_mask : nullptr;
            __half *thi = a.outT_hi ? a.outT_hi + (int64_t)s * a.outT_batch_stride + row : nullptr;
            __half *tlo = a.outT_hi ? a.outT_lo + (int64_t)s * a.outT_batch_stride + row : nullptr;
            float *tf = a.outT_f32 ? a.outT_f32 + (int64_t)s * a.outT_batch_stride + row : nullptr;
#pragma unroll
            for (int j = 0; j < F_EPI_CHUNKS; ++j) {
                const int c0 = (2 * j + half) * 16;
                if (c0 >= a.BN) continue;
                const int col0 = n_blk * a.BN + c0;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int col = col0 + i;
                    float x = accr[j][i] + ((bias != nullptr && col < a.n_valid) ? __ldg(bias + col) : 0.f);
                    if (a.relu) x = relu_nan(x);
                    v[i] = live ? x : 0.f;
                }
                if (mh != nullptr && live) {
                    // ReLU derivative of the forward activation: act > 0 <=> one of its split parts is positive (act >= 0)
                    const uint4 m0 = __ldg(reinterpret_cast<const uint4 *>(mh + col0)), m1 = __ldg(reinterpret_cast<const uint4 *>(mh + col0) + 1);
                    const uint4 l0 = __ldg(reinterpret_cast<const uint4 *>(ml + col0)), l1 = __ldg(reinterpret_cast<const uint4 *>(ml + col0) + 1);
                    const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                    const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 am = __half22float2(*reinterpret_cast<const __half2 *>(&mw[i]));
                        const float2 al = __half22float2(*reinterpret_cast<const __half2 *>(&lw[i]));
                        if (!(am.x > 0.f || al.x > 0.f)) v[2 * i] = 0.f;
                        if (!(am.y > 0.f || al.y > 0.f)) v[2 * i + 1] = 0.f;
                    }
                }
                uint32_t h[8], l[8];
                if (ohi != nullptr || thi != nullptr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn((v[2 * i] - hf.x) * kF16LoScale, (v[2 * i + 1] - hf.y) * kF16LoScale);
                        h[i] = *reinterpret_cast<const uint32_t *>(&hh);
                        l[i] = *reinterpret_cast<const uint32_t *>(&ll);
                    }
                }
                if (live && ohi != nullptr) {
                    if (col0 + 16 <= a.ld_out) {
                        uint4 *ph4 = reinterpret_cast<uint4 *>(ohi + col0), *pl4 = reinterpret_cast<uint4 *>(olo + col0);
                        ph4[0] = make_uint4(h[0], h[1], h[2], h[3]);
                        ph4[1] = make_uint4(h[4], h[5], h[6], h[7]);
                        pl4[0] = make_uint4(l[0], l[1], l[2], l[3]);
                        pl4[1] = make_uint4(l[4], l[5], l[6], l[7]);
                    }
                } else if (live && of != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (col0 + i < a.n_valid) of[col0 + i] = v[i];
                }
                if (tlive) {
                    // transposed copy: lanes are consecutive rows, so every store of a warp is one contiguous segment
                    if (thi != nullptr) {
                        // a lane pair owns rows (r, r + 1) x columns (2 i, 2 i + 1): the even lane stores column 2 i of both rows,
                        // the odd lane column 2 i + 1 -- 4-byte stores, 128 contiguous bytes per warp and column pair (tlive is
                        // warp-uniform: a warp's 32 rows start on a multiple of 32 and ld_t is a multiple of 64)
                        const bool odd = lane & 1;
                        __half *th2 = thi - (odd ? 1 : 0), *tl2 = tlo - (odd ? 1 : 0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t ph = __shfl_xor_sync(0xffffffffu, h[i], 1), pl = __shfl_xor_sync(0xffffffffu, l[i], 1);
                            const uint32_t oh = odd ? __byte_perm(ph, h[i], 0x7632) : __byte_perm(h[i], ph, 0x5410);
                            const uint32_t ol = odd ? __byte_perm(pl, l[i], 0x7632) : __byte_perm(l[i], pl, 0x5410);
                            const int64_t co = (int64_t)(col0 + 2 * i + (odd ? 1 : 0)) * a.ld_t;
                            *reinterpret_cast<uint32_t *>(th2 + co) = oh;
                            *reinterpret_cast<uint32_t *>(tl2 + co) = ol;
                        }
                    } else if (live) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (col0 + i < a.t_valid) tf[(int64_t)(col0 + i) * a.ld_t] = v[i];
                    }
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

// zero columns [c0, c0 + nq * 8) of a half plane [rows][pitch] (pad columns no tile writes): one 16-byte store per thread
__global__ void __launch_bounds__(256) zero_cols_h_kernel(__half *__restrict__ p, int64_t rows, int pitch, int c0, int nq) {
    const int64_t total = rows * nq;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / nq;
        const int q = (int)(i - r * nq);
        *reinterpret_cast<uint4 *>(p + r * pitch + c0 + q * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
}
static int zero_cols_h(__half *p, int64_t rows, int pitch, int c0, int c1, cudaStream_t st) {
    if (c1 <= c0 || rows <= 0) return URSA_OK;
    if ((c0 & 7) || ((c1 - c0) & 7) || (pitch & 7)) {                    // not 16-byte granular: whole plane
        URSA_CUDA(cudaMemsetAsync(p, 0, (size_t)rows * pitch * sizeof(__half), st));
        return URSA_OK;
    }
    const int nq = (c1 - c0) / 8;
    int64_t blocks = (rows * nq + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    zero_cols_h_kernel<<<(int)blocks, 256, 0, st>>>(p, rows, pitch, c0, nq);
    URSA_LAUNCH_CHECK("zero_cols_h_kernel");
    return URSA_OK;
}

// fp32 [rows, cols] (ld, per-batch stride) -> zero-padded hi / lo' half planes [batch][rows_p][cols_p]
__global__ void __launch_bounds__(256) split_f16_kernel(const float *__restrict__ src, int64_t ld_src, int64_t src_batch_stride,
                                                        int rows, int cols, __half *__restrict__ hi, __half *__restrict__ lo,
                                                        int rows_p, int cols_p) {
    const int b = blockIdx.y;
    const int64_t total = (int64_t)rows_p * cols_p;
    const float *sb = src + (int64_t)b * src_batch_stride;
    __half *hb = hi + (int64_t)b * total, *lb = lo + (int64_t)b * total;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c = (int)(i % cols_p);
        const int64_t r = i / cols_p;
        float x = 0.f;
        if (r < rows && c < cols) x = __ldg(sb + r * ld_src + c);
        const __half h = __float2half_rn(x);
        hb[i] = h;
        lb[i] = __float2half_rn((x - __half2float(h)) * kF16LoScale);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
// 3-D half tensor [batch][rows][cols_p] (cols innermost), box = [1][box_rows][64], 128-byte swizzle
static int make_tmap_h(CUtensorMap *tm, const __half *base, int64_t cols_p, int64_t rows, int64_t batch, int box_rows) {
    const uint64_t dims[3] = {(uint64_t)cols_p, (uint64_t)rows, (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)cols_p * 2, (uint64_t)cols_p * 2 * (uint64_t)rows};
    const uint32_t box[3] = {F_BK, (uint32_t)box_rows, 1};
    return make_tensor_map_t(tm, base, 3, dims, strides, box, 128, 1);
}

static int round_up_f(int v, int m) { return (v + m - 1) / m * m; }

// N tile: multiple of 16, <= 128 (the concatenated [B_hi ; B_lo'] operand has N = 2 BN <= 256), minimal padding (ties -> larger)
static int pick_bn_f16(int n_out) {
    const int nr = round_up_f(n_out, 16);
    if (nr <= 128) return nr;
    int best = 128, best_waste = round_up_f(nr, 128) - nr;
    for (int bn = 112; bn >= 64; bn -= 16) {
        const int w = round_up_f(nr, bn) - nr;
        if (w < best_waste) { best = bn; best_waste = w; }
    }
    return best;
}

struct F16Plan {
    int K1p, K2p, BN1, BN3, Np1, Np3, sc;
    size_t x_plane, w1_plane, w2_plane, w3_plane, h_plane, logit_plane;   // elements per (sample) plane
    size_t total_bytes;
};

static F16Plan f16_plan(int S, int64_t N, int in_dim, int hidden, int C) {
    F16Plan p;
    p.K1p = round_up_f(in_dim, F_BK);
    p.K2p = round_up_f(hidden, F_BK);
    p.BN1 = pick_bn_f16(hidden);
    p.BN3 = pick_bn_f16(C);
    p.Np1 = round_up_f(round_up_f(hidden, 16), p.BN1);
    p.Np3 = round_up_f(round_up_f(C, 16), p.BN3);
    p.x_plane = ((size_t)N * p.K1p + 511) & ~(size_t)511;
    p.w1_plane = (size_t)p.Np1 * p.K1p;
    p.w2_plane = (size_t)p.Np1 * p.K2p;
    p.w3_plane = (size_t)p.Np3 * p.K2p;
    p.h_plane = (size_t)N * p.K2p;
    p.logit_plane = ((size_t)N * C + 3) & ~(size_t)3;
    const size_t per_sample_bytes = 2 * 2 * (p.w1_plane + p.w2_plane + p.w3_plane) + 2 * 4 * p.h_plane + 4 * p.logit_plane;
    size_t sc = (size_t)(768ull << 20) / (per_sample_bytes + 1);
    if (sc < 1) sc = 1;
    if (sc > (size_t)S) sc = S;
    if (sc > 32) sc = 32;
    p.sc = (int)sc;
    p.total_bytes = 2 * 2 * p.x_plane + (size_t)p.sc * per_sample_bytes + 4096;
    return p;
}

static int launch_f16_gemm(const __half *a_hi, const __half *a_lo, int64_t a_rows, int64_t a_batch, int Kp, const __half *b_hi,
                           const __half *b_lo, int Np, int BN, int batch, F16GemmArgs g, cudaStream_t st) {
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (int rc = make_tmap_h(&ta_hi, a_hi, Kp, a_rows, a_batch, F_BM)) return rc;
    if (int rc = make_tmap_h(&ta_lo, a_lo, Kp, a_rows, a_batch, F_BM)) return rc;
    if (int rc = make_tmap_h(&tb_hi, b_hi, Kp, Np, g.b_shared ? 1 : batch, BN)) return rc;
    if (int rc = make_tmap_h(&tb_lo, b_lo, Kp, Np, g.b_shared ? 1 : batch, BN)) return rc;
    g.BN = BN;
    g.k_blocks = Kp / F_BK;
    g.a_batched = a_batch > 1 ? 1 : 0;
    g.seg = F_SEG;
    if (const char *e = getenv("URSA_MLP_SEG")) { const int v = atoi(e); if (v >= 1) g.seg = v; }   // accuracy / speed experiments
    g.ntbuf = 512 / (2 * BN) < F_MAX_TBUF ? 512 / (2 * BN) : F_MAX_TBUF;
    uint32_t cols = 32;
    while (cols < (uint32_t)(g.ntbuf * 2 * BN)) cols <<= 1;
    g.tmem_cols = cols;
    const size_t stage_bytes = 2 * F_A_BYTES + 2 * (size_t)BN * F_BK * 2;
    int stages = (int)((size_t)(225 << 10) / stage_bytes);
    if (stages > F_MAX_STAGES) stages = F_MAX_STAGES;
    URSA_REQUIRE(stages >= 2, "mlp_f16_gemm: tile does not fit in shared memory");
    g.stages = stages;
    g.m_blocks = (int)((g.M + F_BM - 1) / F_BM);
    g.n_blocks = Np / BN;
    g.tiles = batch * g.m_blocks * g.n_blocks;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    URSA_CUDA(cudaFuncSetAttribute(mlp_f16_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
    mlp_f16_gemm_kernel<<<grid, F_THREADS, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, g);
    URSA_LAUNCH_CHECK("mlp_f16_gemm_kernel");
    return URSA_OK;
}

size_t mlp_workspace_f16(int S, int64_t N, int in_dim, int hidden, int C) {
    if (in_dim % 4 != 0 || hidden % 4 != 0 || C > 128) return 0;
    return f16_plan(S, N, in_dim, hidden, C).total_bytes;
}

int mlp_forward_f16(const float *bank, int64_t ld_bank, int S, const float *x, int64_t N, int in_dim, int hidden, int C,
                    float *proba_sum, float *entropy_sum, float *logits_out, double gamma, void *workspace,
                    size_t workspace_bytes, cudaStream_t st) {
    if (mlp_workspace_f16(S, N, in_dim, hidden, C) == 0) {
        set_error("URSA_ALGO_TCGEN05_F16 does not cover this MLP shape; use URSA_ALGO_TCGEN05 or URSA_ALGO_FFMA");
        return URSA_ERR_UNSUPPORTED;
    }
    const F16Plan p = f16_plan(S, N, in_dim, hidden, C);
    URSA_REQUIRE(workspace_bytes >= p.total_bytes, "ursa_bma_mlp_forward: workspace too small");
    __half *ws = reinterpret_cast<__half *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __half *x_hi = ws, *x_lo = x_hi + p.x_plane;
    __half *w1_hi = x_lo + p.x_plane, *w1_lo = w1_hi + p.sc * p.w1_plane;
    __half *w2_hi = w1_lo + p.sc * p.w1_plane, *w2_lo = w2_hi + p.sc * p.w2_plane;
    __half *w3_hi = w2_lo + p.sc * p.w2_plane, *w3_lo = w3_hi + p.sc * p.w3_plane;
    __half *h1_hi = w3_lo + p.sc * p.w3_plane, *h1_lo = h1_hi + p.sc * p.h_plane;
    __half *h2_hi = h1_lo + p.sc * p.h_plane, *h2_lo = h2_hi + p.sc * p.h_plane;
    float *lg = reinterpret_cast<float *>(h2_lo + p.sc * p.h_plane);

    const int64_t oW1 = 0, ob1 = oW1 + (int64_t)hidden * in_dim, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)hidden * hidden,
                  oW3 = ob2 + hidden, ob3 = oW3 + (int64_t)C * hidden;
    auto split = [&](const float *src, int64_t ld, int64_t bstride, int rows, int cols, __half *hi, __half *lo, int rows_p,
                     int cols_p, int batch) -> int {
        const int64_t total = (int64_t)rows_p * cols_p;
        int gx = (int)((total + 255) / 256);
        if (gx > 148 * 8) gx = 148 * 8;
        split_f16_kernel<<<dim3(gx, batch), 256, 0, st>>>(src, ld, bstride, rows, cols, hi, lo, rows_p, cols_p);
        URSA_LAUNCH_CHECK("split_f16_kernel");
        return URSA_OK;
    };
    if (int rc = split(x, in_dim, 0, (int)N, in_dim, x_hi, x_lo, (int)N, p.K1p, 1)) return rc;
    // hidden-activation padding columns (Np1..K2p) are never written by a tile: zero those columns of the four planes (a strided
    // memset: 1/9 of the planes at hidden = 400)
    if (int rc = zero_cols_h(h1_hi, 4 * (int64_t)p.sc * N, p.K2p, p.Np1, p.K2p, st)) return rc;

    for (int s0 = 0; s0 < S; s0 += p.sc) {
        const int nb = (S - s0 < p.sc) ? (S - s0) : p.sc;
        const float *bk = bank + (int64_t)s0 * ld_bank;
        if (int rc = split(bk + oW1, in_dim, ld_bank, hidden, in_dim, w1_hi, w1_lo, p.Np1, p.K1p, nb)) return rc;
        if (int rc = split(bk + oW2, hidden, ld_bank, hidden, hidden, w2_hi, w2_lo, p.Np1, p.K2p, nb)) return rc;
        if (int rc = split(bk + oW3, hidden, ld_bank, C, hidden, w3_hi, w3_lo, p.Np3, p.K2p, nb)) return rc;
        F16GemmArgs g;
        g.M = N; g.bias_stride = ld_bank;
        // layer 1: relu(x W1^T + b1) -> h1 (split)
        g.bias = bk + ob1; g.out_hi = h1_hi; g.out_lo = h1_lo; g.out_f32 = nullptr; g.out_batch_stride = (int64_t)p.h_plane;
        g.ld_out = p.K2p; g.n_valid = hidden; g.relu = 1;
        if (int rc = launch_f16_gemm(x_hi, x_lo, N, 1, p.K1p, w1_hi, w1_lo, p.Np1, p.BN1, nb, g, st)) return rc;
        // layer 2: relu(h1 W2^T + b2) -> h2 (split)
        g.bias = bk + ob2; g.out_hi = h2_hi; g.out_lo = h2_lo;
        if (int rc = launch_f16_gemm(h1_hi, h1_lo, N, nb, p.K2p, w2_hi, w2_lo, p.Np1, p.BN1, nb, g, st)) return rc;
        // layer 3: logits = h2 W3^T + b3 (plain fp32)
        g.bias = bk + ob3; g.out_hi = nullptr; g.out_lo = nullptr; g.out_f32 = lg; g.out_batch_stride = (int64_t)p.logit_plane;
        g.ld_out = C; g.n_valid = C; g.relu = 0;
        if (int rc = launch_f16_gemm(h2_hi, h2_lo, N, nb, p.K2p, w3_hi, w3_lo, p.Np3, p.BN3, nb, g, st)) return rc;
        if (int rc = ursa_bma_accumulate(lg, nb, N, C, (int64_t)p.logit_plane, proba_sum, entropy_sum, gamma, (void *)st))
            return rc;
        if (logits_out)
            URSA_CUDA(cudaMemcpy2DAsync(logits_out + (size_t)s0 * N * C, (size_t)N * C * sizeof(float), lg,
                                        p.logit_plane * sizeof(float), (size_t)N * C * sizeof(float), nb,
                                        cudaMemcpyDeviceToDevice, st));
    }
    return URSA_OK;
}


// ---- chain-batched likelihood gradient of the 3-layer MLP on the FP16-split kernel (HMC, BASELINE.json configs[3]) -----------
// Same dataflow as ursa_hmc_mlp_grad (bma_mlp_tc.cu): forward, loss and backward of ALL chains as eight GEMMs whose operands are
// produced in split form by the epilogue of the GEMM before them, activations additionally stored TRANSPOSED because the
// weight-gradient GEMMs contract over the data points.  Here the planes are halves (x = hi + lo' 2^-11) and the GEMMs run on the
// persistent kernel above: the 3xTF32 version spent most of a gradient in the epilogues of one-tile CTAs (K = 224 / 32: a few
// MMAs per tile, then 4 planes of stores with nothing to overlap them).
struct HmcF16Plan {
    int K1p, K2p, K3p, Kq, BNh, Nph, BNc, Npc, BNi, Npi;
    size_t x, xt, w1, w2, w3, w3t, act, actT, dlo, dloT, logits;     // elements per plane
    size_t total_bytes;
};

static HmcF16Plan hmc_f16_plan(int C, int64_t Npts, int in_dim, int hid, int ncls) {
    HmcF16Plan p;
    p.K1p = round_up_f(in_dim, F_BK);
    p.K2p = round_up_f(hid, F_BK);
    p.K3p = round_up_f(ncls, F_BK);
    p.Kq = round_up_f((int)Npts, F_BK);
    p.BNh = pick_bn_f16(hid);    p.Nph = round_up_f(round_up_f(hid, 16), p.BNh);
    p.BNc = pick_bn_f16(ncls);   p.Npc = round_up_f(round_up_f(ncls, 16), p.BNc);
    p.BNi = pick_bn_f16(in_dim); p.Npi = round_up_f(round_up_f(in_dim, 16), p.BNi);
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
    // hi / lo' pairs: X, XT, W1, W2, W2T, W3, W3T, a1, a2, da2, a1T, a2T, da2T, da1T, dlo, dloT (halves) ; plain fp32: logits
    const size_t halves = 2 * (p.x + p.xt + p.w1 + 2 * p.w2 + p.w3 + p.w3t + 3 * p.act + 4 * p.actT + p.dlo + p.dloT);
    p.total_bytes = halves * sizeof(__half) + p.logits * sizeof(float) + 4096 + 40 * 1024;   // planes are padded to 1 KB
    return p;
}

__device__ __forceinline__ void put_split_h(__half *h, __half *l, float x) {
    const __half hh = __float2half_rn(x);
    *h = hh;
    *l = __float2half_rn((x - __half2float(hh)) * kF16LoScale);
}

// theta rows -> split filter planes: W1 [h][in], W2 [h][h], W2^T, W3 [C][h], W3^T  (zero padded to the plane shapes)
__global__ void __launch_bounds__(256) hmc_f16_prep_kernel(const float *__restrict__ theta, int64_t ld, int in_dim, int hid, int ncls,
                                                           HmcF16Plan p, __half *w1h, __half *w1l, __half *w2h, __half *w2l,
                                                           __half *w2th, __half *w2tl, __half *w3h, __half *w3l, __half *w3th,
                                                           __half *w3tl) {
    const int c = blockIdx.y;
    const float *row = theta + (int64_t)c * ld;
    const int64_t oW1 = 0, oW2 = (int64_t)hid * in_dim + hid, oW3 = oW2 + (int64_t)hid * hid + hid;
    const int64_t n1 = (int64_t)p.Nph * p.K1p, n2 = (int64_t)p.Nph * p.K2p, n3 = (int64_t)p.Npc * p.K2p, n3t = (int64_t)p.Nph * p.K3p;
    const int64_t total = n1 + 2 * n2 + n3 + n3t;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        float x = 0.f;
        __half *dh, *dl;
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
        put_split_h(dh, dl, x);
    }
}

// X [N][in] -> X planes [N][K1p] and X^T planes [Npi][Kq]
__global__ void __launch_bounds__(256) hmc_f16_xprep_kernel(const float *__restrict__ x, int64_t Npts, int in_dim, HmcF16Plan p,
                                                            __half *xh, __half *xl, __half *xth, __half *xtl) {
    const int64_t n1 = (int64_t)Npts * p.K1p, n2 = (int64_t)p.Npi * p.Kq;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n1 + n2; i += (int64_t)gridDim.x * 256) {
        float v = 0.f;
        __half *dh, *dl;
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
        put_split_h(dh, dl, v);
    }
}

// one CTA per chain: ce[c] (fixed reduction order) and dlo = softmax - onehot in both layouts (pads stay at the memset zeros)
__global__ void __launch_bounds__(256) hmc_f16_loss_kernel(const float *__restrict__ logits, const int64_t *__restrict__ y, int64_t Npts,
                                                           int ncls, HmcF16Plan p, __half *dh, __half *dl, __half *dth, __half *dtl,
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
            const int64_t a = ((int64_t)c * Npts + n) * p.K3p + k, b = ((int64_t)c * p.Npc + k) * p.Kq + n;
            put_split_h(dh + a, dl + a, d);
            dth[b] = dh[a];
            dtl[b] = dl[a];
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

// bias gradients: db[c][f] = sum_n (hi + lo' 2^-11)[c][f][n] over the transposed planes; one warp per (chain, feature)
__global__ void __launch_bounds__(256) hmc_f16_bias_kernel(const __half *__restrict__ th, const __half *__restrict__ tl, int rows_p, int Kq,
                                                           int n_feat, int C, float *__restrict__ grad, int64_t ld, int64_t off) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= C * n_feat) return;
    const int c = w / n_feat, f = w - c * n_feat;
    const __half *ph = th + ((int64_t)c * rows_p + f) * Kq, *pl = tl + ((int64_t)c * rows_p + f) * Kq;
    float sh = 0.f, sl = 0.f;
    for (int k = lane; k < Kq; k += 32) { sh += __half2float(ph[k]); sl += __half2float(pl[k]); }
    float s = fmaf(sl, kF16LoUnscale, sh);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) grad[(int64_t)c * ld + off + f] = s;
}

}  // namespace ursa

extern "C" size_t ursa_hmc_mlp_grad_f16_workspace(int C, int64_t Npts, int in_dim, int hidden, int ncls) {
    if (C < 1 || Npts < 1 || in_dim < 1 || hidden < 1 || ncls < 1 || in_dim % 4 != 0 || hidden % 4 != 0 || ncls > 128) return 0;
    return ursa::hmc_f16_plan(C, Npts, in_dim, hidden, ncls).total_bytes;
}

extern "C" int ursa_hmc_mlp_grad_f16(const float *theta, int64_t ld, int C, const float *x, const int64_t *y, int64_t Npts,
                                     int in_dim, int hidden, int ncls, float *grad, float *ce, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    using namespace ursa;
    URSA_REQUIRE(theta && x && y && grad && ce && workspace, "ursa_hmc_mlp_grad_f16: null pointer");
    URSA_REQUIRE(ursa_hmc_mlp_grad_f16_workspace(C, Npts, in_dim, hidden, ncls) != 0,
                 "ursa_hmc_mlp_grad_f16: unsupported shape (in_dim %% 4, hidden %% 4, classes <= 128)");
    const int64_t D = (int64_t)hidden * in_dim + hidden + (int64_t)hidden * hidden + hidden + (int64_t)ncls * hidden + ncls;
    URSA_REQUIRE(ld >= D, "ursa_hmc_mlp_grad_f16: ld (%lld) < D (%lld)", (long long)ld, (long long)D);
    const HmcF16Plan p = hmc_f16_plan(C, Npts, in_dim, hidden, ncls);
    URSA_REQUIRE(workspace_bytes >= p.total_bytes, "ursa_hmc_mlp_grad_f16: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    __half *w = reinterpret_cast<__half *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    auto take = [&](size_t n) { __half *r = w; w += (n + 511) & ~(size_t)511; return r; };     // planes stay 1 KB aligned
    __half *xh = take(p.x), *xl = take(p.x), *xth = take(p.xt), *xtl = take(p.xt);
    __half *w1h = take(p.w1), *w1l = take(p.w1), *w2h = take(p.w2), *w2l = take(p.w2), *w2th = take(p.w2), *w2tl = take(p.w2);
    __half *w3h = take(p.w3), *w3l = take(p.w3), *w3th = take(p.w3t), *w3tl = take(p.w3t);
    __half *a1h = take(p.act), *a1l = take(p.act), *a2h = take(p.act), *a2l = take(p.act), *d2h = take(p.act), *d2l = take(p.act);
    __half *a1th = take(p.actT), *a1tl = take(p.actT), *a2th = take(p.actT), *a2tl = take(p.actT);
    __half *d2th = take(p.actT), *d2tl = take(p.actT), *d1th = take(p.actT), *d1tl = take(p.actT);
    __half *dlh = take(p.dlo), *dll = take(p.dlo), *dlth = take(p.dloT), *dltl = take(p.dloT);
    float *logits = reinterpret_cast<float *>(w);
    const int64_t oW1 = 0, ob1 = (int64_t)hidden * in_dim, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)hidden * hidden, oW3 = ob2 + hidden,
                  ob3 = oW3 + (int64_t)ncls * hidden;

    // pads: K2p - Nph columns of the point-major activations (the N tiles cover the rest), the class pads of dlo / dloT
    for (__half *pl : {a1h, a1l, a2h, a2l, d2h, d2l})
        if (int rc = zero_cols_h(pl, (int64_t)C * Npts, p.K2p, p.Nph, p.K2p, st)) return rc;
    URSA_CUDA(cudaMemsetAsync(dlh, 0, (size_t)((char *)logits - (char *)dlh), st));
    hmc_f16_xprep_kernel<<<148 * 4, 256, 0, st>>>(x, Npts, in_dim, p, xh, xl, xth, xtl);
    URSA_LAUNCH_CHECK("hmc_f16_xprep_kernel");
    hmc_f16_prep_kernel<<<dim3(148, C), 256, 0, st>>>(theta, ld, in_dim, hidden, ncls, p, w1h, w1l, w2h, w2l, w2th, w2tl, w3h, w3l, w3th,
                                                      w3tl);
    URSA_LAUNCH_CHECK("hmc_f16_prep_kernel");

    const int64_t act_bs = (int64_t)Npts * p.K2p, actT_bs = (int64_t)p.Nph * p.Kq;
    F16GemmArgs g;
    auto fresh = [&]() {
        F16GemmArgs z;
        z.bias = nullptr; z.bias_stride = 0; z.out_hi = z.out_lo = nullptr; z.out_f32 = nullptr; z.out_batch_stride = 0; z.ld_out = 0;
        z.M = 0; z.n_valid = 0; z.relu = 0;
        return z;
    };
    // ---- forward
    g = fresh(); g.M = Npts; g.bias = theta + ob1; g.bias_stride = ld; g.relu = 1; g.n_valid = hidden;
    g.out_hi = a1h; g.out_lo = a1l; g.out_batch_stride = act_bs; g.ld_out = p.K2p;
    g.outT_hi = a1th; g.outT_lo = a1tl; g.outT_batch_stride = actT_bs; g.ld_t = p.Kq;
    if (int rc = launch_f16_gemm(xh, xl, Npts, 1, p.K1p, w1h, w1l, p.Nph, p.BNh, C, g, st)) return rc;
    g.bias = theta + ob2; g.out_hi = a2h; g.out_lo = a2l; g.outT_hi = a2th; g.outT_lo = a2tl;
    if (int rc = launch_f16_gemm(a1h, a1l, Npts, C, p.K2p, w2h, w2l, p.Nph, p.BNh, C, g, st)) return rc;
    g = fresh(); g.M = Npts; g.bias = theta + ob3; g.bias_stride = ld; g.n_valid = ncls;
    g.out_f32 = logits; g.out_batch_stride = (int64_t)Npts * ncls; g.ld_out = ncls;
    if (int rc = launch_f16_gemm(a2h, a2l, Npts, C, p.K2p, w3h, w3l, p.Npc, p.BNc, C, g, st)) return rc;
    // ---- loss
    hmc_f16_loss_kernel<<<C, 256, 0, st>>>(logits, y, Npts, ncls, p, dlh, dll, dlth, dltl, ce);
    URSA_LAUNCH_CHECK("hmc_f16_loss_kernel");
    // ---- backward
    // da2 = (dlo W3) . [a2 > 0]  -> point-major (A of the da1 GEMM) and transposed (A of the dW2 GEMM)
    g = fresh(); g.M = Npts; g.n_valid = hidden;
    g.out_hi = d2h; g.out_lo = d2l; g.out_batch_stride = act_bs; g.ld_out = p.K2p;
    g.outT_hi = d2th; g.outT_lo = d2tl; g.outT_batch_stride = actT_bs; g.ld_t = p.Kq;
    g.mask_hi = a2h; g.mask_lo = a2l; g.ld_mask = p.K2p; g.mask_batch_stride = act_bs;
    if (int rc = launch_f16_gemm(dlh, dll, Npts, C, p.K3p, w3th, w3tl, p.Nph, p.BNh, C, g, st)) return rc;
    // dW3^T[h][k] = sum_n a2T[h][n] dloT[k][n]  -> stored transposed = W3's own [k][h] layout
    g = fresh(); g.M = hidden; g.n_valid = ncls;
    g.outT_f32 = grad + oW3; g.outT_batch_stride = ld; g.ld_t = hidden; g.t_valid = ncls;
    if (int rc = launch_f16_gemm(a2th, a2tl, p.Nph, C, p.Kq, dlth, dltl, p.Npc, p.BNc, C, g, st)) return rc;
    // dW2[o][i] = sum_n da2T[o][n] a1T[i][n]
    g = fresh(); g.M = hidden; g.n_valid = hidden;
    g.out_f32 = grad + oW2; g.out_batch_stride = ld; g.ld_out = hidden;
    if (int rc = launch_f16_gemm(d2th, d2tl, p.Nph, C, p.Kq, a1th, a1tl, p.Nph, p.BNh, C, g, st)) return rc;
    // da1 = (da2 W2) . [a1 > 0]  -> only its transposed form is needed
    g = fresh(); g.M = Npts; g.n_valid = hidden;
    g.outT_hi = d1th; g.outT_lo = d1tl; g.outT_batch_stride = actT_bs; g.ld_t = p.Kq;
    g.mask_hi = a1h; g.mask_lo = a1l; g.ld_mask = p.K2p; g.mask_batch_stride = act_bs;
    if (int rc = launch_f16_gemm(d2h, d2l, Npts, C, p.K2p, w2th, w2tl, p.Nph, p.BNh, C, g, st)) return rc;
    // dW1[o][k] = sum_n da1T[o][n] XT[k][n]   (XT shared by all chains)
    g = fresh(); g.M = hidden; g.n_valid = in_dim; g.b_shared = 1;
    g.out_f32 = grad + oW1; g.out_batch_stride = ld; g.ld_out = in_dim;
    if (int rc = launch_f16_gemm(d1th, d1tl, p.Nph, C, p.Kq, xth, xtl, p.Npi, p.BNi, C, g, st)) return rc;
    // bias gradients = row sums of the transposed planes
    const int gb1 = (C * hidden + 7) / 8, gb3 = (C * ncls + 7) / 8;
    hmc_f16_bias_kernel<<<gb1, 256, 0, st>>>(d1th, d1tl, p.Nph, p.Kq, hidden, C, grad, ld, ob1);
    hmc_f16_bias_kernel<<<gb1, 256, 0, st>>>(d2th, d2tl, p.Nph, p.Kq, hidden, C, grad, ld, ob2);
    hmc_f16_bias_kernel<<<gb3, 256, 0, st>>>(dlth, dltl, p.Npc, p.Kq, ncls, C, grad, ld, ob3);
    URSA_LAUNCH_CHECK("hmc_f16_bias_kernel");
    return URSA_OK;
}
