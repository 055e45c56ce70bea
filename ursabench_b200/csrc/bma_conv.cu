// K3 (PreResNet): sample-batched BMA forward for models/preresnet.py:90-151 (BasicBlock, depth = 6n+2 < 44)
// straight from the [S, D] weight bank and the [S, nb] BatchNorm running-statistics bank.
//
// Per chunk of (S_c samples x N_c images):
//   prep      bank rows -> packed inference weights: conv filters transposed to [ci][kh][kw][co], eval-mode BN
//             folded to per-channel (a, b) = (gamma/sqrt(var+eps), beta - mean*a)
//   conv3x3   direct convolution on CUDA cores (fp32 FFMA): a CTA owns G images of ONE sample (so the sample's
//             filters are staged in shared memory once), 1024 output pixels, thread tile 4 pixels x 16 channels;
//             the consuming conv applies its input BN + ReLU while staging the tile (zero padding AFTER the
//             activation, as PyTorch pads relu(bn(x))), and adds the residual in its epilogue
//   conv1x1   the stride-2 downsample shortcut on the RAW block input (preresnet.py:44-45)
//   head      BN + ReLU + 8x8 average pool + fc -> logits
//   accumulate (bma_metrics.cu) softmax-average + entropy in sample order
// Activations are NCHW planes per (sample, image) in the workspace; the first conv reads the shared input x.
// This is the fp32 reference-accuracy path (URSA_ALGO_FFMA); see DESIGN.md for the tcgen05 plan.
#include "common.cuh"
#include "preresnet_plan.cuh"

namespace ursa {

size_t preresnet_workspace_tcgen05(int S, int64_t N, int depth, int C);
size_t preresnet_workspace_fused(int S, int64_t N, int depth, int C, int f16);
int preresnet_forward_fused(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf, int S, const float *x,
                            int64_t N, int depth, int C, float *proba_sum, float *entropy_sum, float *logits_out,
                            double gamma, void *workspace, size_t workspace_bytes, int f16, int ws_kept, cudaStream_t st);
int preresnet_forward_tcgen05(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf, int S, const float *x,
                              int64_t N, int depth, int C, float *proba_sum, float *entropy_sum, float *logits_out,
                              double gamma, void *workspace, size_t workspace_bytes, cudaStream_t st);

// -----------------------------------------------------------------------------------------------------------
struct ConvArgs {
    const float *in;        // activations [S_c][N_c][CIN][HIN][HIN]  (in_sample_stride == 0: shared input x)
    const float *res;       // residual   [S_c][N_c][COUT][HOUT][HOUT] or nullptr
    float *out;             // output     [S_c][N_c][COUT][HOUT][HOUT]
    const float *packed;    // [S_c][ld_packed]
    int64_t ld_packed, w_off, bn_off;   // bn_off < 0: no input BN/ReLU (the network's first conv)
    int64_t in_sample_stride, in_image_stride;
    int n_images;           // N_c
};

template <int CIN, int COUT, int STRIDE, int HOUT, int G, int CK>
struct ConvCfg {
    static constexpr int HIN = HOUT * STRIDE;
    static constexpr int HP = (HOUT - 1) * STRIDE + 3;           // staged rows / cols incl. halo
    static constexpr int WP = HP + ((HP % 2 == 0) ? 1 : 0);      // odd row pitch: fewer bank conflicts
    static constexpr int NPX = G * HOUT * HOUT;
    static constexpr int PXT = NPX / 4;                          // pixel-threads (4 pixels each)
    static constexpr int CG = COUT / 16;                         // 16-channel groups
    static constexpr int THREADS = PXT * CG;
    static constexpr int IN_FLOATS = G * CK * HP * WP;
    static constexpr int W_FLOATS = CK * 9 * COUT;
    static constexpr size_t SMEM = (size_t)(IN_FLOATS + W_FLOATS) * sizeof(float);
    static_assert(NPX % 128 == 0 && COUT % 16 == 0 && CIN % CK == 0, "bad conv tiling");
    static_assert(THREADS <= 1024 && THREADS % 32 == 0, "bad thread count");
};

template <int CIN, int COUT, int STRIDE, int HOUT, int G, int CK>
__global__ void __launch_bounds__(ConvCfg<CIN, COUT, STRIDE, HOUT, G, CK>::THREADS, 1)
conv3x3_kernel(const ConvArgs a) {
    using Cfg = ConvCfg<CIN, COUT, STRIDE, HOUT, G, CK>;
    constexpr int HIN = Cfg::HIN, HP = Cfg::HP, WP = Cfg::WP, PXT = Cfg::PXT;
    extern __shared__ __align__(16) float smem[];
    float *in_s = smem;                         // [G][CK][HP][WP]
    float *w_s = smem + Cfg::IN_FLOATS;         // [CK][9][COUT]

    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int n0 = blockIdx.x * G;
    const float *pk = a.packed + (int64_t)s * a.ld_packed;
    const float *wsrc = pk + a.w_off;
    const float *bn = a.bn_off >= 0 ? pk + a.bn_off : nullptr;

    const int cg = tid / PXT, tp = tid - cg * PXT;     // warp-uniform channel group
    int base[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = tp + j * PXT;
        const int g = p / (HOUT * HOUT), r = p - g * (HOUT * HOUT);
        const int h = r / HOUT, w = r - h * HOUT;
        base[j] = g * (CK * HP * WP) + (h * STRIDE) * WP + w * STRIDE;
    }
    float acc[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[j][c] = 0.f;

    for (int c0 = 0; c0 < CIN; c0 += CK) {
        __syncthreads();                         // previous chunk fully consumed
        // stage the input tile: relu(bn(x)) inside the image, 0 in the halo
        for (int i = tid; i < Cfg::IN_FLOATS; i += Cfg::THREADS) {
            const int ww = i % WP;
            const int hh = (i / WP) % HP;
            const int ci = (i / (WP * HP)) % CK;
            const int g = i / (WP * HP * CK);
            const int hi = hh - 1, wi = ww - 1, n = n0 + g;
            float v = 0.f;
            if (ww < HP && hi >= 0 && hi < HIN && wi >= 0 && wi < HIN && n < a.n_images) {
                v = __ldg(a.in + (int64_t)s * a.in_sample_stride + (int64_t)n * a.in_image_stride +
                          ((int64_t)(c0 + ci) * HIN + hi) * HIN + wi);
                if (bn) v = fmaxf(fmaf(__ldg(bn + c0 + ci), v, __ldg(bn + CIN + c0 + ci)), 0.f);
            }
            in_s[i] = v;
        }
        for (int i = tid; i < Cfg::W_FLOATS; i += Cfg::THREADS) w_s[i] = __ldg(wsrc + (int64_t)c0 * 9 * COUT + i);
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int off = ci * (HP * WP) + kh * WP + kw;
                    float av[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) av[j] = in_s[base[j] + off];
                    const float4 *wp = reinterpret_cast<const float4 *>(w_s + (ci * 9 + kh * 3 + kw) * COUT + cg * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w4 = wp[q];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[j][4 * q + 0] = fmaf(av[j], w4.x, acc[j][4 * q + 0]);
                            acc[j][4 * q + 1] = fmaf(av[j], w4.y, acc[j][4 * q + 1]);
                            acc[j][4 * q + 2] = fmaf(av[j], w4.z, acc[j][4 * q + 2]);
                            acc[j][4 * q + 3] = fmaf(av[j], w4.w, acc[j][4 * q + 3]);
                        }
                    }
                }
            }
        }
    }
    // epilogue: (+ residual) -> NCHW planes; lanes are consecutive pixels -> coalesced
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = tp + j * PXT;
        const int g = p / (HOUT * HOUT), r = p - g * (HOUT * HOUT);
        const int n = n0 + g;
        if (n >= a.n_images) continue;
        const int64_t o = (((int64_t)s * a.n_images + n) * COUT + cg * 16) * (HOUT * HOUT) + r;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float v = acc[j][c];
            if (a.res) v += __ldg(a.res + o + (int64_t)c * (HOUT * HOUT));
            a.out[o + (int64_t)c * (HOUT * HOUT)] = v;
        }
    }
}

// 1x1 stride-2 shortcut conv on the raw block input: out[co][h][w] = sum_ci W[ci][co] * in[ci][2h][2w]
template <int CIN, int COUT, int HOUT>
__global__ void __launch_bounds__(256) conv1x1s2_kernel(const ConvArgs a) {
    constexpr int HIN = HOUT * 2;
    __shared__ float w_s[CIN * COUT];
    const int s = blockIdx.y, n = blockIdx.x;
    const float *pk = a.packed + (int64_t)s * a.ld_packed + a.w_off;
    for (int i = threadIdx.x; i < CIN * COUT; i += blockDim.x) w_s[i] = __ldg(pk + i);
    __syncthreads();
    const float *in = a.in + (int64_t)s * a.in_sample_stride + (int64_t)n * a.in_image_stride;
    float *out = a.out + ((int64_t)s * a.n_images + n) * COUT * HOUT * HOUT;
    for (int i = threadIdx.x; i < COUT * HOUT * HOUT; i += blockDim.x) {
        const int r = i % (HOUT * HOUT), co = i / (HOUT * HOUT);
        const int h = r / HOUT, w = r - h * HOUT;
        float acc = 0.f;
#pragma unroll 8
        for (int ci = 0; ci < CIN; ++ci)
            acc = fmaf(__ldg(in + ((int64_t)ci * HIN + 2 * h) * HIN + 2 * w), w_s[ci * COUT + co], acc);
        out[i] = acc;
    }
}

// final BN + ReLU + AvgPool2d(8) + fc; one warp per (sample, image)
__global__ void __launch_bounds__(256) preresnet_head_kernel(const float *__restrict__ act, const float *__restrict__ packed,
                                                              int64_t ld_packed, int64_t bn_off, int64_t fc_off,
                                                              int n_images, int n_pairs, int C, float *__restrict__ logits) {
    __shared__ float feat_s[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 8 + warp;
    if (pair >= n_pairs) return;
    const int s = pair / n_images;
    const float *pk = packed + (int64_t)s * ld_packed;
    const float *x = act + (int64_t)pair * 64 * 64;
    for (int ch = 0; ch < 64; ++ch) {
        const float aa = __ldg(pk + bn_off + ch), bb = __ldg(pk + bn_off + 64 + ch);
        float v = fmaxf(fmaf(aa, __ldg(x + ch * 64 + lane), bb), 0.f) + fmaxf(fmaf(aa, __ldg(x + ch * 64 + 32 + lane), bb), 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) feat_s[warp][ch] = v * (1.f / 64.f);
    }
    __syncwarp();
    const float *fw = pk + fc_off, *fb = pk + fc_off + (int64_t)C * 64;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) acc = fmaf(feat_s[warp][k], __ldg(fw + c * 64 + k), acc);
        logits[(int64_t)pair * C + c] = acc + __ldg(fb + c);
    }
}

// -----------------------------------------------------------------------------------------------------------
constexpr int kChunkSamples = 8, kChunkImages = 512;
constexpr int64_t kActFloats = 16 * 32 * 32;      // largest activation per (sample, image): 64 KB

struct Chunking {
    int sc, nc;
    size_t act_bytes, packed_bytes, logit_bytes, total;
};

static Chunking chunking(int S, int64_t N, const NetPlan &pl) {
    Chunking c;
    c.sc = S < kChunkSamples ? S : kChunkSamples;
    c.nc = (int)(N < kChunkImages ? N : kChunkImages);
    const size_t pairs = (size_t)c.sc * c.nc;
    c.act_bytes = pairs * kActFloats * sizeof(float);
    c.packed_bytes = (size_t)c.sc * pl.packed_floats * sizeof(float);
    c.logit_bytes = ((pairs * pl.C + 3) & ~(size_t)3) * sizeof(float);
    c.total = 4 * c.act_bytes + c.packed_bytes + c.logit_bytes;
    return c;
}

template <int CIN, int COUT, int STRIDE, int HOUT, int G, int CK>
static int launch_conv(ConvArgs a, int sc, cudaStream_t st) {
    using Cfg = ConvCfg<CIN, COUT, STRIDE, HOUT, G, CK>;
    auto kern = conv3x3_kernel<CIN, COUT, STRIDE, HOUT, G, CK>;
    URSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    dim3 grid((a.n_images + G - 1) / G, sc);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    URSA_LAUNCH_CHECK("conv3x3_kernel");
    return URSA_OK;
}

}  // namespace ursa

using namespace ursa;

extern "C" size_t ursa_bma_preresnet_workspace(int S, int64_t N, int depth, int C, int algo) {
    algo &= ~URSA_ALGO_FLAG_WS_KEPT;
    NetPlan pl;
    if (S < 1 || N < 1) return 0;
    if (algo == URSA_ALGO_TCGEN05) return preresnet_workspace_tcgen05(S, N, depth, C);
    if (algo == URSA_ALGO_TCGEN05_FUSED || algo == URSA_ALGO_TCGEN05_FUSED_F16)
        return preresnet_workspace_fused(S, N, depth, C, algo == URSA_ALGO_TCGEN05_FUSED_F16);
    if (algo != URSA_ALGO_FFMA || !build_plan(depth, C, pl)) return 0;
    return chunking(S, N, pl).total;
}

extern "C" int ursa_bma_preresnet_forward(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf,
                                          int S, const float *x, int64_t N, int depth, int C, float *proba_sum,
                                          float *entropy_sum, float *logits_out, double gamma, void *workspace,
                                          size_t workspace_bytes, int algo, void *stream) {
    URSA_REQUIRE(bank && bufbank && x && proba_sum && entropy_sum && workspace, "ursa_bma_preresnet_forward: null pointer");
    URSA_REQUIRE(S >= 1 && N >= 1, "ursa_bma_preresnet_forward: bad shape");
    const int ws_kept = (algo & URSA_ALGO_FLAG_WS_KEPT) != 0;
    algo &= ~URSA_ALGO_FLAG_WS_KEPT;
    if (algo == URSA_ALGO_TCGEN05)
        return preresnet_forward_tcgen05(bank, ld_bank, bufbank, ld_buf, S, x, N, depth, C, proba_sum, entropy_sum,
                                         logits_out, gamma, workspace, workspace_bytes, (cudaStream_t)stream);
    if (algo == URSA_ALGO_TCGEN05_FUSED || algo == URSA_ALGO_TCGEN05_FUSED_F16)
        return preresnet_forward_fused(bank, ld_bank, bufbank, ld_buf, S, x, N, depth, C, proba_sum, entropy_sum,
                                       logits_out, gamma, workspace, workspace_bytes, algo == URSA_ALGO_TCGEN05_FUSED_F16,
                                       ws_kept, (cudaStream_t)stream);
    if (algo != URSA_ALGO_FFMA) {
        set_error("ursa_bma_preresnet_forward: unknown algo %d", algo);
        return URSA_ERR_UNSUPPORTED;
    }
    static thread_local NetPlan pl;     // ~3 KB; rebuilt per call (cheap)
    if (!build_plan(depth, C, pl)) {
        set_error("ursa_bma_preresnet_forward: unsupported depth %d (BasicBlock PreResNet: depth = 6n+2, 8..38)", depth);
        return URSA_ERR_UNSUPPORTED;
    }
    URSA_REQUIRE(ld_bank >= pl.D, "ursa_bma_preresnet_forward: ld_bank (%lld) < D (%lld)", (long long)ld_bank, (long long)pl.D);
    URSA_REQUIRE(ld_buf >= pl.NB, "ursa_bma_preresnet_forward: ld_buf (%lld) < %lld", (long long)ld_buf, (long long)pl.NB);
    const Chunking ck = chunking(S, N, pl);
    URSA_REQUIRE(workspace_bytes >= ck.total, "ursa_bma_preresnet_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char *wsb = reinterpret_cast<char *>(workspace);
    float *actA = reinterpret_cast<float *>(wsb);
    float *actB = reinterpret_cast<float *>(wsb + ck.act_bytes);
    float *actC = reinterpret_cast<float *>(wsb + 2 * ck.act_bytes);
    float *actD = reinterpret_cast<float *>(wsb + 3 * ck.act_bytes);
    float *packed = reinterpret_cast<float *>(wsb + 4 * ck.act_bytes);
    float *logits = reinterpret_cast<float *>(wsb + 4 * ck.act_bytes + ck.packed_bytes);
    const int n = pl.n_blocks;

    for (int s0 = 0; s0 < S; s0 += ck.sc) {
        const int sc = (S - s0 < ck.sc) ? (S - s0) : ck.sc;
        preresnet_prep_kernel<<<dim3(pl.table.n, sc), 256, 0, st>>>(pl.table, bank + (int64_t)s0 * ld_bank, ld_bank,
                                                                    bufbank + (int64_t)s0 * ld_buf, ld_buf, packed,
                                                                    pl.packed_floats);
        URSA_LAUNCH_CHECK("preresnet_prep_kernel");
        for (int64_t i0 = 0; i0 < N; i0 += ck.nc) {
            const int nc = (int)((N - i0 < ck.nc) ? (N - i0) : ck.nc);
            ConvArgs a;
            a.packed = packed; a.ld_packed = pl.packed_floats; a.n_images = nc;
            // network conv1: shared input x, no BN
            a.in = x + i0 * 3 * 32 * 32; a.in_sample_stride = 0; a.in_image_stride = 3 * 32 * 32;
            a.res = nullptr; a.out = actA; a.w_off = pl.conv1_w; a.bn_off = -1;
            if (int rc = launch_conv<3, 16, 1, 32, 2, 3>(a, sc, st)) return rc;
            float *cur = actA, *mid = actB, *nxt = actC, *dsb = actD;
            int ch = 16, hw = 32;
            for (int stg = 0; stg < 3; ++stg) {
                for (int b = 0; b < n; ++b) {
                    const NetPlan::Block &B = pl.blocks[stg][b];
                    const bool down = B.ds >= 0;
                    const int cout = down ? ch * 2 : ch;
                    const int hout = down ? hw / 2 : hw;
                    a.in = cur; a.in_image_stride = (int64_t)ch * hw * hw; a.in_sample_stride = (int64_t)nc * a.in_image_stride;
                    // shortcut
                    const float *res = cur;
                    if (down) {
                        ConvArgs d = a;
                        d.out = dsb; d.w_off = B.ds; d.res = nullptr; d.bn_off = -1;
                        dim3 grid(nc, sc);
                        if (ch == 16) conv1x1s2_kernel<16, 32, 16><<<grid, 256, 0, st>>>(d);
                        else conv1x1s2_kernel<32, 64, 8><<<grid, 256, 0, st>>>(d);
                        URSA_LAUNCH_CHECK("conv1x1s2_kernel");
                        res = dsb;
                    }
                    // conv1 (stride on the first block of stages 2, 3)
                    a.res = nullptr; a.out = mid; a.w_off = B.w1; a.bn_off = B.bn1;
                    int rc;
                    if (ch == 16 && !down) rc = launch_conv<16, 16, 1, 32, 2, 16>(a, sc, st);
                    else if (ch == 16 && down) rc = launch_conv<16, 32, 2, 16, 4, 4>(a, sc, st);
                    else if (ch == 32 && !down) rc = launch_conv<32, 32, 1, 16, 4, 16>(a, sc, st);
                    else if (ch == 32 && down) rc = launch_conv<32, 64, 2, 8, 8, 8>(a, sc, st);
                    else rc = launch_conv<64, 64, 1, 8, 8, 16>(a, sc, st);
                    if (rc) return rc;
                    // conv2 + residual
                    a.in = mid; a.in_image_stride = (int64_t)cout * hout * hout; a.in_sample_stride = (int64_t)nc * a.in_image_stride;
                    a.res = res; a.out = nxt; a.w_off = B.w2; a.bn_off = B.bn2;
                    if (cout == 16) rc = launch_conv<16, 16, 1, 32, 2, 16>(a, sc, st);
                    else if (cout == 32) rc = launch_conv<32, 32, 1, 16, 4, 16>(a, sc, st);
                    else rc = launch_conv<64, 64, 1, 8, 8, 16>(a, sc, st);
                    if (rc) return rc;
                    float *t = cur; cur = nxt; nxt = t;
                    ch = cout; hw = hout;
                }
            }
            const int pairs = sc * nc;
            preresnet_head_kernel<<<(pairs + 7) / 8, 256, 0, st>>>(cur, packed, pl.packed_floats, pl.bn_final, pl.fc, nc,
                                                                  pairs, C, logits);
            URSA_LAUNCH_CHECK("preresnet_head_kernel");
            if (int rc = ursa_bma_accumulate(logits, sc, nc, C, (int64_t)nc * C, proba_sum + i0 * C, entropy_sum + i0, gamma,
                                             stream))
                return rc;
            if (logits_out)
                URSA_CUDA(cudaMemcpy2DAsync(logits_out + ((int64_t)s0 * N + i0) * C, (size_t)N * C * sizeof(float), logits,
                                            (size_t)nc * C * sizeof(float), (size_t)nc * C * sizeof(float), sc,
                                            cudaMemcpyDeviceToDevice, st));
        }
    }
    return URSA_OK;
}
