// K3 (MLP): sample-batched BMA forward for models/mlp.py:8-23 straight from the [S, D] weight bank.
//   h1 = x W1_s^T + b1_s ; h2 = relu(h1) W2_s^T + b2_s ; logits = relu(h2) W3_s^T + b3_s
// followed by the fused softmax-average / entropy accumulation (bma_metrics.cu).
//
// URSA_ALGO_FFMA: fp32 CUDA-core batched GEMM (128x64x16 tiles, 8x4 register tile, double-buffered
// shared memory).  This is the bit-faithful fp32 path every other algo is checked against.
// URSA_ALGO_TCGEN05: 3xTF32 split GEMM on tcgen05 + TMA (bma_mlp_tc.cu).
#include "common.cuh"

namespace ursa {

int mlp_forward_tcgen05(const float *bank, int64_t ld_bank, int S, const float *x, int64_t N, int in_dim, int hidden,
                        int C, float *proba_sum, float *entropy_sum, float *logits_out, double gamma, void *workspace,
                        size_t workspace_bytes, cudaStream_t st);
size_t mlp_workspace_tcgen05(int S, int64_t N, int in_dim, int hidden, int C);
int mlp_forward_f16(const float *bank, int64_t ld_bank, int S, const float *x, int64_t N, int in_dim, int hidden, int C,
                    float *proba_sum, float *entropy_sum, float *logits_out, double gamma, void *workspace,
                    size_t workspace_bytes, cudaStream_t st);
size_t mlp_workspace_f16(int S, int64_t N, int in_dim, int hidden, int C);

constexpr int BM = 128, BN = 64, BK = 16, kGemmThreads = 256;

struct GemmArgs {
    const float *A;      // [batch?][M, K] row-major, lda
    const float *W;      // [batch][Nout, K] row-major, ldw = K
    const float *bias;   // [batch][Nout]
    float *Y;            // [batch][M, Nout] row-major, ldy
    int64_t strideA, strideW, strideBias, strideY;   // per-sample strides in elements (strideA may be 0)
    int64_t M;
    int Nout, K, lda, ldy;
    int relu_out;
};

// Y[b] = A[b] * W[b]^T + bias[b]   (optionally ReLU on store)
__global__ void __launch_bounds__(kGemmThreads) mlp_gemm_kernel(const GemmArgs a) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int b = blockIdx.z;
    const float *A = a.A + (int64_t)b * a.strideA;
    const float *W = a.W + (int64_t)b * a.strideW;
    const float *bias = a.bias + (int64_t)b * a.strideBias;
    float *Y = a.Y + (int64_t)b * a.strideY;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;             // 16 x 16 threads; thread tile 8 (M) x 4 (N)
    const int K = a.K;

    // global->smem mapping: A tile 128 x 16 = 512 float4 (2 per thread); W tile 64 x 16 = 256 float4 (1 per thread)
    const int lrow = tid >> 2, lk4 = (tid & 3) * 4;
    const bool vecA = (a.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
    const bool vecW = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15u) == 0);

    auto load4 = [&](const float *base, int64_t row, int64_t nrows, int ld, int k, bool vec) -> float4 {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows) {
            const float *ptr = base + row * ld + k;
            if (vec && k + 4 <= K) {
                r = __ldg(reinterpret_cast<const float4 *>(ptr));
            } else {
                if (k + 0 < K) r.x = __ldg(ptr + 0);
                if (k + 1 < K) r.y = __ldg(ptr + 1);
                if (k + 2 < K) r.z = __ldg(ptr + 2);
                if (k + 3 < K) r.w = __ldg(ptr + 3);
            }
        }
        return r;
    };
    auto stage_store = [&](int buf, const float4 &a0, const float4 &a1, const float4 &w0) {
        As[buf][lk4 + 0][lrow] = a0.x; As[buf][lk4 + 1][lrow] = a0.y; As[buf][lk4 + 2][lrow] = a0.z; As[buf][lk4 + 3][lrow] = a0.w;
        As[buf][lk4 + 0][lrow + 64] = a1.x; As[buf][lk4 + 1][lrow + 64] = a1.y; As[buf][lk4 + 2][lrow + 64] = a1.z; As[buf][lk4 + 3][lrow + 64] = a1.w;
        Bs[buf][lk4 + 0][lrow] = w0.x; Bs[buf][lk4 + 1][lrow] = w0.y; Bs[buf][lk4 + 2][lrow] = w0.z; Bs[buf][lk4 + 3][lrow] = w0.w;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    float4 ra0 = load4(A, m0 + lrow, a.M, a.lda, lk4, vecA);
    float4 ra1 = load4(A, m0 + lrow + 64, a.M, a.lda, lk4, vecA);
    float4 rw0 = load4(W, n0 + lrow, a.Nout, K, lk4, vecW);
    stage_store(0, ra0, ra1, rw0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            const int k = (kt + 1) * BK + lk4;
            ra0 = load4(A, m0 + lrow, a.M, a.lda, k, vecA);
            ra1 = load4(A, m0 + lrow + 64, a.M, a.lda, k, vecA);
            rw0 = load4(W, n0 + lrow, a.Nout, K, k, vecW);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            const float4 w = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stage_store(buf ^ 1, ra0, ra1, rw0);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.Nout) {
                float v = acc[i][j] + __ldg(bias + n);
                if (a.relu_out) v = fmaxf(v, 0.f);
                Y[m * a.ldy + n] = v;
            }
        }
    }
}

static int launch_gemm(const GemmArgs &g, int batch, cudaStream_t st) {
    dim3 grid((unsigned)((g.M + BM - 1) / BM), (unsigned)((g.Nout + BN - 1) / BN), (unsigned)batch);
    mlp_gemm_kernel<<<grid, kGemmThreads, 0, st>>>(g);
    URSA_LAUNCH_CHECK("mlp_gemm_kernel");
    return URSA_OK;
}

// samples per pass so that the two hidden activations stay around <= 128 MB (L2-sized working set)
static int ffma_chunk(int S, int64_t N, int hidden) {
    const int64_t per = N * (int64_t)hidden * 4 * 2;
    int64_t sc = per > 0 ? (int64_t)(128ll << 20) / per : S;
    if (sc < 1) sc = 1;
    return (int)(sc < S ? sc : S);
}

}  // namespace ursa

using namespace ursa;

extern "C" size_t ursa_bma_mlp_workspace(int S, int64_t N, int in_dim, int hidden, int C, int algo) {
    if (S < 1 || N < 1 || in_dim < 1 || hidden < 1 || C < 1) return 0;
    if (algo == URSA_ALGO_TCGEN05) return mlp_workspace_tcgen05(S, N, in_dim, hidden, C);
    if (algo == URSA_ALGO_TCGEN05_F16) return mlp_workspace_f16(S, N, in_dim, hidden, C);
    if (algo != URSA_ALGO_FFMA) return 0;
    const int sc = ffma_chunk(S, N, hidden);
    return (size_t)sc * (size_t)N * (size_t)(2 * hidden + C) * sizeof(float);
}

extern "C" int ursa_bma_mlp_forward(const float *bank, int64_t ld_bank, int S, const float *x, int64_t N, int in_dim,
                                    int hidden, int C, float *proba_sum, float *entropy_sum, float *logits_out,
                                    double gamma, void *workspace, size_t workspace_bytes, int algo, void *stream) {
    URSA_REQUIRE(bank && x && proba_sum && entropy_sum && workspace, "ursa_bma_mlp_forward: null pointer");
    URSA_REQUIRE(S >= 1 && N >= 1 && in_dim >= 1 && hidden >= 1 && C >= 1, "ursa_bma_mlp_forward: bad shape");
    const int64_t D = (int64_t)hidden * in_dim + hidden + (int64_t)hidden * hidden + hidden + (int64_t)C * hidden + C;
    URSA_REQUIRE(ld_bank >= D, "ursa_bma_mlp_forward: ld_bank (%lld) < D (%lld)", (long long)ld_bank, (long long)D);
    URSA_REQUIRE(workspace_bytes >= ursa_bma_mlp_workspace(S, N, in_dim, hidden, C, algo), "ursa_bma_mlp_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (algo == URSA_ALGO_TCGEN05)
        return mlp_forward_tcgen05(bank, ld_bank, S, x, N, in_dim, hidden, C, proba_sum, entropy_sum, logits_out, gamma,
                                   workspace, workspace_bytes, st);
    if (algo == URSA_ALGO_TCGEN05_F16)
        return mlp_forward_f16(bank, ld_bank, S, x, N, in_dim, hidden, C, proba_sum, entropy_sum, logits_out, gamma, workspace,
                               workspace_bytes, st);
    URSA_REQUIRE(algo == URSA_ALGO_FFMA, "ursa_bma_mlp_forward: unknown algo %d", algo);

    const int64_t oW1 = 0, ob1 = oW1 + (int64_t)hidden * in_dim, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)hidden * hidden,
                  oW3 = ob2 + hidden, ob3 = oW3 + (int64_t)C * hidden;
    const int sc = ffma_chunk(S, N, hidden);
    float *h1 = reinterpret_cast<float *>(workspace);
    float *h2 = h1 + (size_t)sc * N * hidden;
    float *lg = h2 + (size_t)sc * N * hidden;
    for (int s0 = 0; s0 < S; s0 += sc) {
        const int nb = (S - s0 < sc) ? (S - s0) : sc;
        const float *bk = bank + (int64_t)s0 * ld_bank;
        GemmArgs g;
        g.M = N; g.strideW = ld_bank; g.strideBias = ld_bank;
        // layer 1: x is shared by all samples (strideA = 0)
        g.A = x; g.lda = in_dim; g.strideA = 0; g.W = bk + oW1; g.bias = bk + ob1; g.K = in_dim; g.Nout = hidden;
        g.Y = h1; g.ldy = hidden; g.strideY = N * (int64_t)hidden; g.relu_out = 1;
        if (int rc = launch_gemm(g, nb, st)) return rc;
        // layer 2
        g.A = h1; g.lda = hidden; g.strideA = N * (int64_t)hidden; g.W = bk + oW2; g.bias = bk + ob2; g.K = hidden;
        g.Y = h2;
        if (int rc = launch_gemm(g, nb, st)) return rc;
        // layer 3 (no activation)
        g.A = h2; g.W = bk + oW3; g.bias = bk + ob3; g.Nout = C; g.Y = lg; g.ldy = C; g.strideY = N * (int64_t)C;
        g.relu_out = 0;
        if (int rc = launch_gemm(g, nb, st)) return rc;
        if (int rc = ursa_bma_accumulate(lg, nb, N, C, N * (int64_t)C, proba_sum, entropy_sum, gamma, stream)) return rc;
        if (logits_out)
            URSA_CUDA(cudaMemcpyAsync(logits_out + (size_t)s0 * N * C, lg, (size_t)nb * N * C * sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
    }
    return URSA_OK;
}
