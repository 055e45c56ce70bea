// K3-epilogue / K4: BMA accumulation from logits and the metric counters.
//  accumulate  reference tasks/prediction.py:52-75 (+ util.py:126-144)
//  metrics     reference tasks/prediction.py:79-102, 152-194
// One warp owns one test row; classes are strided over lanes (C <= 1024).  The sample loop is sequential
// per row, so the fp32 accumulation order equals the reference's list order.  Metric partials are
// reduced in a fixed order (warp -> block -> second launch) so counters and fp64 sums are reproducible.
#include "bma_epilogue.cuh"
#include "common.cuh"

namespace ursa {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kMaxPerLane = 32;          // C <= 32 * 32
constexpr int kMaxBins = 64;

template <int PER_LANE>
__global__ void __launch_bounds__(kRowThreads) bma_accumulate_kernel(const float *__restrict__ logits, int64_t S,
                                                                      int64_t N, int C, int64_t ld_sample,
                                                                      float *__restrict__ proba_sum,
                                                                      float *__restrict__ entropy_sum,
                                                                      float one_minus_gamma, float gamma_over_c) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= N) return;
    float P[PER_LANE];
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) P[j] = (lane + 32 * j < C) ? proba_sum[row * C + lane + 32 * j] : 0.f;
    float E = entropy_sum[row];
    for (int64_t s = 0; s < S; ++s) {
        const float *lr = logits + s * ld_sample + row * C;
        float x[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) x[j] = (lane + 32 * j < C) ? __ldg(lr + lane + 32 * j) : 0.f;
        softmax_accumulate_row<PER_LANE>(x, C, lane, one_minus_gamma, gamma_over_c, P, E);
    }
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j)
        if (lane + 32 * j < C) proba_sum[row * C + lane + 32 * j] = P[j];
    if (lane == 0) entropy_sum[row] = E;
}

// ---------------------------------------------------------------------------------------------
// workspace layout per block: int64[1 + 2*nb] then double[2 + nb]
struct MetricArgs {
    const float *proba_sum;
    const int64_t *targets;
    int64_t N;
    int C, n_bins;
    float num_samples, one_minus_gamma, gamma_over_c;
    int32_t *pred_out;
    float *conf_out;
    int64_t *ws_i64;
    double *ws_f64;
};

__global__ void __launch_bounds__(kRowThreads) bma_metrics_kernel(const MetricArgs a, int rows_per_block) {
    __shared__ long long s_cnt[kRowWarps][1 + 2 * kMaxBins];
    __shared__ double s_f64[kRowWarps][2 + kMaxBins];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = a.n_bins, C = a.C;
    for (int i = lane; i < 1 + 2 * nb; i += 32) s_cnt[warp][i] = 0;
    for (int i = lane; i < 2 + nb; i += 32) s_f64[warp][i] = 0.0;
    __syncwarp();
    const double step = 1.0 / (double)nb;                              // np.linspace: bounds[b] = b * step, last = 1.0
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
    for (int r = warp; r < rows_per_block; r += kRowWarps) {           // fixed row -> warp assignment
        const int64_t row = row0 + r;
        if (row >= a.N) break;
        const float *pr = a.proba_sum + row * C;
        const int64_t y = a.targets[row];
        float best = -INFINITY;
        int besti = 0x7fffffff;
        double brier = 0.0;
        float py = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float p = __fdiv_rn(__ldg(pr + c), a.num_samples);   // ensemble_proba / S  (:82)
            if (p > best) { best = p; besti = c; }                     // first maximum within the lane
            const double d = (double)p - (c == y ? 1.0 : 0.0);         // :192-194
            brier += d * d;
            if (c == y) py = p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                             // argmax: larger value, then smaller index
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        brier = warp_sum_d(brier);
        py = warp_sum(py);                                             // exactly one lane holds it
        if (lane == 0) {
            const bool ok = (int64_t)besti == y;
            const double conf = (double)best;
            int bin = -1;
            for (int b = 0; b < nb; ++b) {
                const double lo = (double)b * step;
                const double hi = (b == nb - 1) ? 1.0 : (double)(b + 1) * step;
                if (conf > lo && conf <= hi) bin = b;                  // (lo, hi]  (:169-170)
            }
            s_cnt[warp][0] += ok ? 1 : 0;
            if (bin >= 0) {
                s_cnt[warp][1 + bin] += 1;
                s_cnt[warp][1 + nb + bin] += ok ? 1 : 0;
                s_f64[warp][2 + bin] += conf;
            }
            const float q = __fadd_rn(__fmul_rn(a.one_minus_gamma, py), a.gamma_over_c);   // :88-90
            s_f64[warp][0] += -log((double)q);
            s_f64[warp][1] += brier;
            if (a.pred_out) a.pred_out[row] = besti;
            if (a.conf_out) a.conf_out[row] = best;
        }
    }
    __syncthreads();
    const int ni = 1 + 2 * nb, nf = 2 + nb;
    for (int i = threadIdx.x; i < ni; i += kRowThreads) {
        long long t = 0;
        for (int w = 0; w < kRowWarps; ++w) t += s_cnt[w][i];
        a.ws_i64[(int64_t)blockIdx.x * ni + i] = t;
    }
    for (int i = threadIdx.x; i < nf; i += kRowThreads) {
        double t = 0.0;
        for (int w = 0; w < kRowWarps; ++w) t += s_f64[w][i];
        a.ws_f64[(int64_t)blockIdx.x * nf + i] = t;
    }
}

__global__ void bma_metrics_reduce_kernel(const int64_t *ws_i64, const double *ws_f64, int nblocks, int nb,
                                          int64_t *out_i64, double *out_f64) {
    const int ni = 1 + 2 * nb, nf = 2 + nb;
    for (int i = threadIdx.x; i < ni; i += blockDim.x) {
        long long t = 0;
        for (int b = 0; b < nblocks; ++b) t += ws_i64[(int64_t)b * ni + i];
        out_i64[i] = t;
    }
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        double t = 0.0;
        for (int b = 0; b < nblocks; ++b) t += ws_f64[(int64_t)b * nf + i];
        out_f64[i] = t;
    }
}

constexpr int kMetricRowsPerBlock = 64;
static inline int metric_blocks(int64_t N) { return (int)((N + kMetricRowsPerBlock - 1) / kMetricRowsPerBlock); }

}  // namespace ursa

using namespace ursa;

extern "C" int ursa_bma_accumulate(const float *logits, int64_t S, int64_t N, int C, int64_t ld_sample,
                                   float *proba_sum, float *entropy_sum, double gamma, void *stream) {
    URSA_REQUIRE(logits && proba_sum && entropy_sum, "ursa_bma_accumulate: null pointer");
    URSA_REQUIRE(S >= 0 && N >= 0 && C >= 1 && C <= 32 * kMaxPerLane, "ursa_bma_accumulate: need 1 <= C <= %d", 32 * kMaxPerLane);
    URSA_REQUIRE(ld_sample >= N * C, "ursa_bma_accumulate: ld_sample < N*C");
    if (S == 0 || N == 0) return URSA_OK;
    const float omg = (float)(1.0 - gamma);
    const float goc = (float)(gamma * 1.0 / (double)C);
    const int grid = (int)((N + kRowWarps - 1) / kRowWarps);
    cudaStream_t st = (cudaStream_t)stream;
#define URSA_ACC(PL) bma_accumulate_kernel<PL><<<grid, kRowThreads, 0, st>>>(logits, S, N, C, ld_sample, proba_sum, entropy_sum, omg, goc)
    if (C <= 32) URSA_ACC(1);
    else if (C <= 128) URSA_ACC(4);
    else if (C <= 256) URSA_ACC(8);
    else URSA_ACC(kMaxPerLane);
#undef URSA_ACC
    URSA_LAUNCH_CHECK("bma_accumulate_kernel");
    return URSA_OK;
}

extern "C" size_t ursa_bma_metrics_workspace(int64_t N, int n_bins) {
    if (N < 0 || n_bins < 1 || n_bins > kMaxBins) return 0;
    const size_t nblk = (size_t)metric_blocks(N > 0 ? N : 1);
    return nblk * ((size_t)(1 + 2 * n_bins) * sizeof(int64_t) + (size_t)(2 + n_bins) * sizeof(double));
}

extern "C" int ursa_bma_metrics(const float *proba_sum, int64_t N, int C, float num_samples, const int64_t *targets,
                                double gamma, int n_bins, int64_t *out_i64, double *out_f64, int32_t *pred_out,
                                float *conf_out, void *workspace, size_t workspace_bytes, void *stream) {
    URSA_REQUIRE(proba_sum && targets && out_i64 && out_f64 && workspace, "ursa_bma_metrics: null pointer");
    URSA_REQUIRE(N >= 1 && C >= 1, "ursa_bma_metrics: need N >= 1 and C >= 1");
    URSA_REQUIRE(n_bins >= 1 && n_bins <= kMaxBins, "ursa_bma_metrics: n_bins must be in [1, %d]", kMaxBins);
    URSA_REQUIRE(num_samples > 0.f, "ursa_bma_metrics: num_samples must be > 0");
    URSA_REQUIRE(workspace_bytes >= ursa_bma_metrics_workspace(N, n_bins), "ursa_bma_metrics: workspace too small");
    URSA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "ursa_bma_metrics: workspace must be 8-byte aligned");
    const int nblk = metric_blocks(N);
    MetricArgs a;
    a.proba_sum = proba_sum; a.targets = targets; a.N = N; a.C = C; a.n_bins = n_bins;
    a.num_samples = num_samples;
    a.one_minus_gamma = (float)(1.0 - gamma);
    a.gamma_over_c = (float)(gamma * 1.0 / (double)C);
    a.pred_out = pred_out; a.conf_out = conf_out;
    a.ws_i64 = reinterpret_cast<int64_t *>(workspace);
    a.ws_f64 = reinterpret_cast<double *>(a.ws_i64 + (size_t)nblk * (1 + 2 * n_bins));
    cudaStream_t st = (cudaStream_t)stream;
    bma_metrics_kernel<<<nblk, kRowThreads, 0, st>>>(a, kMetricRowsPerBlock);
    URSA_LAUNCH_CHECK("bma_metrics_kernel");
    bma_metrics_reduce_kernel<<<1, 128, 0, st>>>(a.ws_i64, a.ws_f64, nblk, n_bins, out_i64, out_f64);
    URSA_LAUNCH_CHECK("bma_metrics_reduce_kernel");
    return URSA_OK;
}
