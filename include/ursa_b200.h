/*
 * ursa_b200.h -- C ABI of the B200-native URSABench hot path.
 *
 * The reference (reml-lab/URSABench) is pure Python on PyTorch and has no FFI:
 * its "plugin interface" for this path is the Python class API
 * (inference/inference_base.py:12-56, tasks/task_base.py:4-20).  This header is
 * the boundary *below* that API: each entry point replaces the implicit ATen
 * launch sequence of one reference function (cited per function, paths relative
 * to /root/reference/URSABench/).  `ursabench_b200/_C.py` binds it with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers + sizes, no torch / C++ types;
 *  - the caller owns all memory; nothing here allocates device memory
 *    (workspaces are sized by the *_workspace() helpers and passed in);
 *  - every launch is asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - return value: 0 = URSA_OK, negative = error; ursa_last_error() returns a
 *    thread-local message for the last failing call on this thread;
 *  - no global mutable state besides that message and the launch counter: thread-compatible;
 *  - all arithmetic is fp32 unless stated; "n" counts elements, "ld" strides
 *    are in elements.
 */
#ifndef URSA_B200_H
#define URSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define URSA_OK               0
#define URSA_ERR_INVALID     -1   /* bad argument (null pointer, bad size, misalignment) */
#define URSA_ERR_CUDA        -2   /* CUDA runtime error; message carries cudaGetErrorString */
#define URSA_ERR_UNSUPPORTED -3   /* shape / architecture not covered by this build */

#define URSA_ABI_VERSION 1

int ursa_abi_version(void);
const char *ursa_last_error(void);
/* Number of kernel launches this library has issued in this process so far (monotonic, all threads, all devices):
 * a diagnostic counter for benchmarks ("gpu_launches"); nothing on the path reads it. */
uint64_t ursa_launch_count(void);
/* Per-kernel timing of the PreResNet BMA forward with CUDA events recorded on the launching stream (diagnostic for
 * bench.py's `roofline`: the kernel's average launch duration measured live inside the timed step).  Between
 * ursa_profile_begin(capacity) and ursa_profile_end() every instrumented launch records an event pair (at most
 * `capacity` pairs); ursa_profile_end synchronises on them and returns, per kind, the summed milliseconds and the
 * number of launches.  Process-wide; do not use from two threads at once. */
#define URSA_PROF_STAGE_C16  0   /* preresnet_stage16_kernel<16> / preresnet_stage_kernel<16> */
#define URSA_PROF_STAGE_C32  1
#define URSA_PROF_STAGE_C64  2
#define URSA_PROF_STEM       3
#define URSA_PROF_SHORTCUT   4
#define URSA_PROF_CONV_S2    5   /* stride-2 transition conv (conv3x3_tc_kernel) */
#define URSA_PROF_HEAD       6   /* head + softmax-average accumulation */
#define URSA_PROF_PREP       7
#define URSA_PROF_KINDS      8
int ursa_profile_begin(int capacity);
int ursa_profile_end(double *ms_sum, int64_t *count, int n_kinds);
/* SM count and compute capability of the current device. */
int ursa_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------
 * K1  fused SG-MCMC update  (replaces optimSGHMC.step, inference/optim_sghmc.py:30-68:
 *     7-8 ATen launches per parameter tensor -> ONE launch over the flat buffer)
 *
 *   d = g + wd_over_n * p                                   (:47-48; skipped when wd_over_n == 0)
 *   momentum != 0:  if FIRST: v = d                         (:51-52)
 *                   v = momentum * v - lr * d ; u = v       (:53/:56, :60)
 *   momentum == 0:  u = -lr * d                             (:62)
 *   NOISE:          u = u + (z * noise_mul) / noise_div     (:63-64; z ~ N(0,1))
 *   p = p + u                                               (:65)
 *   momentum != 0:  v = u                                   (:66-67; the noise is stored in the momentum)
 *   snapshot != NULL: snapshot[i] = p[i]                    (replaces deepcopy(model.cpu()), sghmc.py:99)
 *   ZERO_GRAD:      g = 0                                   (fused optimizer.zero_grad(), sghmc.py:79)
 *
 * `noise` != NULL : z is read from `noise` (parity mode; bit-exact vs. the reference's CPU arithmetic).
 * `noise` == NULL : z is drawn in-register: Philox4x32-10, key = seed, counter =
 *                   ((elem_offset+i)/4, step); Box-Muller; lane (elem_offset+i)%4.  The scale is applied
 *                   as one multiply by noise_mul/noise_div.
 * A [chains, D] buffer is just n = chains*D elements (the update is elementwise and the schedule is
 * shared); ranks pass their global element base in `elem_offset` so chains never share a noise stream.
 * p, g, v, snapshot, noise must be 16-byte aligned.  v may be NULL iff momentum == 0.
 * ---------------------------------------------------------------------- */
#define URSA_STEP_FIRST     1u
#define URSA_STEP_NOISE     2u
#define URSA_STEP_ZERO_GRAD 4u

int ursa_sgmcmc_step(float *p, float *g, float *v, float *snapshot, const float *noise,
                     int64_t n, float lr, float momentum, float wd_over_n,
                     float noise_mul, float noise_div, uint32_t flags,
                     uint64_t seed, uint64_t step, uint64_t elem_offset, void *stream);

/* Same update with the per-step scalars read from DEVICE memory, so the launch can sit inside a captured CUDA graph
 * whose replays see new values: dyn = [lr, momentum, wd_over_n, noise_scale (= noise_mul/noise_div; 0 gates the
 * noise off), step_lo, step_hi] (six 32-bit words; the step counter words are raw bit patterns).  The momentum
 * branch is taken iff v != NULL.  In external-noise mode the noise term is z * noise_scale. */
int ursa_sgmcmc_step_dyn(float *p, float *g, float *v, float *snapshot, const float *noise, int64_t n,
                         const float *dyn, uint32_t flags, uint64_t seed, uint64_t elem_offset, void *stream);

/* Write the six dyn words from by-value arguments (a 1-thread launch: the values are captured at launch time, so
 * the host may run ahead of the device without racing on a staging buffer). */
int ursa_sgmcmc_set_dyn(float *dyn, float lr, float momentum, float wd_over_n, float noise_scale, uint64_t step,
                        void *stream);

/* Fill `out[n]` with the N(0,1) stream ursa_sgmcmc_step would consume (test / diagnostics hook). */
int ursa_philox_normal(float *out, int64_t n, uint64_t seed, uint64_t step, uint64_t elem_offset, void *stream);

/* ------------------------------------------------------------------------
 * K2a SWAG moment + deviation-row update  (replaces SWA._collect_model, inference/swa.py:79-90 and
 *     CovarianceSpace.collect_vector, inference/subspaces.py:85-89: T D2H copies + 7 CPU passes + a
 *     K x D torch.cat -> one pass, 24 B/param, ring row written in place)
 *   mean = mean * keep + w / denom ; sq = sq * keep + (w*w) / denom ; dev_row = w - mean
 *   keep = n/(n+1), denom = n+1 computed by the caller in double (swa.py:83-88).
 * ---------------------------------------------------------------------- */
int ursa_swag_collect(const float *w, float *mean, float *sq_mean, float *dev_row,
                      int64_t n, float keep, float denom, void *stream);

/* var = max(sq_mean - mean^2, clamp)     (SWA._get_mean_and_variance, inference/swa.py:106-108) */
int ursa_swag_variance(const float *mean, const float *sq_mean, float *var, int64_t n, float clamp,
                       void *stream);

/* ------------------------------------------------------------------------
 * K2b batched SWAG draw  (the formula of inference/swag.py:85-97; the reference discards its draw at :98 and
 *     its low-rank branch raises -- see DESIGN.md "reference quirks")
 *   out[s, :] = mean + sqrt(var) * z1[s, :] + (ring^T z2[s, :]) / rank_div        s = 0..S-1
 * ring: [K, ld_ring] deviation rows (K == 0 -> diagonal draw); z2: [S, K] device, row-major;
 * z1: [S, ld_z1] device, or NULL to draw z1 in-register from Philox4x32-10 (key = seed, counter = (block, step)): element
 * (s, d) is normal s % 6 of block (s / 6) * D + d -- a block serves SIX consecutive draws of one column; its 128 bits are cut
 * into three (24-bit radius uniform, 18-bit angle) Box-Muller pairs (csrc/common.cuh::box_muller6; oracle/restate.py::
 * draw_normals restates it).  A diagonal draw with mean = 0, var = 1 returns exactly this stream.  All S draws are produced in
 * ONE pass over the ring: (K + 2 + S) * 4 B/param instead of S * (K + 3) * 4.  The K x S contraction runs on tcgen05 (3xTF32,
 * fp32 accumulate in TMEM).  Rows (ring, out, z1) must be 16-byte aligned: ld_* % 4 == 0.  K <= URSA_DRAW_MAX_K; any S >= 1
 * (draws are processed URSA_DRAW_MAX_S per launch, so the ring is read once per group of 30 draws; the Philox stream does not
 * depend on the grouping).
 * ---------------------------------------------------------------------- */
#define URSA_DRAW_MAX_S 30
#define URSA_DRAW_MAX_K 24

int ursa_swag_draw(float *out, int64_t ld_out, const float *mean, const float *var,
                   const float *ring, int64_t ld_ring, int K, const float *z2,
                   const float *z1, int64_t ld_z1, int S, int64_t D, float rank_div,
                   uint64_t seed, uint64_t step, void *stream);

/* ------------------------------------------------------------------------
 * K2c Gram matrix of the deviation ring, gram[K, K] (fp64, device) = R R^T over ring[K, D] in one streaming pass.
 *     First half of the PCA subspace (inference/subspaces.py:116-156 runs a randomized SVD of the K x D matrix on
 *     the host): with K <= URSA_DRAW_MAX_K rows, s and U come from eigh(gram) and the components s V^T = U^T R from
 *     ursa_swag_draw with z2 = U^T, var = 0 (the same call evaluates SubspaceModel.forward, mean + P^T t,
 *     inference/projection_model.py:13-14).
 * ---------------------------------------------------------------------- */
int ursa_swag_gram(const float *ring, int64_t ld_ring, int K, int64_t D, double *gram, void *stream);

/* ------------------------------------------------------------------------
 * K3  BMA accumulation  (replaces the inner loop of Prediction.update_statistics,
 *     tasks/prediction.py:52-75: per sample softmax twice + 2 D2H + CPU accumulate)
 *   for s in order: p = softmax(logits[s, i, :]) ; proba_sum[i, :] += p ;
 *                   entropy_sum[i] += -sum_c q log q,  q = (1-gamma) p + gamma/C   (util.py:126-144)
 * logits: [S, N, C] with sample stride ld_sample (elements).  Accumulation order over s is sequential,
 * as in the reference, so results do not depend on the launch shape.
 * ---------------------------------------------------------------------- */
int ursa_bma_accumulate(const float *logits, int64_t S, int64_t N, int C, int64_t ld_sample,
                        float *proba_sum, float *entropy_sum, double gamma, void *stream);

/* ------------------------------------------------------------------------
 * K4  BMA metric counters  (replaces Prediction.get_performance_metrics + _get_ece + _get_brier,
 *     tasks/prediction.py:79-102,152-194: numpy on [N, C] -> counters the host turns into metrics)
 *   pbar = proba_sum / num_samples (fp32 division) ; pred = first argmax ; conf = max
 *   out_i64 = [correct, bin_count[n_bins], bin_correct[n_bins]]            (exact)
 *   out_f64 = [nll_sum, brier_sum, bin_conf_sum[n_bins]]                   (fp64, fixed reduction order)
 *   bins are (b/n_bins, (b+1)/n_bins] compared in fp64 (np.linspace(0,1,n_bins+1), :160-170)
 *   nll_sum = sum_i -log((1-gamma) pbar[i, y_i] + gamma/C) ; brier_sum = sum_i sum_c (pbar - onehot)^2
 *   pred_out / conf_out ([N], nullable) receive the per-row argmax / confidence.
 * Deterministic: per-block partials in `workspace`, reduced in block order by a second launch.
 * ---------------------------------------------------------------------- */
size_t ursa_bma_metrics_workspace(int64_t N, int n_bins);
int ursa_bma_metrics(const float *proba_sum, int64_t N, int C, float num_samples, const int64_t *targets,
                     double gamma, int n_bins, int64_t *out_i64, double *out_f64,
                     int32_t *pred_out, float *conf_out, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ------------------------------------------------------------------------
 * K3  sample-batched BMA forward, MLP  (replaces S x ceil(N/B) calls of model(x) + the accumulation above
 *     for models/mlp.py:8-23).  bank: [S, ld_bank] flat weight vectors in model.parameters() order
 *     (fc1.weight[h,in], fc1.bias[h], fc2.weight[h,h], fc2.bias[h], fc3.weight[C,h], fc3.bias[C]).
 *     x: [N, in_dim].  Accumulates into proba_sum [N, C] / entropy_sum [N] in sample order (layer 3 writes logits to
 *     the workspace, then ONE ursa_bma_accumulate launch per chunk: tiles of different samples finish in any order).
 *     logits_out (nullable): [S, N, C].
 *     algo: URSA_ALGO_FFMA (fp32 CUDA cores), URSA_ALGO_TCGEN05 (3xTF32 on tcgen05 + TMA, one tile per CTA) or
 *     URSA_ALGO_TCGEN05_F16 (the product path: 2xFP16-split operands -- half the operand bytes and twice the K per MMA of
 *     3xTF32 at the same 22 significant bits -- on a persistent tcgen05 kernel whose MMAs run under the previous tile's
 *     epilogue; operands beyond fp16's range, |x| > 65 504, surface as NaN logits, never as finite wrong values).
 * ---------------------------------------------------------------------- */
#define URSA_ALGO_FFMA    0
#define URSA_ALGO_TCGEN05 1
#define URSA_ALGO_TCGEN05_FUSED 2   /* PreResNet only: stage-fused 3xTF32 kernel, activations in shared memory, residual in TMEM */
#define URSA_ALGO_FLAG_WS_KEPT 0x100 /* OR-ed into `algo` of ursa_bma_preresnet_forward: the workspace is the one the caller's PREVIOUS
                                      * call of this entry used, with the same min(S, 8), N, depth and C, and nothing else has written
                                      * to it since -- the FP16-split engine then skips re-zeroing the pad positions of its plane images */
#define URSA_ALGO_TCGEN05_F16 4     /* MLP: persistent 2xFP16-split GEMM kernel (csrc/bma_mlp_f16.cu); WideResNet: the FP16-split
                                     * instantiation of the conv kernel.  fp16's range: overflow surfaces as NaN logits */
#define URSA_ALGO_TCGEN05_FUSED_F16 3 /* PreResNet only: stage-fused kernel on 2xFP16-split operands (22 significant bits, fp32
                                       * accumulate), MMA / epilogue wavefront per 128-position tile.  Activations above ~1e6
                                       * overflow FP16 and surface as NaN logits (never as finite wrong values). */

size_t ursa_bma_mlp_workspace(int S, int64_t N, int in_dim, int hidden, int C, int algo);
int ursa_bma_mlp_forward(const float *bank, int64_t ld_bank, int S, const float *x, int64_t N,
                         int in_dim, int hidden, int C, float *proba_sum, float *entropy_sum,
                         float *logits_out, double gamma, void *workspace, size_t workspace_bytes,
                         int algo, void *stream);

/* ------------------------------------------------------------------------
 * K3' batched GEMM on the MLP engine (3xTF32, two-level accumulation, fp32 in / out):
 *       out[b, m, n] = sum_k A[b or shared, m, k] * B[b, n, k]  (+ bias[b, n])  (ReLU)
 *     A: [batch or 1][M][lda] (a_batch_stride = 0: one A shared by the batch), B: [batch][N][ldb], out: [batch][M][ldo].
 *     Used by the chain-batched HMC likelihood gradient of the MLPs (config 4), where hamiltorch / torch.func run
 *     one autograd graph per chain (inference/hmc.py:71-75).  Workspace holds the TF32 hi / lo planes.
 * ---------------------------------------------------------------------- */
size_t ursa_gemm_nt_3xtf32_workspace(int batch, int64_t M, int N, int K, int a_batched);
int ursa_gemm_nt_3xtf32(const float *A, int64_t lda, int64_t a_batch_stride, const float *B, int64_t ldb,
                        int64_t b_batch_stride, const float *bias, int64_t bias_stride, int relu,
                        float *out, int64_t ldo, int64_t out_batch_stride, int batch, int64_t M, int N, int K,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * K5' chain-batched likelihood gradient of the 3-layer MLP for HMC (replaces the per-chain autograd pass hamiltorch runs
 *     per leapfrog step, inference/hmc.py:71-75; BASELINE.json configs[3]).  theta: [C, ld] chain states in the MLP's flat
 *     layout (fc1.weight[h,in], fc1.bias, fc2.weight[h,h], fc2.bias, fc3.weight[k,h], fc3.bias); x: [N, in] shared by all
 *     chains; y: [N] int64 labels.  grad[c, :D] = d/dtheta sum_n CE(f_theta_c(x_n), y_n), ce[c] = that sum (fp32, fixed
 *     reduction order).  Eight 3xTF32 tcgen05 GEMMs; every operand is produced in split form by the epilogue of the GEMM
 *     before it, activations also transposed for the weight-gradient GEMMs.  in % 4 == 0, h % 4 == 0, k <= 128.
 * ---------------------------------------------------------------------- */
size_t ursa_hmc_mlp_grad_workspace(int C, int64_t N, int in_dim, int hidden, int n_classes);
int ursa_hmc_mlp_grad(const float *theta, int64_t ld, int C, const float *x, const int64_t *y, int64_t N,
                      int in_dim, int hidden, int n_classes, float *grad, float *ce,
                      void *workspace, size_t workspace_bytes, void *stream);
/* The same gradient on the persistent 2xFP16-split GEMM kernel (csrc/bma_mlp_f16.cu; planes are halves, x = hi + lo' 2^-11):
 * the product path of inference.HMC for the reference's MLPs.  Operands beyond fp16's range (|x| > 65 504) make grad / ce
 * non-finite -- never finite-but-wrong; callers fall back to ursa_hmc_mlp_grad, which has fp32's range. */
size_t ursa_hmc_mlp_grad_f16_workspace(int C, int64_t N, int in_dim, int hidden, int n_classes);
int ursa_hmc_mlp_grad_f16(const float *theta, int64_t ld, int C, const float *x, const int64_t *y, int64_t N,
                          int in_dim, int hidden, int n_classes, float *grad, float *ce,
                          void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * K3  sample-batched BMA forward, PreResNet (BasicBlock, depth = 6n+2 < 44; models/preresnet.py:90-151)
 *     bank: [S, ld_bank] parameters; bufbank: [S, ld_buf] BatchNorm running stats in named_buffers()
 *     order with the int64 num_batches_tracked entries dropped (mean, var per BN layer);
 *     x: [N, 3, 32, 32] NCHW.  Eval-mode BN (eps 1e-5) + ReLU are folded into the consuming conv.
 *     URSA_ALGO_TCGEN05_FUSED_F16 (the product path): conv1 + stage 1, stage 2 and stage 3 each run as ONE persistent
 *     tcgen05 kernel with the activations resident in shared memory (inputs arrive as TMA-copied plane images, outputs
 *     leave by TMA tensor stores), the two stride-2 convs with their 1x1 shortcuts as one FP16-split implicit GEMM each,
 *     and -- when logits_out is NULL -- the head kernel ends in the softmax-average / entropy epilogue (no logits round
 *     trip).  Other algos write logits to the workspace and call the ursa_bma_accumulate kernel.
 * ---------------------------------------------------------------------- */
size_t ursa_bma_preresnet_workspace(int S, int64_t N, int depth, int C, int algo);
int ursa_bma_preresnet_forward(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf,
                               int S, const float *x, int64_t N, int depth, int C,
                               float *proba_sum, float *entropy_sum, float *logits_out, double gamma,
                               void *workspace, size_t workspace_bytes, int algo, void *stream);

/* ------------------------------------------------------------------------
 * K3  BMA forward, WideResNet (WRN-depth-widen, WideBasic blocks, biased convs, stride on conv2;
 *     models/wideresnet.py:78-120 -- BASELINE.json configs[2] is WRN-28-10 with C = 100).
 *     depth = 6n+4 (n <= 8), widen even in [2, 16].  bank / bufbank / x as for the PreResNet entry point.
 *     One posterior sample at a time over chunks of images; every 3x3 conv (and the 1x1 shortcut convs,
 *     folded into conv2's K loop) runs as a persistent tcgen05 implicit GEMM; the head kernel (BN + ReLU +
 *     pool + linear) ends in the softmax-average / entropy epilogue.  algo: URSA_ALGO_TCGEN05_F16 (2xFP16-split
 *     operands, the product path: 1.3x the 3xTF32 engine and closer to an fp64 forward; fp16's range) or
 *     URSA_ALGO_TCGEN05 (3xTF32, fp32's range); the workspace query returns 0 for an unsupported shape.
 * ---------------------------------------------------------------------- */
size_t ursa_bma_wrn_workspace(int S, int64_t N, int depth, int widen, int C, int algo);
int ursa_bma_wrn_forward(const float *bank, int64_t ld_bank, const float *bufbank, int64_t ld_buf,
                         int S, const float *x, int64_t N, int depth, int widen, int C,
                         float *proba_sum, float *entropy_sum, float *logits_out, double gamma,
                         void *workspace, size_t workspace_bytes, int algo, void *stream);

/* ------------------------------------------------------------------------
 * K3b BatchNorm re-estimation for one posterior sample (replaces util.bn_update, util.py:212-247, called once per SWAG
 *     sample at inference/swag.py:123-124): ONE train-mode pass of the WideResNet over x [N, 3, 32, 32] in batches of
 *     `batch` images (the last may be ragged): every BatchNorm normalises with its batch statistics, the running
 *     statistics are reset and re-estimated with the cumulative momentum b / (n + b) (unbiased running variance, as
 *     PyTorch).  bank_row: [D] parameters; buf_row: [nb] running statistics, overwritten.  Same tcgen05 conv kernel as
 *     ursa_bma_wrn_forward with a statistics epilogue (fp64 sums); batch must be even, 2..512.
 * ---------------------------------------------------------------------- */
size_t ursa_wrn_bn_update_workspace(int64_t N, int batch, int depth, int widen, int C);
int ursa_wrn_bn_update(const float *bank_row, float *buf_row, const float *x, int64_t N, int batch,
                       int depth, int widen, int C, void *workspace, size_t workspace_bytes, void *stream);
/* the same with the engine chosen by the caller: URSA_ALGO_TCGEN05 (3xTF32, = ursa_wrn_bn_update) or URSA_ALGO_TCGEN05_F16
 * (2xFP16-split conv kernel: faster and closer to an fp64 pass; activations beyond fp16's range make the statistics
 * non-finite -- callers then repeat the pass on URSA_ALGO_TCGEN05) */
int ursa_wrn_bn_update_algo(const float *bank_row, float *buf_row, const float *x, int64_t N, int batch,
                            int depth, int widen, int C, void *workspace, size_t workspace_bytes, int algo, void *stream);

/* The same for PreResNets (the north-star model), SAMPLE-BATCHED: S posterior samples (rows of bank / bufbank) take the
 * train-mode pass together, 8 per launch.  Layer by layer on the 3xTF32 tcgen05 conv kernel with a statistics epilogue
 * (batch statistics need the whole batch's conv output, so the fused stage kernels cannot be used): per conv
 * conv -> finalize (this batch's (a, b), running statistics) -> apply.  bufbank rows [S, ld_buf] are overwritten
 * (running_mean, running_var of every BatchNorm in named_buffers() order).  depth = 6n+2 in 8..38; batch even, 2..512. */
size_t ursa_preresnet_bn_update_workspace(int S, int64_t N, int batch, int depth, int C);
int ursa_preresnet_bn_update(const float *bank, int64_t ld_bank, float *bufbank, int64_t ld_buf, int S,
                             const float *x, int64_t N, int batch, int depth, int C,
                             void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * K5  chain-batched HMC  (replaces the call into hamiltorch.sample_model made by HMC.sample,
 *     inference/hmc.py:62-85.  hamiltorch is a third-party dependency that is NOT vendored in the reference
 *     (unpinned git HEAD, util.py:11) -- these entry points follow its published leapfrog / Metropolis algorithm as
 *     restated in oracle/restate.py::hmc_*; parity is unpinned.)
 *   State: theta, r, saved are [C, ld] fp32 (C independent chains, ld % 4 == 0); elementwise entry points take the
 *   batch as n = C*ld flat elements because step size, mass and prior precision are shared (hmc.py:64-75).
 *
 *   ursa_hmc_momentum   r = sqrt_mass * z          z ~ N(0,1): `noise` (parity mode) or Philox as in K1
 *   ursa_hmc_leapfrog   grad_logp = -(tau_out * g_nll + tau * theta)      g_nll = d/dtheta sum_i CE_i (autograd)
 *                       r += kick * grad_logp ; if drift != 0: theta += drift * r ; snapshot = theta (nullable)
 *                       (kick = eps/2 for the first and last half steps, eps between; drift = eps * inv_mass, 0 for
 *                        the closing half kick)                             20 B/param
 *   ursa_hmc_energy     sum_theta2[c] = sum_d theta[c,d]^2, sum_r2[c] = sum_d r[c,d]^2 in fp64, fixed order; the
 *                       caller forms H = tau_out*CE_sum + tau/2*sum_theta2 + D/2*log(2*pi/tau) + inv_mass/2*sum_r2
 *   ursa_hmc_accept     accept[c] = isfinite(H) && log u[c] <= min(0, h_old[c] - h_new[c]);  u from `logu` (parity)
 *                       or Philox (one block per chain: chain_offset + c, `step`).  Accepted chains commit
 *                       theta -> saved (and keep_src -> keep_dst when given), rejected chains restore
 *                       saved -> theta; `out` (nullable, [C, ld_out]) receives the resulting state
 *                       (keep_dst's when given, else theta's).
 * ---------------------------------------------------------------------- */
int ursa_hmc_momentum(float *r, const float *noise, int64_t n, float sqrt_mass, uint64_t seed, uint64_t step,
                      uint64_t elem_offset, void *stream);
int ursa_hmc_leapfrog(float *theta, float *r, const float *g_nll, float *snapshot, int64_t n, float kick,
                      float drift, float tau, float tau_out, void *stream);
size_t ursa_hmc_energy_workspace(int64_t C, int64_t D);
int ursa_hmc_energy(const float *theta, const float *r, int64_t C, int64_t D, int64_t ld, double *sum_theta2,
                    double *sum_r2, void *workspace, size_t workspace_bytes, void *stream);
int ursa_hmc_accept(float *theta, float *saved, float *keep_dst, const float *keep_src, float *out, int64_t ld_out,
                    int64_t C, int64_t ld, const double *h_old, const double *h_new, const float *logu,
                    int32_t *accept, uint64_t seed, uint64_t step, uint64_t chain_offset, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* URSA_B200_H */
