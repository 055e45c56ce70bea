"""numpy restatement of the reference's hot-path arithmetic (CPU oracle).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line it restates; paths are relative to
``/root/reference/URSABench/``.  All state is fp32 unless noted.  Fused
multiply-adds are emulated through float64 (the product of two fp32 values is
exact in fp64), because torch's CPU ``add(alpha=...)`` kernel is an FMA
(verified in this container: 0 mismatches / 1e6 against the fp64 emulation,
3076 / 1e6 against separate multiply + add).
"""
import math

import numpy as np

F32 = np.float32


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def fma32(a, b, c):
    """round_to_f32(a*b + c) with a single rounding (a, b, c fp32)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


# --------------------------------------------------------------------------
# a1  optimSGHMC.step            inference/optim_sghmc.py:30-68
# --------------------------------------------------------------------------
def sgmcmc_step(p, g, v, z, lr, momentum, weight_decay, n_train, first_step, add_noise):
    """One optimSGHMC update over a flat fp32 vector.

    p, g      parameters / gradients                        (optim_sghmc.py:43-46)
    v         momentum buffer, ignored when ``first_step``  (:51-52)
    z         standard-normal draws used when ``add_noise`` (:63-64)
    returns   (p_new, v_new)  -- v_new is None when momentum == 0
    """
    p = _f32(p)
    d = _f32(g)
    if weight_decay != 0:                                     # :47-48  d = g + (wd/N) p
        d = fma32(F32(weight_decay / n_train), p, d)
    neg_lr = F32(-lr)
    if momentum != 0:
        buf = d.copy() if first_step else _f32(v)             # :51-52 clone(d) on first step
        buf = fma32(neg_lr, d, buf * F32(momentum))           # :53/:56 buf.mul_(mu).add_(d, alpha=-lr)
        u = buf                                               # :60
    else:
        u = d * neg_lr                                        # :62
    if add_noise:                                             # :63-64  (z * s) / N, then add
        s = F32(math.sqrt(2 * (1 - momentum) * lr))
        u = u + (_f32(z) * s) / F32(n_train)
    p_new = p + u                                             # :65
    return p_new, (u if momentum != 0 else None)              # :66-67 noise is stored in the momentum


# --------------------------------------------------------------------------
# a3  cSGHMC schedule and gates  inference/csghmc.py:30-31,42-44,64-72,89-93,106
# --------------------------------------------------------------------------
def csghmc_num_batch(n_train, batch_size):
    return max(1, n_train / batch_size + 1)                   # :30-32 (float, over-counts by >=1)


def csghmc_lr(lr_0, epoch, batch_idx, num_batch, cycle_length, num_cycles):
    total_iterations = cycle_length * num_cycles * num_batch  # :42-44
    per_cycle = total_iterations // num_cycles                # float floor-division (:67)
    r = epoch * num_batch + batch_idx                         # :65
    inner = np.pi * (r % per_cycle)
    inner /= per_cycle
    return 0.5 * (np.cos(inner) + 1) * lr_0                   # :69-70


def csghmc_noise_gate(epochs_run, cycle_length, burn_in_epochs, num_samples_per_cycle):
    return (epochs_run % cycle_length) + 1 > (cycle_length - burn_in_epochs - num_samples_per_cycle)  # :89-90


def csghmc_sample_gate(epochs_run_after_increment, cycle_length, num_samples_per_cycle):
    return ((epochs_run_after_increment - 1) % cycle_length) >= (cycle_length - num_samples_per_cycle)  # :106


# --------------------------------------------------------------------------
# a8  SWA lr schedule            inference/swa.py:92-101
# --------------------------------------------------------------------------
def swa_schedule(epoch, burn_in_epochs, lr_init, swag_lr):
    t = epoch / burn_in_epochs
    ratio = swag_lr / lr_init
    if t <= 0.5:
        factor = 1.0
    elif t <= 0.9:
        factor = 1.0 - (1.0 - ratio) * (t - 0.5) / 0.4
    else:
        factor = ratio
    return lr_init * factor


# --------------------------------------------------------------------------
# a4  SWA._collect_model         inference/swa.py:79-90
# a5  CovarianceSpace ring       inference/subspaces.py:85-92
# a6  _get_mean_and_variance     inference/swa.py:106-108
# --------------------------------------------------------------------------
def swag_collect(w, mean, sq_mean, n):
    """n = num_models_collected as seen by _collect_model (a python int)."""
    w = _f32(w)
    keep = F32(n / (n + 1.0))                                 # :83 python float -> fp32 scalar
    denom = F32(n + 1.0)
    mean_new = _f32(mean) * keep + w / denom                  # :83-84 mul_ then add_ (alpha=1: plain add)
    sq_new = _f32(sq_mean) * keep + (w * w) / denom           # :87-88
    dev = w - mean_new                                        # :89
    return mean_new, sq_new, dev


def ring_push(ring, dev, max_rank=20):
    """ring: [r, D] array of the most recent deviations, oldest first."""
    ring = np.asarray(ring, np.float32).reshape(-1, dev.shape[0])
    if ring.shape[0] + 1 > max_rank:                          # subspaces.py:86-87
        ring = ring[1:]
    return np.concatenate([ring, _f32(dev)[None, :]], axis=0)  # :88


def ring_get_space(ring):
    return _f32(ring) / F32((ring.shape[0] - 1) ** 0.5)       # subspaces.py:92


def swag_variance(mean, sq_mean, var_clamp=1e-30):
    mean = _f32(mean)
    return np.maximum(_f32(sq_mean) - mean * mean, F32(var_clamp))  # swa.py:107


# --------------------------------------------------------------------------
# a7  SWAG draw                  inference/swag.py:85-97
# The reference discards its draw (:98) and its low-rank branch crashes (:90);
# this is the intended formula and is the spec for the kernel.
# --------------------------------------------------------------------------
def swag_draw(mean, var, z1, ring=None, z2=None, max_rank=20):
    """z1: [S, D]; ring: [K, D]; z2: [S, K].  Returns [S, D]."""
    mean = _f32(mean)
    std = np.sqrt(_f32(var))
    var_sample = std[None, :] * _f32(z1)                      # :86 / :88-89
    if ring is None:
        return mean[None, :] + var_sample                     # :86 Normal(mean, sqrt(var))
    cov = (_f32(z2).astype(np.float64) @ _f32(ring).astype(np.float64)).astype(np.float32)  # :90-94 D^T z2
    cov = cov / F32((max_rank - 1) ** 0.5)                    # :95 uses max_rank, not the current rank
    return mean[None, :] + (var_sample + cov)                 # :96-97


# --------------------------------------------------------------------------
# a14 util.central_smoothing / compute_predictive_entropy   util.py:126-144
# a10 Prediction.update_statistics                          tasks/prediction.py:52-75
# --------------------------------------------------------------------------
def softmax_rows(logits):
    """log_softmax(x).exp() in fp32 (tasks/prediction.py:60)."""
    x = _f32(logits)
    m = x.max(axis=-1, keepdims=True)
    s = x - m
    lse = np.log(np.exp(s).sum(axis=-1, keepdims=True, dtype=np.float32))
    return np.exp(s - lse).astype(np.float32)


def central_smoothing(proba, gamma=1e-4):
    proba = _f32(proba)
    return (F32(1 - gamma) * proba + F32(gamma * 1 / proba.shape[-1])).astype(np.float32)  # util.py:134


def predictive_entropy(proba):
    proba = _f32(proba)
    return -(proba * np.log(proba)).sum(axis=-1, dtype=np.float32)  # util.py:144


def bma_accumulate(logits, proba_sum=None, entropy_sum=None):
    """logits: [S, N, C].  Accumulates in list order like prediction.py:56-64."""
    logits = _f32(logits)
    S, N, C = logits.shape
    proba_sum = np.zeros((N, C), np.float32) if proba_sum is None else _f32(proba_sum).copy()
    entropy_sum = np.zeros((N,), np.float32) if entropy_sum is None else _f32(entropy_sum).copy()
    for s in range(S):
        p = softmax_rows(logits[s])
        proba_sum += p                                        # :60
        entropy_sum += predictive_entropy(central_smoothing(p))  # :61-63
    return proba_sum, entropy_sum


# --------------------------------------------------------------------------
# a11-a13 metrics                tasks/prediction.py:79-102, 152-194
# --------------------------------------------------------------------------
def bma_counters(proba_sum, num_samples, targets, n_bins=15, gamma=1e-4):
    """Integer / fp64 counters that the CUDA metric kernel must reproduce.

    Returns a dict with
      correct (int), bin_count[15] (int64), bin_correct[15] (int64)  -- exact-match quantities
      bin_conf_sum[15], nll_sum, brier_sum (float64)
      pred[N] (int64), conf[N] (float32)
    """
    pbar = _f32(proba_sum) / F32(num_samples)                 # prediction.py:82 (fp32 true division)
    targets = np.asarray(targets, np.int64)
    N, C = pbar.shape
    pred = pbar.argmax(axis=1)                                # first maximum (:83, :164)
    conf = pbar.max(axis=1)
    acc = pred == targets
    bounds = np.linspace(0, 1, n_bins + 1)                    # float64 (:160)
    conf64 = conf.astype(np.float64)
    bin_count = np.zeros(n_bins, np.int64)
    bin_correct = np.zeros(n_bins, np.int64)
    bin_conf_sum = np.zeros(n_bins, np.float64)
    for b in range(n_bins):                                   # (lo, hi]  (:169-170)
        inb = np.logical_and(conf64 > bounds[b], conf64 <= bounds[b + 1])
        bin_count[b] = inb.sum()
        bin_correct[b] = acc[inb].sum()
        bin_conf_sum[b] = conf64[inb].sum()
    smoothed = central_smoothing(pbar, gamma)                 # :88-90
    nll_sum = float(-np.log(smoothed[np.arange(N), targets].astype(np.float64)).sum())
    onehot = np.zeros(pbar.shape)                             # float64 (:192)
    onehot[np.arange(N), targets] = 1.0
    brier_sum = float(((pbar - onehot) ** 2).sum())
    return dict(correct=int(acc.sum()), bin_count=bin_count, bin_correct=bin_correct,
                bin_conf_sum=bin_conf_sum, nll_sum=nll_sum, brier_sum=brier_sum,
                pred=pred.astype(np.int64), conf=conf, n=N)


def metrics_from_counters(c):
    """error_rate / nll / ll / brier_score / ece from the counters (fp64 host math)."""
    n = c["n"]
    ece = 0.0
    for b in range(len(c["bin_count"])):
        cnt = int(c["bin_count"][b])
        if cnt > 0:                                           # prediction.py:172
            delta = c["bin_conf_sum"][b] / cnt - c["bin_correct"][b] / cnt
            ece += abs(delta) * (cnt / n)                     # :177
    nll = c["nll_sum"] / n
    return dict(error_rate=1 - c["correct"] / n, nll=nll, ll=-nll,
                brier_score=c["brier_sum"] / n, ece=ece)


def get_ece(preds, targets, n_bins=15):
    """Line-for-line semantics of _get_ece incl. float32 means (prediction.py:152-182)."""
    bounds = np.linspace(0, 1, n_bins + 1)
    conf = np.max(preds, 1)
    pred = np.argmax(preds, 1)
    acc = pred == targets
    total = 0.0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        inb = np.logical_and(conf > lo, conf <= hi)
        prop = np.mean(inb)
        if prop > 0:
            total += np.abs(np.mean(conf[inb]) - np.mean(acc[inb])) * prop
    return total


def get_brier(preds, targets):
    onehot = np.zeros(preds.shape)                            # prediction.py:192-194
    onehot[np.arange(len(targets)), targets] = 1.0
    return np.mean(np.sum((preds - onehot) ** 2, axis=1))


# --------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller: the in-register noise source of the CUDA kernels
# (not in the reference, which calls torch.randn_like at optim_sghmc.py:64;
# this restates Salmon et al. 2011 so the device stream can be checked on CPU).
# --------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: (k0, k1).  Returns uint32 [..., 4]."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0 = np.uint32(key[0])
    k1 = np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _PHILOX_W0)
            k1 = np.uint32(k1 + _PHILOX_W1)
    return np.stack([c0, c1, c2, c3], axis=-1)


def philox_normals(n, seed, step, elem_offset=0):
    """The Gaussian stream of ``ursa_sgmcmc_step`` in Philox mode for flat
    element indices [elem_offset, elem_offset + n): element i uses lane i % 4 of
    the block with counter (i//4 lo, i//4 hi, step lo, step hi), key = seed."""
    idx = np.arange(elem_offset, elem_offset + n, dtype=np.uint64)
    blk = idx >> np.uint64(2)
    lane = (idx & np.uint64(3)).astype(np.int64)
    ublk, inv = np.unique(blk, return_inverse=True)
    ctr = np.zeros((ublk.shape[0], 4), np.uint32)
    ctr[:, 0] = (ublk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (ublk >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32((step >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    z = box_muller(r)                                         # [nblk, 4]
    return z[inv, lane].astype(np.float32), r


def box_muller(r):
    """r: uint32 [..., 4] -> normals [..., 4]: (r0,r1)->(z0,z1), (r2,r3)->(z2,z3)."""
    r = np.asarray(r, np.uint32).astype(np.float64)
    two_m32 = 2.0 ** -32
    u1a = r[..., 0] * two_m32 + 2.0 ** -33                    # (0, 1)
    u2a = r[..., 1] * two_m32 + 2.0 ** -33
    u1b = r[..., 2] * two_m32 + 2.0 ** -33
    u2b = r[..., 3] * two_m32 + 2.0 ** -33
    ra = np.sqrt(-2.0 * np.log(u1a))
    rb = np.sqrt(-2.0 * np.log(u1b))
    ta = 2.0 * np.pi * u2a
    tb = 2.0 * np.pi * u2b
    return np.stack([ra * np.cos(ta), ra * np.sin(ta), rb * np.cos(tb), rb * np.sin(tb)], axis=-1)


def draw_normals(S, D, seed, step):
    """The N(0, 1) stream of ``ursa_swag_draw`` in Philox mode (csrc/common.cuh::box_muller6): element (s, d) is normal
    s % 6 of the block with counter ((s // 6) * D + d, step), key = seed; the block's 128 bits x:y:z:w are cut from the top
    into three (24-bit radius uniform a, 18-bit angle t) pairs.  Returns float64 [S, D]."""
    G = (S + 5) // 6
    blk = (np.arange(G, dtype=np.uint64)[:, None] * np.uint64(D) + np.arange(D, dtype=np.uint64)[None, :]).reshape(-1)
    ctr = np.zeros((blk.shape[0], 4), np.uint32)
    ctr[:, 0] = (blk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (blk >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32((step >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.uint64)
    bits = (r[:, 0] << np.uint64(32) | r[:, 1]), (r[:, 2] << np.uint64(32) | r[:, 3])      # (x:y, z:w)

    def field(top, width):                                    # `width` bits whose MSB is bit `top` of the 128-bit string (127 = MSB of x)
        lo = top - width + 1
        if lo >= 64:
            v = bits[0] >> np.uint64(lo - 64)
        elif top < 64:
            v = bits[1] >> np.uint64(lo)
        else:
            v = (bits[0] << np.uint64(64 - lo)) | (bits[1] >> np.uint64(lo))
        return (v & np.uint64((1 << width) - 1)).astype(np.float64)

    z = np.empty((G, 6, D))
    top = 127
    for j in range(3):
        a, t = field(top, 24), field(top - 24, 18)
        top -= 42
        u = ((a + 0.5) * 2.0 ** -24).astype(np.float32).astype(np.float64)       # the device rounds u to fp32 (a = 2^24 - 1 -> 1.0)
        rad = np.sqrt(-2.0 * np.log(u))
        th = 2.0 * np.pi * (t + 0.5) * 2.0 ** -18
        z[:, 2 * j] = (rad * np.cos(th)).reshape(G, D)
        z[:, 2 * j + 1] = (rad * np.sin(th)).reshape(G, D)
    return z.reshape(6 * G, D)[:S]


# --------------------------------------------------------------------------
# a15  HMC                        inference/hmc.py:62-85  ->  hamiltorch.sample_model
#
# PARITY UNPINNED.  The arithmetic lives in the third-party package `hamiltorch`
# (git+https://github.com/AdamCobb/hamiltorch, unpinned HEAD; the demo notebook's install log shows
# hamiltorch-0.4.0.dev1) which is neither vendored under /root/reference nor installed here.  What follows restates
# its published algorithm (hamiltorch/samplers.py: gibbs, leapfrog, hamiltonian, sample with Sampler.HMC and a
# diagonal `inv_mass` vector; hamiltorch/samplers.py::define_model_log_prob with Normal(0, tau^-1/2) priors and
# `multi_class_linear_output`), anchored on the reference's own call site: hmc.py:64-75 passes one `tau` for every
# tensor, `inv_mass = ones/mass`, `burn=-1`, `tau_out=1`, and hmc.py:78-81 thins the returned list by L.
# --------------------------------------------------------------------------
def hmc_momentum(z, mass):
    """gibbs(): Normal(0, mass**0.5).sample() = sqrt(mass) * z."""
    return _f32(z) * F32(math.sqrt(mass))


def hmc_leapfrog_update(theta, r, g_nll, kick, drift, tau, tau_out=1.0):
    """One kick (+ drift) of the leapfrog integrator on flat fp32 vectors.

    grad log p(theta) = -(tau_out * g_nll + tau * theta)   (ll = -tau_out * sum CE ; prior Normal(0, tau^-1/2))
    momentum += kick * grad ; params += drift * momentum    (hamiltorch leapfrog: kick = eps/2 before the loop and
    as the closing correction, eps inside; drift = eps * inv_mass, 0 for the closing half kick)
    """
    theta, r, g = _f32(theta), _f32(r), _f32(g_nll)
    glp = -fma32(F32(tau), theta, F32(tau_out) * g)
    r = r + F32(kick) * glp
    if drift != 0:
        theta = theta + F32(drift) * r
    return theta, r


def hmc_energy_sums(theta, r):
    """(sum theta^2, sum r^2) in float64 (fp32 products are exact in fp64)."""
    t, q = np.asarray(theta, np.float64), np.asarray(r, np.float64)
    return float((t * t).sum()), float((q * q).sum())


def hmc_hamiltonian(ce_sum, sum_theta2, sum_r2, D, tau, inv_mass, tau_out=1.0):
    """hamiltonian(): H = -log p + 0.5 * r . (inv_mass * r);  log p = -tau_out*CE_sum + sum_d log N(theta_d; 0, tau^-1/2)."""
    log_prior = -0.5 * tau * sum_theta2 - 0.5 * D * math.log(2.0 * math.pi / tau)
    return tau_out * ce_sum - log_prior + 0.5 * inv_mass * sum_r2


def hmc_accept(h_old, h_new, log_u):
    """sample(): rho = min(0, h_old - h_new); accept iff rho >= log u.  Non-finite energies reject (LogProbError)."""
    if not (math.isfinite(h_old) and math.isfinite(h_new)):
        return False
    return min(0.0, h_old - h_new) >= log_u


def hmc_chain(theta0, nll_and_grad, z_list, logu_list, step_size, L, tau, mass, tau_out=1.0):
    """Single-chain hamiltorch.sample(...) loop with injected momentum noise z_list[n] and log-uniforms.

    nll_and_grad(theta) -> (sum CE as float, d/dtheta sum CE as fp32 vector).  Returns (ret, accepts): ``ret`` is
    hamiltorch's returned list -- the initial point followed by the L leapfrog positions of every iteration
    (burn = -1 => every iteration is recorded; a rejected iteration re-appends the previous L entries) -- and
    the per-iteration accept flags.  The reference wrapper then takes ret[burn*L::L] (hmc.py:80).
    """
    inv_mass = 1.0 / mass
    theta = _f32(theta0).copy()
    D = theta.size
    ret = [theta.copy()]
    accepts = []
    for n in range(len(z_list)):
        r = hmc_momentum(z_list[n], mass)
        ce, g = nll_and_grad(theta)
        h_old = hmc_hamiltonian(ce, *hmc_energy_sums(theta, r), D, tau, inv_mass, tau_out)
        cur, traj = theta, []
        for step in range(L):
            kick = 0.5 * step_size if step == 0 else step_size
            cur, r = hmc_leapfrog_update(cur, r, g, kick, step_size * inv_mass, tau, tau_out)
            ce, g = nll_and_grad(cur)
            traj.append(cur.copy())
        _, r = hmc_leapfrog_update(cur, r, g, 0.5 * step_size, 0.0, tau, tau_out)
        h_new = hmc_hamiltonian(ce, *hmc_energy_sums(cur, r), D, tau, inv_mass, tau_out)
        ok = hmc_accept(h_old, h_new, logu_list[n])
        accepts.append(ok)
        if ok:
            theta = cur
            ret.extend(traj)
        else:
            ret.extend(ret[-L:])
    return ret, accepts


# --------------------------------------------------------------------------
# SURVEY 8(f).3  PCASpace.get_space (integer pca_rank)      inference/subspaces.py:116-131,154-156
#                SubspaceModel.forward                      inference/projection_model.py:13-14
# The reference runs sklearn's randomized_svd (n_iter = 5, flip_sign -> svd_flip, u-based) on A = ring / sqrt(max(1, rank-1));
# for rank <= max_rank <= 24 rows that converges to the exact SVD, restated here with numpy's.
# --------------------------------------------------------------------------
def pca_space(ring_rows, rank, pca_rank):
    """ring_rows: [r, D] (any row order); returns s[:k, None] * Vt[:k] as float32, k = max(1, min(pca_rank, rank))."""
    A = _f32(ring_rows).astype(np.float64) / (max(1, int(rank) - 1)) ** 0.5          # :121
    U, sv, Vt = np.linalg.svd(A, full_matrices=False)
    idx = np.argmax(np.abs(U), axis=0)                                                # svd_flip(u_based_decision=True)
    signs = np.sign(U[idx, np.arange(U.shape[1])])
    signs[signs == 0] = 1.0
    Vt = Vt * signs[:, None]
    k = max(1, min(int(pca_rank), int(rank)))                                         # :128
    return (sv[:k, None] * Vt[:k]).astype(np.float32)                                 # :156


# --------------------------------------------------------------------------
# SURVEY 8(f).3  PCASpace.get_space, pca_rank = 'mle'        inference/subspaces.py:123-153
# The rank selection calls ``sklearn.decomposition.pca._assess_dimension_`` (subspaces.py:13,143-146), a private function of a
# third-party dependency that is absent here: the four-argument form (spectrum, rank, n_samples, n_features) exists in
# scikit-learn <= 0.22.x (the reference pins no version; 0.23 renamed it to ``_assess_dimension(spectrum, rank, n_samples)``
# and 0.24 removed the ``sklearn.decomposition.pca`` module).  Its published algorithm is the Laplace-approximated evidence of
# Minka, "Automatic choice of dimensionality for PCA" (NIPS 2000), eq. 30, restated below term by term.  Pinned in
# tests/test_pca_space.py against the installed scikit-learn's three-argument ``_assess_dimension`` (same formula with
# n_features = len(spectrum), which is exactly how the reference calls it: n_features = min(shape) = len(eigs); the <= 0.22
# exponent of the 2 pi term is (m + rank + 1) / 2, the later one (m + rank) / 2 -- a constant over ranks).
# --------------------------------------------------------------------------
def assess_dimension_sklearn022(spectrum, rank, n_samples, n_features):
    """log p(data | rank) of Minka's PCA model for a spectrum of covariance eigenvalues (descending)."""
    spectrum = np.asarray(spectrum, dtype=np.float64)
    if rank > len(spectrum):
        raise ValueError("The tested rank cannot exceed the rank of the dataset")
    pu = -rank * math.log(2.0)                                              # p(U): area of the Stiefel manifold
    for i in range(rank):
        pu += math.lgamma((n_features - i) / 2.0) - math.log(math.pi) * (n_features - i) / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        pl = -float(np.sum(np.log(spectrum[:rank]))) * n_samples / 2.0      # retained eigenvalues
    if rank == n_features:
        pv, v = 0.0, 1.0
    else:
        v = float(np.sum(spectrum[rank:])) / (n_features - rank)            # ML noise variance
        with np.errstate(divide="ignore", invalid="ignore"):
            pv = -float(np.log(v)) * n_samples * (n_features - rank) / 2.0    # numpy's log: v = 0 gives +inf, not an error
    m = n_features * rank - rank * (rank + 1.0) / 2.0                       # degrees of freedom of U
    pp = math.log(2.0 * math.pi) * (m + rank + 1.0) / 2.0
    pa = 0.0                                                                # log |A_Z|, the Hessian determinant
    spectrum_ = spectrum.copy()
    spectrum_[rank:n_features] = v
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(rank):
            for j in range(i + 1, len(spectrum)):
                pa += float(np.log((spectrum[i] - spectrum[j]) * (1.0 / spectrum_[j] - 1.0 / spectrum_[i]))) + math.log(n_samples)
    return pu + pl + pv + pp - pa / 2.0 - rank * math.log(n_samples) / 2.0


def pca_mle_rank(eigs, n_rows, n_cols):
    """The reference's post-selection (subspaces.py:133-151): eigs = s**2 of the [n_rows, n_cols] matrix A; returns
    (ll, corrected_ll, chosen rank = nanargmax(corrected_ll)).  Candidate ranks are 0 .. len(eigs) - 1."""
    eigs = np.asarray(eigs, dtype=np.float64)
    ll = np.zeros(len(eigs))
    correction = np.zeros(len(eigs))
    for rank in range(len(eigs)):
        m = n_cols * rank - rank * (rank + 1) / 2.0                          # :139
        correction[rank] = 0.5 * m * np.log(n_rows)                          # :140
        ll[rank] = assess_dimension_sklearn022(eigs, rank, n_samples=max(n_rows, n_cols), n_features=min(n_rows, n_cols))
    corrected = ll - correction
    return ll, corrected, int(np.nanargmax(corrected))


def pca_space_mle(ring_rows, rank):
    """get_space of PCASpace(pca_rank='mle') (subspaces.py:123-153): full-rank SVD, Minka's criterion, first k components."""
    full = pca_space(ring_rows, rank, rank).astype(np.float64)               # s[:, None] * Vt, all r components
    A = _f32(ring_rows)
    eigs = np.sum(full * full, axis=1)                                       # s**2
    ll, corrected, k = pca_mle_rank(eigs, A.shape[0], A.shape[1])
    return full[:k].astype(np.float32), k, ll, corrected


def subspace_project(mean, cov_factor, t):
    """mean + cov_factor^T t (projection_model.py:14)."""
    return (_f32(mean).astype(np.float64) + _f32(cov_factor).astype(np.float64).T @ _f32(t).astype(np.float64)).astype(np.float32)
