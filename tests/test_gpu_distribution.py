"""Distributional validation of the samplers on a Gaussian target, THROUGH THE CLASS API (north star: "sampler
distributional correctness is validated on a Gaussian target: chain mean and covariance within stated Monte Carlo
error").

Target: Bayesian linear regression, theta in R^d with correlated design columns, full-batch gradients (no minibatch
noise), Gaussian prior.  With the reference's conventions (mean loss, weight decay lambda / N, noise
sqrt(2 (1 - mu) lr T) / N; optim_sghmc.py:47-48,63-64) the stationary law of SGLD / SGHMC / cSGHMC is

    N(theta*, T * [N (2 X^T X + lambda I)]^-1),   theta* = (2 X^T X + lambda I)^-1 2 X^T y,

up to an O(lr) discretisation bias (checked by a CPU simulation of the reference's update when this test was written).
One sampler object drives CHAINS independent chains at once: the "model" holds theta as a [CHAINS, d] parameter and the
loss is the SUM over chains of each chain's mean squared error, so every chain sees exactly its own gradient while the
flat-buffer K1 update, the Philox noise, the momentum branch ("noise stored in v") and the class's schedules and gates
are the production code path.  Monte-Carlo error: with n = CHAINS x kept samples (thinned samples of one chain are
positively correlated, so n_eff >= CHAINS is what the bounds below assume) the standard error of a mean component is
sigma / sqrt(n_eff) and of a covariance entry ~ sqrt((S_ii S_jj + S_ij^2) / n_eff); bounds are 5 standard errors plus
the stated discretisation allowance.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
D_DIM, N_DATA, CHAINS = 4, 32, 4096


class ChainBatchedRegression(torch.nn.Module):
    def __init__(self, chains, d):
        super().__init__()
        self.theta = torch.nn.Parameter(torch.zeros(chains, d))

    def reset_parameters(self):
        with torch.no_grad():
            self.theta.zero_()

    def forward(self, x):                       # [B, d] -> [B, chains]
        return x @ self.theta.t()


def _sum_of_chain_mse(pred, y):                 # sum over chains of the chain's MEAN squared error
    return ((pred - y[:, None]) ** 2).mean(0).sum()


def _problem():
    rng = np.random.RandomState(0)
    X = rng.randn(N_DATA, D_DIM)
    X[:, 1] = 0.8 * X[:, 0] + 0.6 * X[:, 1]                 # correlated columns -> correlated posterior
    X[:, 3] = 0.5 * X[:, 2] - 0.5 * X[:, 0] + 0.7 * X[:, 3]
    y = X @ np.array([0.5, -1.0, 0.3, 0.8]) + 0.1 * rng.randn(N_DATA)
    prior_std = 2.0
    lam = 1.0 / prior_std ** 2
    A = 2 * X.T @ X + lam * np.eye(D_DIM)
    mean = np.linalg.solve(A, 2 * X.T @ y)
    cov_T1 = np.linalg.inv(N_DATA * A)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.from_numpy(X).float(), torch.from_numpy(y).float()),
                                         batch_size=N_DATA, shuffle=False)
    return loader, prior_std, mean, cov_T1


def _check(samples, mean, cov, n_eff, disc=0.04):
    """samples [n, d] float64; disc = allowed relative O(lr) bias of the covariance."""
    m = samples.mean(0)
    S = np.cov(samples.T)
    se_mean = np.sqrt(np.diag(cov) / n_eff)
    assert np.all(np.abs(m - mean) <= 5 * se_mean + 1e-4), (m - mean, se_mean)
    for i in range(D_DIM):
        for j in range(D_DIM):
            se = np.sqrt((cov[i, i] * cov[j, j] + cov[i, j] ** 2) / n_eff)
            tol = 5 * se + disc * np.sqrt(cov[i, i] * cov[j, j])
            assert abs(S[i, j] - cov[i, j]) <= tol, (i, j, S[i, j], cov[i, j], tol)
    # the correlation structure is really there (not a diagonal law passing by accident)
    corr = cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    assert np.abs(corr - np.eye(D_DIM)).max() > 0.3
    return S


def _stack(handles):
    return np.concatenate([h.theta.detach().double().cpu().numpy() for h in handles], 0)


@pytest.mark.parametrize("alpha,lr,temperature", [(0.3, 0.05, 1.0), (1.0, 0.15, 1.0), (0.3, 0.05, float(N_DATA))])
def test_sghmc_class_samples_the_gaussian_target(alpha, lr, temperature):
    """SGHMC (momentum branch: the noise is stored in v) and its alpha = 1 limit (SGLD arithmetic) through
    ``inference.SGHMC``; temperature = N recovers the Bayesian posterior N(theta*, (2 X^T X + lambda I)^-1)."""
    import ursabench_b200 as U
    loader, prior_std, mean, cov1 = _problem()
    torch.manual_seed(11)
    hyp = {"lr": lr, "prior_std": prior_std, "num_samples": 40, "alpha": alpha, "burn_in_epochs": 400,
           "temperature": temperature}
    inf = U.inference.SGHMC(hyp, ChainBatchedRegression(CHAINS, D_DIM), loader, device=DEV)
    inf.loss_criterion = _sum_of_chain_mse
    out = inf.sample()
    assert len(out) == 40
    # one step per epoch (full batch); the cosine schedule has annealed lr to ~2 % of its start when sampling begins, so
    # the kept samples carry almost no discretisation bias but are strongly correlated along a chain: n_eff = CHAINS
    kept = _stack(out[:4])
    S = _check(kept, mean, temperature * cov1, n_eff=CHAINS)
    assert S.shape == (D_DIM, D_DIM)


def test_csghmc_class_samples_the_gaussian_target():
    """cSGHMC: exploration (no noise) then sampling inside every cycle, per-iteration cosine step size, samples taken
    at the end of the cycles -- the production schedule and gates, momentum branch."""
    import ursabench_b200 as U
    loader, prior_std, mean, cov1 = _problem()
    torch.manual_seed(12)
    # noise is on for the last burn_in_epochs + num_samples_per_cycle = 352 epochs of each 400-epoch cycle (csghmc.py:89-90)
    hyp = {"lr_0": 0.08, "prior_std": prior_std, "num_samples_per_cycle": 2, "cycle_length": 400, "burn_in_epochs": 350,
           "num_cycles": 2, "alpha": 0.3}
    inf = U.inference.cSGHMC(hyp, ChainBatchedRegression(CHAINS, D_DIM), loader, device=DEV)
    inf.loss_criterion = _sum_of_chain_mse
    out = inf.sample()
    assert len(out) == 4
    # samples come from the low-lr tail of each cycle, where the discretisation bias vanishes but the chain has had
    # ~350 noisy epochs to equilibrate; allow 6 % on the covariance scale
    _check(_stack(out), mean, cov1, n_eff=CHAINS, disc=0.06)


def test_swag_low_rank_draws_follow_the_documented_gaussian_law():
    """Row a7, low-rank branch (swag.py:88-97): the reference's own branch raises, so no golden can pin it; the
    formula is pinned to the restatement elsewhere (test_k2_draw_matches_oracle) and the LAW is checked here.  12 000 draws
    of the production path (30 draws per launch, Philox z1 generated in the kernel, z2 from torch's device generator the
    way ``SWAG._draw_into_bank`` does) must have mean = SWA mean and covariance diag(var) + R^T R / (max_rank - 1) within
    6 standard errors per entry, and draws of the same launch must be uncorrelated."""
    from ursabench_b200 import _C
    D, K, S, LAUNCHES, max_rank = 301, 6, 30, 400, 20                     # D: two column tiles + a ragged tail, ld = 304
    rng = np.random.RandomState(3)
    ld = (D + 3) // 4 * 4
    mean = rng.randn(D)
    var = 0.05 + rng.rand(D)
    ring = rng.randn(K, D) * np.linspace(3.0, 0.5, K)[:, None]

    def padded(a):
        t = torch.zeros(a.shape[:-1] + (ld,), device=DEV)
        t[..., :D] = torch.from_numpy(a).float().to(DEV)
        return t

    mean_d, var_d, ring_d = padded(mean), padded(var), padded(ring)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(1234)
    out = torch.empty(LAUNCHES, S, ld, device=DEV)
    for step in range(LAUNCHES):
        z2 = torch.randn(S, K, device=DEV, generator=gen)
        _C.swag_draw(out[step], mean_d, var_d, D, ring=ring_d, z2=z2, rank_div=float((max_rank - 1) ** 0.5), seed=77, step=step)
    x = out[..., :D].double().cpu().numpy()                               # [LAUNCHES, S, D]
    mean32, var32, ring32 = (t[..., :D].double().cpu().numpy() for t in (mean_d, var_d, ring_d))
    cov = np.diag(var32) + ring32.T @ ring32 / (max_rank - 1)
    flat = x.reshape(-1, D)
    n = flat.shape[0]
    se_mean = np.sqrt(np.diag(cov) / n)
    assert np.all(np.abs(flat.mean(0) - mean32) <= 6 * se_mean), np.abs((flat.mean(0) - mean32) / se_mean).max()
    emp = np.cov(flat.T)
    se_cov = np.sqrt((np.outer(np.diag(cov), np.diag(cov)) + cov ** 2) / n)
    z = np.abs(emp - cov) / se_cov
    assert z.max() <= 6.0, (z.max(), np.unravel_index(z.argmax(), z.shape))
    # not a diagonal law passing by accident: the low-rank part carries real correlations
    corr = cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    assert np.abs(corr - np.eye(D)).max() > 0.3
    # draws s != s' of one launch share the ring pass and the Philox blocks (six normals per block): they must still be
    # independent.  Standardised residuals of draw 0 vs draw 1, and of two draws served by the SAME Philox block (0 and 5)
    res = (x - mean32) / np.sqrt(np.diag(cov))
    for a, b in ((0, 1), (0, 5), (6, 29)):
        per_launch = (res[:, a, :] * res[:, b, :]).mean(1)                # one number per launch, iid over launches, mean 0
        t = per_launch.mean() / (per_launch.std(ddof=1) / np.sqrt(LAUNCHES))
        assert abs(t) <= 5.0, (a, b, t)
