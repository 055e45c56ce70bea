"""GPU tests of the chain-batched HMC path (K5 kernels through the C ABI + the ``HMC`` class).  hamiltorch is a
third-party dependency absent from the reference tree, so parity is against the restatement of its published
algorithm (oracle/restate.py::hmc_*, PARITY UNPINNED) plus distributional checks on a conjugate Gaussian target."""
import math

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


@pytest.fixture(scope="module")
def C():
    from ursabench_b200 import _C
    _C.lib()
    return _C


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [0, 4, 1024, 199_212, 2_000_004])
def test_k5_momentum_external_and_philox(C, n):
    rng = np.random.RandomState(1)
    z = rng.randn(n).astype(np.float32)
    r = torch.empty(n, device=DEV)
    C.hmc_momentum(r, math.sqrt(0.1919), noise=dev(z))
    assert np.array_equal(r.cpu().numpy(), R.hmc_momentum(z, 0.1919))
    if 0 < n <= 199_212:
        C.hmc_momentum(r, math.sqrt(0.1919), seed=5, step=3, elem_offset=8)
        zz, _ = R.philox_normals(n, 5, 3, 8)
        np.testing.assert_allclose(r.cpu().numpy(), zz * np.float32(math.sqrt(0.1919)), rtol=0, atol=2e-5)   # MUFU log2 / sin / cos vs. libm
        # same stream as K1's generator
        ref = torch.empty(n, device=DEV)
        C.philox_normal(ref, 5, 3, 8)
        assert torch.equal(r, ref * np.float32(math.sqrt(0.1919)))


@pytest.mark.parametrize("n", [4, 260, 1_000_000, 3 * 199_212])
@pytest.mark.parametrize("mode", ["first", "middle_snap", "last"])
def test_k5_leapfrog_bit_exact_vs_oracle(C, n, mode):
    rng = np.random.RandomState(2)
    th, r, g = (rng.randn(n).astype(np.float32) for _ in range(3))
    eps, inv_mass, tau = 2.09e-4, 1 / 0.1919, 100.0
    kick = 0.5 * eps if mode != "middle_snap" else eps
    drift = 0.0 if mode == "last" else eps * inv_mass
    t_d, r_d, g_d = dev(th), dev(r), dev(g)
    snap = torch.zeros(n, device=DEV) if mode == "middle_snap" else None
    C.hmc_leapfrog(t_d, r_d, g_d, kick=kick, drift=drift, tau=tau, tau_out=1.0, snapshot=snap)
    te, re = R.hmc_leapfrog_update(th, r, g, kick, drift, tau)
    assert np.array_equal(r_d.cpu().numpy(), re)
    assert np.array_equal(t_d.cpu().numpy(), te)
    assert np.array_equal(g_d.cpu().numpy(), g)
    if snap is not None:
        assert np.array_equal(snap.cpu().numpy(), te)


def test_k5_leapfrog_argument_errors(C):
    t = torch.zeros(8, device=DEV)
    with pytest.raises(ValueError):
        C.hmc_leapfrog(t, t.clone(), torch.zeros(4, device=DEV), kick=0.1, drift=0.1, tau=1.0)
    with pytest.raises(ValueError):
        C.hmc_leapfrog(t, t.clone(), t.clone(), kick=0.1, drift=0.0, tau=1.0, snapshot=t.clone())   # snapshot w/o drift
    with pytest.raises(ValueError):
        C.hmc_leapfrog(t.cpu(), t, t, kick=0.1, drift=0.1, tau=1.0)
    with pytest.raises(ValueError):
        C.hmc_momentum(torch.zeros(6, device=DEV), 1.0)                                             # n % 4 != 0


@pytest.mark.parametrize("chains,D", [(1, 5), (3, 8191), (7, 8192), (4, 8193), (128, 199_210), (2, 1_000_003)])
def test_k5_energy_matches_fp64(C, chains, D):
    ld = (D + 3) // 4 * 4
    rng = np.random.RandomState(3)
    th = np.zeros((chains, ld), np.float32)
    r = np.zeros((chains, ld), np.float32)
    th[:, :D] = rng.randn(chains, D) * 0.3
    r[:, :D] = rng.randn(chains, D)
    th[:, D:] = 7.0                                   # padding must not leak into the sums
    r[:, D:] = -3.0
    sums, _ = C.hmc_energy(dev(th), dev(r), D)
    s = sums.cpu().numpy()
    for c in range(chains):
        a, b = R.hmc_energy_sums(th[c, :D], r[c, :D])
        assert abs(s[0, c] - a) <= 1e-12 * a and abs(s[1, c] - b) <= 1e-12 * b
    again, _ = C.hmc_energy(dev(th), dev(r), D)
    assert torch.equal(sums, again)                   # deterministic reduction order


def test_k5_accept_commit_restore_and_flags(C):
    chains, ld = 6, 1028
    rng = np.random.RandomState(4)
    theta, saved, cand, kept = (rng.randn(chains, ld).astype(np.float32) for _ in range(4))
    h_old = np.array([10.0, 10.0, 10.0, 10.0, 10.0, float("nan")])
    h_new = np.array([9.0, 10.5, 10.5, float("inf"), float("nan"), 9.0])
    logu = np.log(np.array([0.9, 0.9, 0.5, 1e-9, 1e-9, 1e-9], np.float32))   # exp(-0.5) = 0.607
    want = [R.hmc_accept(float(a), float(b), float(u)) for a, b, u in zip(h_old, h_new, logu)]
    assert want == [True, False, True, False, False, False]
    t_d, s_d, c_d, k_d = dev(theta), dev(saved), dev(cand), dev(kept)
    out = torch.zeros(chains, ld + 4, device=DEV)[:, :ld]
    acc = torch.zeros(chains, dtype=torch.int32, device=DEV)
    C.hmc_accept(t_d, s_d, dev(h_old), dev(h_new), acc, logu=dev(logu))
    assert acc.cpu().tolist() == [int(w) for w in want]
    for c, w in enumerate(want):
        exp = theta[c] if w else saved[c]
        assert np.array_equal(t_d[c].cpu().numpy(), exp) and np.array_equal(s_d[c].cpu().numpy(), exp)
    # with the keep pair + strided output rows: out = kept-state, theta/saved as before
    t_d, s_d = dev(theta), dev(saved)
    C.hmc_accept(t_d, s_d, dev(h_old), dev(h_new), acc, logu=dev(logu), keep_dst=k_d, keep_src=c_d, out=out)
    for c, w in enumerate(want):
        exp = cand[c] if w else kept[c]
        assert np.array_equal(k_d[c].cpu().numpy(), exp) and np.array_equal(out[c].cpu().numpy(), exp)
    # Philox uniforms: acceptance frequency of a fixed energy gap matches exp(-gap)
    n = 20_000
    th = torch.zeros(n, 4, device=DEV)
    acc = torch.zeros(n, dtype=torch.int32, device=DEV)
    C.hmc_accept(th, th.clone(), torch.zeros(n, dtype=torch.float64, device=DEV),
                 torch.full((n,), 0.7, dtype=torch.float64, device=DEV), acc, seed=11, step=2)
    rate = acc.float().mean().item()
    assert abs(rate - math.exp(-0.7)) < 4 * math.sqrt(0.25 / n)


# ------------------------------------------------------------------------------------------------ class API
def _mlp_problem(n=96, d=12, c=3, seed=0):
    from ursabench_b200 import models
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    y = torch.randint(0, c, (n,), generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)
    torch.manual_seed(seed)
    return models.MLP(16, d, c), x, y, loader


def _cpu_nll_and_grad(model, x, y):
    import copy
    m = copy.deepcopy(model).cpu().double()          # fp64 CPU gradient of the same module

    def f(theta):
        off = 0
        for p in m.parameters():
            n = p.numel()
            p.data = torch.from_numpy(np.asarray(theta[off:off + n], np.float64)).view(p.shape).clone()
            p.grad = None
            off += n
        loss = torch.nn.functional.cross_entropy(m(x.double()), y, reduction="sum")
        loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy().astype(np.float32)
        return float(loss.item()), g
    return f


@pytest.mark.parametrize("burn", [0, 2, -1, -2])
def test_hmc_class_trajectory_matches_restatement(burn):
    from ursabench_b200 import inference
    model, x, y, loader = _mlp_problem()
    hyp = {"step_size": 5e-3, "num_samples": 6, "L": 4, "tau": 10.0, "burn": burn, "mass": 0.5}
    inf = inference.HMC(hyperparameters=dict(hyp), model=model, train_loader=loader, device=DEV)
    D, ld = inf.D, inf.ld
    rng = np.random.RandomState(7)
    z = [rng.randn(1, ld).astype(np.float32) for _ in range(hyp["num_samples"])]
    logu = [np.log(rng.rand(1).astype(np.float32)) for _ in range(hyp["num_samples"])]
    logu[4] = np.array([1e-3], np.float32)            # > 0 can never be accepted -> a guaranteed rejection
    inf._inject = ([dev(a) for a in z], [dev(a) for a in logu])
    theta0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    samples = inf.sample()
    ret, accepts = R.hmc_chain(theta0, _cpu_nll_and_grad(model, x, y), [a[0, :D] for a in z], [float(u[0]) for u in logu],
                               hyp["step_size"], hyp["L"], hyp["tau"], hyp["mass"])
    assert accepts[4] is False and any(accepts)
    want = ret[burn * hyp["L"]::hyp["L"]]             # the reference wrapper's thinning (hmc.py:80)
    assert len(samples) == len(want)
    for s, w in zip(samples, want):
        got = torch.cat([p.detach().reshape(-1) for p in s.parameters()]).numpy()
        np.testing.assert_allclose(got, w, rtol=0, atol=2e-5)
    assert all(isinstance(s, torch.nn.Module) for s in samples)
    assert abs(float(inf.acceptance_rate[0]) - np.mean(accepts)) < 1e-12


def test_hmc_default_hyperparameters_and_api_surface():
    from ursabench_b200 import inference
    model, x, y, loader = _mlp_problem(seed=1)
    inf = inference.HMC(hyperparameters=None, model=model, train_loader=loader, device=DEV)
    assert (inf.step_size, inf.num_samples, inf.L, inf.tau, inf.burn, inf.mass) == (0.001, 10, 1, 0.1, -1, 1.0)
    out = inf.sample()
    assert len(out) == 1                              # samples[-1*L::L] keeps one entry (reference :80 with burn = -1)
    inf.update_hyp({"step_size": 1e-3, "num_samples": 3, "L": 2, "tau": 1.0, "burn": 0, "mass": 1.0})
    out = inf.sample(debug=False)
    assert len(out) == 4                              # initial point + one state per iteration
    logits = out[-1](x)                               # handles behave like CPU modules
    assert logits.shape == (96, 3)
    with pytest.raises(RuntimeError):
        inference.HMC(hyperparameters=None, model=model, train_loader=loader, device=torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        inference.HMC(hyperparameters=None, model="not a module", train_loader=loader, device=DEV)


def test_hmc_conjugate_gaussian_target_many_chains():
    """Bayesian linear regression: posterior N(P^-1 X^T y, P^-1), P = tau I + X^T X.  2048 chains started from one
    point; after a short burn-in the cross-chain mean / covariance must match within Monte-Carlo error."""
    from ursabench_b200 import inference
    d, n, tau = 6, 40, 2.0
    g = torch.Generator().manual_seed(3)
    X = torch.randn(n, d, generator=g)
    yv = X @ torch.randn(d, generator=g) + 0.5 * torch.randn(n, generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(X, yv), batch_size=n, shuffle=False)
    model = torch.nn.Linear(d, 1, bias=False)
    chains = 2048
    torch.manual_seed(123)
    # L * eps * omega stays within ~[0.9, 2.0] rad for every posterior mode (omega^2 in [~17, ~80]): no HMC resonance
    hyp = {"step_size": 0.0275, "num_samples": 40, "L": 8, "tau": tau, "burn": 40, "mass": 1.0, "num_chains": chains}
    inf = inference.HMC(hyperparameters=hyp, model=model, train_loader=loader, model_loss="regression", device=DEV)
    out = inf.sample()
    assert len(out) == chains
    w = inf.bank.w[:chains, :d].double().cpu().numpy()
    P = tau * np.eye(d) + (X.T @ X).double().numpy()
    cov = np.linalg.inv(P)
    mean = cov @ (X.T @ yv).double().numpy()
    se = np.sqrt(np.diag(cov) / chains)
    assert np.all(np.abs(w.mean(0) - mean) < 5 * se)
    emp = np.cov(w.T)
    assert np.all(np.abs(emp - cov) < 6 * np.sqrt((np.outer(np.diag(cov), np.diag(cov)) + cov ** 2) / chains))
    acc = inf.acceptance_rate.numpy()
    assert 0.6 < acc.mean() <= 1.0
    # chains on another rank (disjoint Philox base) draw different momenta
    r1, r2 = torch.empty(chains, 8, device=DEV), torch.empty(chains, 8, device=DEV)
    from ursabench_b200 import _C
    _C.hmc_momentum(r1, 1.0, seed=inf.seed, step=1, elem_offset=0)
    _C.hmc_momentum(r2, 1.0, seed=inf.seed, step=1, elem_offset=chains * 8)
    assert not torch.equal(r1, r2)


def test_hmc_mlp_gemm_gradient_matches_vmap_grad():
    """The chain-batched GEMM formulation of the MLP likelihood gradient (HMC._build_grad_fn_mlp) against torch.func's
    vmap(grad) of the module: same fp32 arithmetic up to summation order."""
    from ursabench_b200 import inference, models
    torch.manual_seed(3)
    x, y = torch.randn(70, 1, 6, 6), torch.randint(0, 7, (70,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)
    hyp = {"step_size": 1e-3, "num_samples": 1, "L": 2, "tau": 10.0, "burn": 0, "mass": 1.0, "num_chains": 5}
    inf = inference.HMC(hyperparameters=dict(hyp), model=models.MLP(40, 36, 7), train_loader=loader, device=DEV)
    inf.sample()
    assert inf.grad_engine == "mlp_tcgen05_fused_f16"       # the hand-written chain-batched forward + backward (FP16-split) is the default
    th = torch.randn(5, inf.ld, device=DEV) * 0.2
    g0, ce0 = torch.zeros_like(th), torch.zeros(5, device=DEV)
    inf._grad(th, g0, ce0)                                   # ursa_hmc_mlp_grad
    inf._grad_fn = inf._build_grad_fn()                      # vmap(grad) of the module
    gv, cev = torch.zeros_like(th), torch.zeros(5, device=DEV)
    inf._grad(th, gv, cev)
    assert (g0 - gv)[:, :inf.D].abs().max().item() < 2e-5 * gv.abs().max().item()
    assert torch.allclose(ce0, cev, rtol=2e-6, atol=1e-4)
    from ursabench_b200.tasks._engine import _arch_of
    theta = torch.randn(5, inf.ld, device=DEV) * 0.2
    inf._grad_fn = inf._build_grad_fn_mlp_tc(_arch_of(inf.model))        # the same GEMMs on ursa_gemm_nt_3xtf32
    g1, ce1 = torch.zeros_like(theta), torch.zeros(5, device=DEV)
    inf._grad(theta, g1, ce1)
    inf._grad_fn = inf._build_grad_fn()                     # the generic engine
    g2, ce2 = torch.zeros_like(theta), torch.zeros(5, device=DEV)
    inf._grad(theta, g2, ce2)
    scale = g2.abs().max().item()
    assert (g1 - g2).abs().max().item() < 2e-5 * scale
    assert torch.allclose(ce1, ce2, rtol=2e-6, atol=1e-4)
    inf._grad_fn = inf._build_grad_fn_mlp(_arch_of(inf.model))          # the cuBLAS fp32 formulation of the same GEMMs
    g3, ce3 = torch.zeros_like(theta), torch.zeros(5, device=DEV)
    inf._grad(theta, g3, ce3)
    assert (g3 - g2).abs().max().item() < 2e-5 * scale and torch.allclose(ce3, ce2, rtol=2e-6, atol=1e-4)
    # a module the fast path does not cover keeps the generic engine
    inf2 = inference.HMC(hyperparameters=dict(hyp), model=torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(36, 7)),
                         train_loader=loader, device=DEV)
    inf2.sample()
    assert inf2.grad_engine == "vmap"


@pytest.mark.parametrize("batch,M,N,K,shared,relu", [(3, 70, 40, 36, True, True), (2, 257, 10, 200, False, False),
                                                       (4, 10, 200, 1000, False, False), (1, 130, 129, 31, True, False)])
def test_gemm_nt_3xtf32_matches_fp64(batch, M, N, K, shared, relu):
    """ursa_gemm_nt_3xtf32 (the MLP engine as a plain batched GEMM): ragged M / N / K, shared or batched A, strided operands,
    bias + ReLU epilogue, against an fp64 product."""
    from ursabench_b200 import _C
    torch.manual_seed(batch * 1000 + M)
    A = torch.randn((M, K + 3) if shared else (batch, M, K + 3), device=DEV)[..., :K]          # lda > K
    Bm = torch.randn(batch, N, K + 5, device=DEV)[..., :K]
    bias = torch.randn(batch, N, device=DEV)
    out = torch.full((batch, M, N + 2), float("nan"), device=DEV)[..., :N]                      # ldo > N
    _C.gemm_nt(A, Bm, out, bias=bias, relu=relu)
    ref = torch.matmul((A if not shared else A[None]).double(), Bm.double().transpose(1, 2)) + bias.double()[:, None, :]
    if relu:
        ref = ref.clamp_min(0)
    err = (out.double() - ref).abs().max().item()
    assert err < 3e-6 * max(1.0, ref.abs().max().item()) * (K ** 0.5) / 8 + 1e-6, err


def test_hmc_trajectory_graph_replay_equals_eager():
    """The leapfrog trajectory captured as ONE CUDA graph (iterations >= 2 replay it) must reproduce the eager loop bit
    for bit: same chains, same accept decisions, same kept samples."""
    from ursabench_b200 import inference, models
    torch.manual_seed(4)
    x, y = torch.randn(70, 1, 6, 6), torch.randint(0, 7, (70,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=70, shuffle=False)
    hyp = {"step_size": 2e-3, "num_samples": 6, "L": 3, "tau": 10.0, "burn": 0, "mass": 1.0, "num_chains": 4}
    outs = []
    for use_graph in (False, True):
        torch.manual_seed(11)
        inf = inference.HMC(hyperparameters=dict(hyp), model=models.MLP(40, 36, 7), train_loader=loader, device=DEV)
        inf.use_cuda_graph = use_graph
        handles = inf.sample()
        assert inf.graph_replays == (5 if use_graph else 0)
        outs.append((torch.stack([inf.bank.w[h._ursa_row, :inf.D] for h in handles]).clone(), inf.acceptance_rate.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert float(outs[0][1].mean()) > 0.2


@pytest.mark.parametrize("engine", ["tf32", "f16"])
@pytest.mark.parametrize("C,N,in_dim,hid,ncls", [(3, 70, 36, 40, 7), (2, 257, 784, 200, 10), (1, 1000, 16, 8, 3), (4, 33, 20, 132, 100)])
def test_hmc_mlp_fused_gradient_matches_autograd(C, N, in_dim, hid, ncls, engine):
    """``ursa_hmc_mlp_grad`` / ``ursa_hmc_mlp_grad_f16`` (eight tcgen05 GEMMs, 3xTF32 or persistent 2xFP16-split, with split / transposed operands handed from epilogue to epilogue)
    against fp64 autograd of the same MLP: gradient of the summed cross entropy per chain, and its value; ragged point
    counts, hidden widths that are not a multiple of the N tile, more classes than one 16-column chunk."""
    from ursabench_b200 import _C
    torch.manual_seed(C * 1000 + N)
    D = hid * in_dim + hid + hid * hid + hid + ncls * hid + ncls
    ld = (D + 3) // 4 * 4
    theta = torch.zeros(C, ld, device=DEV)
    theta[:, :D] = torch.randn(C, D, device=DEV) * (1.5 / in_dim ** 0.5)
    x = torch.randn(N, in_dim, device=DEV)
    y = torch.randint(0, ncls, (N,), device=DEV)
    g = torch.full((C, ld), 7.0, device=DEV)
    ce = torch.zeros(C, device=DEV)
    ws = _C.hmc_mlp_grad(theta, x, y, in_dim, hid, ncls, g, ce, engine=engine)
    assert ws is not None
    o = [0, hid * in_dim, hid * in_dim + hid, hid * in_dim + hid + hid * hid, hid * in_dim + 2 * hid + hid * hid,
         hid * in_dim + 2 * hid + hid * hid + ncls * hid, D]
    for c in range(C):
        t = theta[c, :D].double().clone().requires_grad_(True)
        W1, b1, W2, b2, W3, b3 = (t[o[i]:o[i + 1]] for i in range(6))
        h1 = torch.relu(x.double() @ W1.view(hid, in_dim).t() + b1)
        h2 = torch.relu(h1 @ W2.view(hid, hid).t() + b2)
        loss = torch.nn.functional.cross_entropy(h2 @ W3.view(ncls, hid).t() + b3, y, reduction="sum")
        loss.backward()
        scale = t.grad.abs().max().item()
        assert (g[c, :D].double() - t.grad).abs().max().item() < 2e-5 * scale, (c, scale)
        assert abs(ce[c].item() - loss.item()) < 2e-6 * abs(loss.item()) + 1e-4
    if ld > D:
        assert (g[:, D:] == 7.0).all()                        # padding untouched
    with pytest.raises(ValueError):
        _C.hmc_mlp_grad(theta, x, y.int(), in_dim, hid, ncls, g, ce)
    assert _C.lib().ursa_hmc_mlp_grad_workspace(2, 10, 30, 40, 5) == 0      # in_dim % 4 != 0: not covered
