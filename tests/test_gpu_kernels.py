"""GPU parity tests proper: every kernel is called through the C ABI (ursabench_b200._C -> libursa_b200.so)
and compared with the golden fixtures of the live reference and with the CPU oracle on seeded inputs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _npz(name):
    return np.load(os.path.join(GOLD, name))


def _json(name):
    return json.load(open(os.path.join(GOLD, name)))


@pytest.fixture(scope="module")
def C():
    from ursabench_b200 import _C
    _C.lib()
    return _C


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------- K1
CASES = ["sgld_wd_noise", "sgld_nowd_nonoise", "sghmc_wd_noise", "sghmc_nowd_noise", "sghmc_wd_nonoise"]


@pytest.mark.parametrize("case", CASES)
def test_k1_matches_reference_golden(C, case):
    """identical grad + noise -> the reference's optimSGHMC.step results (north star: rtol 1e-6; observed: bit-exact)."""
    g = _npz("sgmcmc_step.npz")
    lr0, mom, wd, n_train, noise = g[case + "/hyper"]
    p = dev(g[case + "/init"])
    v = torch.zeros_like(p) if mom != 0 else None
    for t in range(4):
        lr = float(g["%s/lr%d" % (case, t)])
        C.sgmcmc_step(p, dev(g["%s/g%d" % (case, t)]), v, None, dev(g["%s/z%d" % (case, t)]) if noise else None,
                      lr=lr, momentum=mom, wd_over_n=wd / n_train, noise_mul=np.sqrt(2 * (1 - mom) * lr),
                      noise_div=n_train, first_step=(t == 0), add_noise=bool(noise))
        ref = g["%s/p%d" % (case, t)]
        got = p.cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=0)
        assert np.array_equal(got, ref), "not bit-exact: %d elements differ" % (got != ref).sum()
        if mom != 0:
            assert np.array_equal(v.cpu().numpy(), g["%s/v%d" % (case, t)])


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 1 << 20, (1 << 20) + 3, 2_000_003])
@pytest.mark.parametrize("mom", [0.0, 0.9])
def test_k1_matches_oracle_sizes(C, n, mom):
    rng = np.random.RandomState(n % 1000 + int(mom * 10))
    p0, g0, v0, z0 = (rng.randn(n).astype(np.float32) for _ in range(4))
    lr, wd, N = 0.0123, 7.5, 50000
    for first in (True, False):
        p, g, v, z = dev(p0), dev(g0), dev(v0), dev(z0)
        snap = torch.full_like(p, -7.0)
        C.sgmcmc_step(p, g, v if mom else None, snap, z, lr=lr, momentum=mom, wd_over_n=wd / N,
                      noise_mul=np.sqrt(2 * (1 - mom) * lr), noise_div=N, first_step=first, add_noise=True,
                      zero_grad=True)
        pe, ve = R.sgmcmc_step(p0, g0, v0, z0, lr, mom, wd, N, first, True)
        assert np.array_equal(p.cpu().numpy(), pe)
        assert np.array_equal(snap.cpu().numpy(), pe)            # thinned-sample snapshot row
        assert not g.any()                                       # fused zero_grad
        if mom:
            assert np.array_equal(v.cpu().numpy(), ve)


def test_k1_philox_stream_matches_oracle(C):
    n, seed, step = 4099, 0x1234ABCD5678, 42
    out = torch.empty(n, device="cuda")
    C.philox_normal(out, seed, step)
    z, _ = R.philox_normals(n, seed, step)
    np.testing.assert_allclose(out.cpu().numpy(), z, atol=3e-5, rtol=1e-4)       # MUFU lg2/sin/cos vs fp64
    out2 = torch.empty(1000, device="cuda")
    C.philox_normal(out2, seed, step, elem_offset=2048)
    assert torch.equal(out2, out[2048:3048])                                     # counter = global element index
    C.philox_normal(out2, seed, step + 1, elem_offset=2048)
    assert not torch.equal(out2, out[2048:3048])


@pytest.mark.parametrize("mom", [0.0, 0.5])
def test_k1_philox_mode_consistent_and_distribution(C, mom):
    n = (1 << 21) + 1
    torch.manual_seed(0)
    p0, g0, v0 = torch.randn(n, device="cuda"), torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    lr, N = 0.01, 100.0
    kw = dict(lr=lr, momentum=mom, wd_over_n=0.0, noise_mul=np.sqrt(2 * (1 - mom) * lr), noise_div=N)
    p_a, v_a = p0.clone(), v0.clone()
    C.sgmcmc_step(p_a, g0, v_a if mom else None, **kw, add_noise=True, seed=99, step=5)
    p_b, v_b = p0.clone(), v0.clone()
    C.sgmcmc_step(p_b, g0, v_b if mom else None, **kw, add_noise=False)
    z = torch.empty(n, device="cuda")
    C.philox_normal(z, 99, 5)
    scale = np.float32(kw["noise_mul"] / N)
    zhat = ((p_a - p_b) / scale).double()
    # the injected noise is the documented Philox stream (up to fp32 rounding of p + u)
    assert (zhat - z.double()).abs().max().item() < 5e-3
    assert abs(z.mean().item()) < 4.0 / np.sqrt(n) and abs(z.var().item() - 1) < 0.01
    assert abs((z ** 3).mean().item()) < 0.02 and abs((z ** 4).mean().item() - 3) < 0.05
    # different step / seed -> different stream; same -> identical
    p_c = p0.clone(); v_c = v0.clone()
    C.sgmcmc_step(p_c, g0, v_c if mom else None, **kw, add_noise=True, seed=99, step=5)
    assert torch.equal(p_a, p_c)
    C.sgmcmc_step(p_c, g0, v_c if mom else None, **kw, add_noise=True, seed=99, step=6)
    assert not torch.equal(p_a, p_c)


def test_k1_gaussian_target_sgld_distribution(C):
    """Distributional check on a Gaussian target: U(theta) = N_train/2 * theta^2/sigma^2 scaled like the reference's
    mean-loss convention.  SGLD with the reference's scalings (noise sqrt(2 lr)/N, grad of the mean loss) samples
    N(0, sigma_post^2); we run 65536 independent 1-d chains (one flat buffer) and compare mean / variance."""
    chains, steps, burn = 1 << 16, 600, 200
    N = 10.0                      # "dataset size": the posterior is exp(-N * L(theta)), L = theta^2 / (2 s2)
    s2 = 0.5
    lr = 0.02 * N                 # effective step h = lr / N on the un-normalised log posterior
    p = torch.zeros(chains, device="cuda")
    acc1 = torch.zeros_like(p, dtype=torch.float64)
    acc2 = torch.zeros_like(p, dtype=torch.float64)
    for t in range(steps):
        g = p / s2                                            # d/dtheta of the mean loss
        C.sgmcmc_step(p, g, None, None, None, lr=lr, momentum=0.0, wd_over_n=0.0,
                      noise_mul=np.sqrt(2 * lr), noise_div=N, add_noise=True, seed=7, step=t)
        if t >= burn:
            acc1 += p.double()
            acc2 += p.double() ** 2
    # stationary variance of the Euler discretisation: with h = lr/N on U = N theta^2/(2 s2), a = 1 - lr/s2:
    # var = (2 lr / N^2) / (1 - a^2)
    a = 1 - lr / s2
    var_expected = (2 * lr / N ** 2) / (1 - a ** 2)
    m = (acc1 / (steps - burn)).mean().item()
    var = (acc2 / (steps - burn)).mean().item()
    assert abs(m) < 5e-3
    assert abs(var - var_expected) / var_expected < 0.02


def test_k1_argument_errors(C):
    p = torch.zeros(8)
    with pytest.raises(ValueError):
        C.sgmcmc_step(p, p, None, lr=0.1, momentum=0.0, wd_over_n=0.0)          # CPU tensors: no CPU path
    pc = torch.zeros(8, device="cuda")
    with pytest.raises(ValueError):
        C.sgmcmc_step(pc, pc.clone(), None, lr=0.1, momentum=0.5, wd_over_n=0.0)  # momentum without v
    with pytest.raises(ValueError):
        C.sgmcmc_step(pc, pc.clone(), pc.clone(), lr=0.1, momentum=-0.5, wd_over_n=0.0)


# ----------------------------------------------------------------------------- K2
@pytest.mark.parametrize("mode", ["textbook", "swa_biased", "compat_n0"])
def test_k2_collect_matches_reference_golden(C, mode):
    g = _npz("swa_collect.npz")
    ws = g[mode + "/w"]
    D = ws.shape[1]
    K = 3
    mean, sq = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ld = (D + 3) // 4 * 4
    ring = torch.zeros(K, ld, device="cuda")
    n = 0
    for k in range(ws.shape[0]):
        if mode == "swa_biased":
            n += 1
        C.swag_collect(dev(ws[k]), mean, sq, ring[k % K], 0 if mode == "compat_n0" else n)
        if mode == "textbook":
            n += 1
        assert np.array_equal(mean.cpu().numpy(), g["%s/mean%d" % (mode, k)])
        assert np.array_equal(sq.cpu().numpy(), g["%s/sq%d" % (mode, k)])
        ref_ring = g["%s/ring%d" % (mode, k)]                     # oldest first
        rows = [ring[j % K, :D].cpu().numpy() for j in range(max(0, k + 1 - K), k + 1)]
        assert np.array_equal(np.stack(rows), ref_ring)
    var = torch.empty(D, device="cuda")
    C.swag_variance(mean, sq, var)
    assert np.array_equal(var.cpu().numpy(), g[mode + "/var"])


@pytest.mark.parametrize("D,K,S", [(1000, 0, 1), (1000, 5, 4), (4099, 20, 30), (1024 * 3 + 1, 24, 32), (2_050_001, 20, 8),
                                   (513, 3, 16), (4099, 20, 70), (1000, 0, 33)])      # S > 32: grouped launches, same stream
def test_k2_draw_matches_oracle(C, D, K, S):
    rng = np.random.RandomState(D % 97)
    ld = (D + 3) // 4 * 4
    mean, var = rng.randn(D).astype(np.float32), (rng.rand(D) + 0.01).astype(np.float32)
    z1 = rng.randn(S, D).astype(np.float32)
    ring = rng.randn(K, D).astype(np.float32) if K else None
    z2 = rng.randn(S, K).astype(np.float32) if K else None

    def padded(a):
        t = torch.zeros(a.shape[0], ld, device="cuda")
        t[:, :D] = dev(a)
        return t

    out = torch.full((S, ld), 123.0, device="cuda")
    C.swag_draw(out, dev(mean), dev(var), D, ring=padded(ring) if K else None, z2=dev(z2) if K else None,
                z1=padded(z1), rank_div=float((20 - 1) ** 0.5))
    ref = R.swag_draw(mean, var, z1, ring, z2, max_rank=20)
    np.testing.assert_allclose(out[:, :D].cpu().numpy(), ref, rtol=2e-5, atol=2e-5)
    if ld > D:
        assert (out[:, D:] == 123.0).all()                        # padding untouched
    if not K:
        assert np.array_equal(out[:, :D].cpu().numpy(), ref)      # diag draw: same roundings as mean + std * z
    # Philox mode == external mode fed with the documented stream: element (s, d) = normal s % 6 of block (s // 6) * D + d
    # (a diagonal draw with mean 0, var 1 returns the stream itself; it is pinned against the oracle's restatement)
    out_p = torch.empty((S, ld), device="cuda")
    C.swag_draw(out_p, dev(mean), dev(var), D, ring=padded(ring) if K else None, z2=dev(z2) if K else None,
                z1=None, rank_div=float((20 - 1) ** 0.5), seed=5, step=9)
    z1s = C.swag_draw_noise(S, D, 5, 9, "cuda")
    np.testing.assert_allclose(z1s[:, :D].cpu().numpy(), R.draw_normals(S, D, 5, 9), atol=2e-4, rtol=1e-4)   # MUFU vs fp64 (lg2.approx near u = 1: |dz| <= 1e-4 with probability 1e-6)
    out_e = torch.empty((S, ld), device="cuda")
    C.swag_draw(out_e, dev(mean), dev(var), D, ring=padded(ring) if K else None, z2=dev(z2) if K else None,
                z1=z1s, rank_div=float((20 - 1) ** 0.5))
    assert torch.equal(out_p[:, :D], out_e[:, :D])


# ----------------------------------------------------------------------------- K3 epilogue / K4
@pytest.mark.parametrize("name,S", [("mlp", 5), ("preresnet8", 2)])
def test_k3_accumulate_matches_reference_golden(C, name, S):
    g = _npz("prediction.npz")
    logits = dev(g[name + "/logits"])
    _, N, Cc = logits.shape
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    C.bma_accumulate(logits[:3].contiguous(), P, E)               # accumulates across calls like update_statistics
    if S > 3:
        C.bma_accumulate(logits[3:].contiguous(), P, E)
    elif S < 3:
        pass
    ref_p = g[name + "/ensemble_proba"]
    np.testing.assert_allclose(P.cpu().numpy(), ref_p, atol=1e-5, rtol=0)         # north star: 1e-5
    np.testing.assert_allclose(P.cpu().numpy(), ref_p, atol=5e-7, rtol=2e-6)      # observed
    np.testing.assert_allclose(E.cpu().numpy(), g[name + "/entropy"], atol=2e-6, rtol=1e-5)


def test_k3_accumulate_c100_and_wide(C):
    rng = np.random.RandomState(1)
    for Cc in (2, 33, 100, 257, 1000):
        logits = (rng.randn(3, 50, Cc) * 3).astype(np.float32)
        P, E = torch.zeros(50, Cc, device="cuda"), torch.zeros(50, device="cuda")
        C.bma_accumulate(dev(logits), P, E)
        rp, re = R.bma_accumulate(logits)
        np.testing.assert_allclose(P.cpu().numpy(), rp, atol=1e-6, rtol=1e-5)
        np.testing.assert_allclose(E.cpu().numpy(), re, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["rand_c10", "rand_c100", "rand_c2", "edges_c4"])
def test_k4_counters_bit_exact_and_metrics(C, name):
    g = _npz("metrics_edge.npz")
    ref = _json("metrics_edge.json")[name]
    psum, y, S = g[name + "/proba_sum"], g[name + "/y"], int(g[name + "/S"])
    oi, of, pred, conf = C.bma_metrics(dev(psum), S, dev(y), want_rows=True)
    oi, of = oi.cpu().numpy(), of.cpu().numpy()
    c = R.bma_counters(psum, S, y)
    nb = 15
    assert oi[0] == c["correct"]
    assert np.array_equal(oi[1:1 + nb], c["bin_count"])
    assert np.array_equal(oi[1 + nb:], c["bin_correct"])
    assert np.array_equal(pred.cpu().numpy(), c["pred"])
    assert np.array_equal(conf.cpu().numpy(), c["conf"])           # same fp32 division
    np.testing.assert_allclose(of[2:], c["bin_conf_sum"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(of[0], c["nll_sum"], rtol=1e-12)
    np.testing.assert_allclose(of[1], c["brier_sum"], rtol=1e-12)
    m = R.metrics_from_counters(dict(correct=int(oi[0]), bin_count=oi[1:1 + nb], bin_correct=oi[1 + nb:],
                                     bin_conf_sum=of[2:], nll_sum=of[0], brier_sum=of[1], n=len(y)))
    assert m["error_rate"] == pytest.approx(ref["error_rate"], abs=1e-15)
    assert m["brier_score"] == pytest.approx(ref["brier_score"], rel=1e-12)
    assert m["nll"] == pytest.approx(ref["nll"], rel=2e-6)
    assert m["ece"] == pytest.approx(ref["ece"], abs=2e-7)
    # deterministic: fixed reduction order
    oi2, of2, _, _ = C.bma_metrics(dev(psum), S, dev(y))
    assert torch.equal(oi2.cpu(), torch.from_numpy(oi)) and torch.equal(of2.cpu(), torch.from_numpy(of))


# ----------------------------------------------------------------------------- K3 forward, MLP
@pytest.mark.parametrize("name,S", [("mlp", 5), ("mlp_c100", 3)])
def test_k3_mlp_forward_matches_reference_golden(C, name, S):
    g = _npz("prediction.npz")
    hidden, in_dim, Cc = (int(v) for v in g[name + "/arch"])
    bank = dev(g[name + "/bank"])
    x = dev(g[name + "/x"].reshape(g[name + "/x"].shape[0], -1))
    N = x.shape[0]
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, Cc, device="cuda")
    C.bma_mlp_forward(bank, S, x, in_dim, hidden, Cc, P, E, logits_out=logits, algo=C.ALGO_FFMA)
    if name + "/logits" in g:
        np.testing.assert_allclose(logits.cpu().numpy(), g[name + "/logits"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(P.cpu().numpy(), g[name + "/ensemble_proba"], atol=1e-5, rtol=0)   # north star
    np.testing.assert_allclose(E.cpu().numpy(), g[name + "/entropy"], atol=1e-5, rtol=1e-5)


def test_k3_mlp_forward_config1_shape_vs_torch(C):
    """MLP 784-400-400-10 (config 1), S=6, N=1000, against a plain PyTorch fp32 forward on the GPU with TF32 off."""
    from ursabench_b200.models import MLP
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    S, N = 6, 1000
    ms = [MLP(400, 784, 10).cuda() for _ in range(S)]
    for m in ms:
        for p in m.parameters():
            p.data.mul_(2.0)
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms])
    x = torch.randn(N, 1, 28, 28, device="cuda")
    P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, 10, device="cuda")
    C.bma_mlp_forward(bank, S, x.view(N, -1).contiguous(), 784, 400, 10, P, E, logits_out=logits)
    with torch.no_grad():
        ref = torch.stack([m(x) for m in ms])
    assert (logits - ref).abs().max().item() < 2e-5
    pref = torch.softmax(ref.double(), -1).sum(0)
    assert (P.double() - pref).abs().max().item() < 1e-5


# ----------------------------------------------------------------------------- K3 forward, PreResNet
def _algo(C, name):
    return {"ffma": C.ALGO_FFMA, "tcgen05": C.ALGO_TCGEN05, "fused": C.ALGO_TCGEN05_FUSED,
            "fused16": C.ALGO_TCGEN05_FUSED_F16}[name]


@pytest.mark.parametrize("algo", ["ffma", "tcgen05", "fused", "fused16"])
def test_k3_preresnet8_forward_matches_reference_golden(C, algo):
    g = _npz("prediction.npz")
    algo = _algo(C, algo)
    bank, bufs = dev(g["preresnet8/bank"]), dev(g["preresnet8/buffers"])
    x = dev(g["preresnet8/x"].astype(np.float32))
    S, N, Cc = 2, x.shape[0], 10
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, Cc, device="cuda")
    C.bma_preresnet_forward(bank, bufs, S, x, 8, Cc, P, E, logits_out=logits, algo=algo)
    torch.cuda.synchronize()
    ref = g["preresnet8/logits"]
    err = np.abs(logits.cpu().numpy() - ref).max()
    assert err < 1e-4 * max(1.0, np.abs(ref).max()), err
    np.testing.assert_allclose(P.cpu().numpy(), g["preresnet8/ensemble_proba"], atol=1e-5, rtol=0)   # north star
    np.testing.assert_allclose(E.cpu().numpy(), g["preresnet8/entropy"], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("algo", ["ffma", "tcgen05", "fused", "fused16"])
@pytest.mark.parametrize("depth,S,N,Cc", [(20, 3, 70, 10), (14, 9, 5, 100), (20, 1, 513, 10), (8, 2, 7, 10)])
def test_k3_preresnet_forward_vs_torch_fp32(C, depth, S, N, Cc, algo):
    """PreResNet-20 (config 2/5) with random BN statistics against a plain PyTorch fp32 forward (TF32 off);
    N not a multiple of the per-CTA image group, S crossing the sample-chunk size, image chunking (N > 512)."""
    from ursabench_b200.models import PreResNet
    torch.manual_seed(depth + S)
    ms = []
    for s in range(S):
        m = PreResNet(num_classes=Cc, depth=depth).cuda().eval()
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data.uniform_(0.5, 1.5)
                mod.bias.data.normal_(0, 0.2)
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 2.0)
        m.fc.weight.data.mul_(4.0)
        ms.append(m)
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms])
    bufs = torch.stack([torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]) for m in ms])
    x = torch.randn(N, 3, 32, 32, device="cuda")
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, Cc, device="cuda")
    C.bma_preresnet_forward(bank, bufs, S, x, depth, Cc, P, E, logits_out=logits,
                            algo=_algo(C, algo))
    torch.cuda.synchronize()
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        ref = torch.stack([m(x) for m in ms])
    scale = max(1.0, ref.abs().max().item())
    assert (logits - ref).abs().max().item() < 1e-4 * scale
    pbar = torch.softmax(ref.double(), -1).mean(0)                  # BMA probabilities: north star 1e-5
    assert (P.double() / S - pbar).abs().max().item() < 1e-5


def test_k3_preresnet_workspace_kept_flag(C):
    """URSA_ALGO_FLAG_WS_KEPT: a second call on the caller's untouched workspace skips re-zeroing the plane-image padding and returns
    the same bits; without the flag a scribbled workspace is re-initialised by the call itself."""
    from ursabench_b200.models import PreResNet
    torch.manual_seed(11)
    ms = [PreResNet(num_classes=10, depth=8).eval() for _ in range(3)]
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms]).cuda()
    bufs = torch.stack([torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]) for m in ms]).cuda()
    x = torch.randn(77, 3, 32, 32, device="cuda")
    outs = []
    ws = None
    for step in range(3):
        P, E = torch.zeros(77, 10, device="cuda"), torch.zeros(77, device="cuda")
        if step == 2:
            ws.fill_(float("nan"))                                        # somebody else used the memory: no flag
        algo = C.ALGO_TCGEN05_FUSED_F16 | (C.ALGO_FLAG_WS_KEPT if step == 1 else 0)
        ws = C.bma_preresnet_forward(bank, bufs, 3, x, 8, 10, P, E, algo=algo, workspace=ws)
        outs.append((P.clone(), E.clone()))
    assert torch.isfinite(outs[0][0]).all()
    for P, E in outs[1:]:
        assert torch.equal(P, outs[0][0]) and torch.equal(E, outs[0][1])


# ----------------------------------------------------------------------------- K3 forward, MLP on tcgen05 (3xTF32, 2xFP16-split)
@pytest.mark.parametrize("engine", ["ALGO_TCGEN05", "ALGO_TCGEN05_F16"])
@pytest.mark.parametrize("name,S", [("mlp", 5), ("mlp_c100", 3)])
def test_k3_mlp_tcgen05_matches_reference_golden(C, name, S, engine):
    g = _npz("prediction.npz")
    hidden, in_dim, Cc = (int(v) for v in g[name + "/arch"])
    bank = dev(g[name + "/bank"])
    x = dev(g[name + "/x"].reshape(g[name + "/x"].shape[0], -1))
    N = x.shape[0]
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, Cc, device="cuda")
    C.bma_mlp_forward(bank, S, x, in_dim, hidden, Cc, P, E, logits_out=logits, algo=getattr(C, engine))
    torch.cuda.synchronize()
    if name + "/logits" in g:
        np.testing.assert_allclose(logits.cpu().numpy(), g[name + "/logits"], atol=3e-5, rtol=1e-5)
    np.testing.assert_allclose(P.cpu().numpy(), g[name + "/ensemble_proba"], atol=1e-5, rtol=0)   # north star
    np.testing.assert_allclose(E.cpu().numpy(), g[name + "/entropy"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("hidden,S,N", [(400, 6, 1000), (200, 3, 333), (600, 2, 129), (400, 40, 700)])
def test_k3_mlp_tcgen05_vs_ffma_config1_shapes(C, hidden, S, N):
    """MLP 784-h-h-10 (config 1 and its siblings): the tensor-core paths against the fp32 CUDA-core path and a
    float64 forward -- 3xTF32 and 2xFP16-split must stay at fp32-level accuracy (a single TF32 / FP16 pass would be ~1e-3).
    S = 40 spans two sample chunks of the persistent FP16-split kernel (tiles of many samples per CTA)."""
    from ursabench_b200.models import MLP
    torch.manual_seed(hidden)
    ms = [MLP(hidden, 784, 10).cuda() for _ in range(S)]
    for m in ms:
        for p in m.parameters():
            p.data.mul_(2.0)
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms])
    x = torch.randn(N, 784, device="cuda")
    outs = {}
    for algo in (C.ALGO_FFMA, C.ALGO_TCGEN05, C.ALGO_TCGEN05_F16):
        P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
        logits = torch.empty(S, N, 10, device="cuda")
        C.bma_mlp_forward(bank, S, x, 784, hidden, 10, P, E, logits_out=logits, algo=algo)
        torch.cuda.synchronize()
        outs[algo] = (logits, P)
    with torch.no_grad():
        ref = torch.stack([m.double()(x.double()) for m in ms])
    scale = max(1.0, ref.abs().max().item())
    e_ffma = (outs[C.ALGO_FFMA][0].double() - ref).abs().max().item()
    e_tc = (outs[C.ALGO_TCGEN05][0].double() - ref).abs().max().item()
    assert e_ffma < 2e-5 * scale
    assert e_tc < 4e-5 * scale, (e_tc, e_ffma)
    e_f16 = (outs[C.ALGO_TCGEN05_F16][0].double() - ref).abs().max().item()
    assert e_f16 < 4e-5 * scale, (e_f16, e_tc, e_ffma)
    pbar_ref = torch.softmax(ref, -1).mean(0)                       # the BMA probabilities (north star: 1e-5)
    assert (outs[C.ALGO_TCGEN05][1].double() / S - pbar_ref).abs().max().item() < 5e-6
    assert (outs[C.ALGO_TCGEN05_F16][1].double() / S - pbar_ref).abs().max().item() < 5e-6
    assert (outs[C.ALGO_FFMA][1].double() / S - pbar_ref).abs().max().item() < 5e-6


def test_k3_mlp_f16_range_overflow_is_loud(C):
    """An input beyond fp16's range must come out as NaN probabilities for that image (never finite-but-wrong); the other
    images are untouched."""
    from ursabench_b200.models import MLP
    torch.manual_seed(3)
    m = MLP(200, 784, 10).cuda()
    bank = torch.cat([p.detach().reshape(-1) for p in m.parameters()])[None].contiguous()
    x = torch.randn(300, 784, device="cuda")
    x[7, 5] = 1e6
    P, E = torch.zeros(300, 10, device="cuda"), torch.zeros(300, device="cuda")
    C.bma_mlp_forward(bank, 1, x, 784, 200, 10, P, E, algo=C.ALGO_TCGEN05_F16)
    assert not torch.isfinite(P[7]).all()
    keep = torch.ones(300, dtype=torch.bool, device="cuda")
    keep[7] = False
    with torch.no_grad():
        ref = torch.softmax(m.double()(x.double()), -1)
    assert (P[keep].double() - ref[keep]).abs().max().item() < 1e-5
