"""The CPU oracle (oracle/restate.py) against the fixtures generated from the live reference
(oracle/gen_golden.py) and the reference's only known-answer artefact (notebook lr trace)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _npz(name):
    return np.load(os.path.join(GOLD, name))


def _json(name):
    return json.load(open(os.path.join(GOLD, name)))


CASES = ["sgld_wd_noise", "sgld_nowd_nonoise", "sghmc_wd_noise", "sghmc_nowd_noise", "sghmc_wd_nonoise"]


@pytest.mark.parametrize("case", CASES)
def test_sgmcmc_step_bit_exact(case):
    g = _npz("sgmcmc_step.npz")
    lr0, mom, wd, n_train, noise = g[case + "/hyper"]
    p = g[case + "/init"]
    v = None
    for t in range(4):
        p, v = R.sgmcmc_step(p, g["%s/g%d" % (case, t)], v, g["%s/z%d" % (case, t)],
                             float(g["%s/lr%d" % (case, t)]), mom, wd, int(n_train),
                             first_step=(t == 0), add_noise=bool(noise))
        ref_p = g["%s/p%d" % (case, t)]
        assert np.array_equal(p, ref_p), "step %d: %d mismatching elements" % (t, (p != ref_p).sum())
        if mom != 0:
            assert np.array_equal(v, g["%s/v%d" % (case, t)])


def test_first_step_quirk_q1():
    # p=1, g=2, wd/N=.04, mu=.9, lr=.1 -> 2.632
    p, v = R.sgmcmc_step(np.float32([1.0]), np.float32([2.0]), None, None, 0.1, 0.9, 0.04 * 10, 10, True, False)
    assert abs(float(p[0]) - 2.632) < 1e-6


def test_notebook_lr_known_answer():
    nb = _json("schedules.json")["notebook"]
    num_batch = R.csghmc_num_batch(nb["n_train"], nb["batch_size"])
    assert num_batch == 601.0
    assert len(nb["trace"]) == 220
    for epochs_run, lr in nb["trace"]:
        got = R.csghmc_lr(nb["lr_0"], epochs_run - 1, nb["last_batch_idx"], num_batch, nb["cycle_length"],
                          nb["num_cycles"])
        assert abs(got - lr) < 1e-15, (epochs_run, got, lr)


def test_csghmc_schedule_and_gates_live():
    for cfg in _json("schedules.json")["live"]:
        h = cfg["hyper"]
        nb = R.csghmc_num_batch(cfg["n_train"], cfg["batch_size"])
        assert nb == cfg["num_batch"]
        for epoch, b, lr in cfg["lrs"]:
            assert R.csghmc_lr(h["lr_0"], epoch, b, nb, h["cycle_length"], h["num_cycles"]) == lr
        for e, gate in enumerate(cfg["noise_gate_by_epochs_run"]):
            assert R.csghmc_noise_gate(e, h["cycle_length"], h["burn_in_epochs"], h["num_samples_per_cycle"]) == gate
        for e, gate in enumerate(cfg["sample_gate_by_epoch_index"]):
            assert R.csghmc_sample_gate(e + 1, h["cycle_length"], h["num_samples_per_cycle"]) == gate


def test_swa_schedule():
    s = _json("schedules.json")["swa_schedule"]
    h = s["hyper"]
    for e, lr in enumerate(s["lr"]):
        assert R.swa_schedule(e, h["burn_in_epochs"], h["lr_init"], h["swag_lr"]) == lr


@pytest.mark.parametrize("mode", ["textbook", "swa_biased", "compat_n0"])
def test_swa_collect_moments_and_ring(mode):
    g = _npz("swa_collect.npz")
    ws = g[mode + "/w"]
    D = ws.shape[1]
    mean = np.zeros(D, np.float32)
    sq = np.zeros(D, np.float32)
    ring = np.zeros((0, D), np.float32)
    n = 0
    for k in range(ws.shape[0]):
        if mode == "swa_biased":
            n += 1
        mean, sq, dev = R.swag_collect(ws[k], mean, sq, 0 if mode == "compat_n0" else n)
        if mode == "textbook":
            n += 1
        ring = R.ring_push(ring, dev, max_rank=3)
        assert np.array_equal(mean, g["%s/mean%d" % (mode, k)])
        assert np.array_equal(sq, g["%s/sq%d" % (mode, k)])
        assert np.array_equal(ring, g["%s/ring%d" % (mode, k)])
        assert ring.shape[0] == int(g["%s/rank%d" % (mode, k)][0]) == min(k + 1, 3)
    assert np.array_equal(R.swag_variance(mean, sq), g[mode + "/var"])
    assert np.array_equal(R.ring_get_space(ring), g[mode + "/space"])
    if mode == "compat_n0":
        assert (g[mode + "/var"] == np.float32(1e-30)).all() and (ring == 0).all()


def test_swag_diag_draw_formula():
    g = _npz("swa_collect.npz")
    out = R.swag_draw(g["normal/mean"], g["normal/std"] ** 2, g["normal/z"][None, :])[0]
    np.testing.assert_allclose(out, g["normal/draw"], rtol=2e-6, atol=1e-7)


def test_swag_lowrank_draw_matches_dense_formula():
    rng = np.random.RandomState(0)
    D, K, S = 513, 5, 4
    mean, var = rng.randn(D).astype(np.float32), rng.rand(D).astype(np.float32)
    ring = rng.randn(K, D).astype(np.float32)
    z1, z2 = rng.randn(S, D).astype(np.float32), rng.randn(S, K).astype(np.float32)
    out = R.swag_draw(mean, var, z1, ring, z2, max_rank=20)
    ref = mean + np.sqrt(var) * z1 + (z2 @ ring) / np.sqrt(19.0)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_swag_compat_facts():
    f = _json("swag_compat_facts.json")
    assert f["num_models_collected"] == 0 and f["all_samples_identical"] and f["sample_equals_last_iterate"]
    assert f["variance_all_clamp"] and f["ring_all_zero"] and "subspace" in f["full_cov_raises"]


def test_smoothing_entropy():
    g = _npz("metrics_edge.npz")
    assert np.array_equal(R.central_smoothing(g["smooth/in"]), g["smooth/out"])
    np.testing.assert_allclose(R.predictive_entropy(g["smooth/out"]), g["smooth/entropy"], rtol=2e-6)


@pytest.mark.parametrize("name", ["mlp", "preresnet8"])
def test_bma_accumulate_from_logits(name):
    g = _npz("prediction.npz")
    P, E = R.bma_accumulate(g[name + "/logits"])
    np.testing.assert_allclose(P, g[name + "/ensemble_proba"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(E, g[name + "/entropy"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["rand_c10", "rand_c100", "rand_c2", "edges_c4"])
def test_metrics_vs_reference(name):
    g = _npz("metrics_edge.npz")
    ref = _json("metrics_edge.json")[name]
    psum, y, S = g[name + "/proba_sum"], g[name + "/y"], int(g[name + "/S"])
    c = R.bma_counters(psum, S, y)
    m = R.metrics_from_counters(c)
    assert m["error_rate"] == pytest.approx(ref["error_rate"], abs=1e-15)
    assert m["brier_score"] == pytest.approx(ref["brier_score"], rel=1e-12)
    assert m["nll"] == pytest.approx(ref["nll"], rel=2e-6)
    assert m["ll"] == pytest.approx(ref["ll"], rel=2e-6)
    assert m["ece"] == pytest.approx(ref["ece"], abs=2e-7)      # reference averages in fp32
    pbar = psum / np.float32(S)
    assert R.get_ece(pbar, y) == ref["ece_direct"]              # line-for-line restatement: exact
    assert R.get_brier(pbar, y) == ref["brier_direct"]
    assert c["bin_count"].sum() == (c["conf"] > 0).sum()


def test_metrics_on_prediction_fixture():
    g = _npz("prediction.npz")
    ref = _json("prediction_metrics.json")
    for name, S in (("mlp", 5), ("mlp_c100", 3), ("preresnet8", 2)):
        c = R.bma_counters(g[name + "/ensemble_proba"], S, g[name + "/y"])
        m = R.metrics_from_counters(c)
        for k in ("error_rate", "nll", "ll", "brier_score", "ece"):
            assert m[k] == pytest.approx(ref[name][k], rel=3e-6, abs=2e-7), (name, k)


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    out = R.philox4x32_10(np.zeros((1, 4), np.uint32), (0, 0))[0]
    assert [hex(int(v)) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = R.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [hex(int(v)) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    out = R.philox4x32_10(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32),
                          (0xa4093822, 0x299f31d0))[0]
    assert [hex(int(v)) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_draw_normals_stream():
    # the SWAG draw's six-normals-per-block stream (oracle restatement of csrc/common.cuh::box_muller6)
    z = R.draw_normals(13, 20000, seed=77, step=3)
    assert z.shape == (13, 20000) and abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01 and abs((z ** 4).mean() - 3) < 0.06
    assert np.abs(np.corrcoef(z) - np.eye(13)).max() < 0.04                      # the six draws of a block are uncorrelated
    assert np.array_equal(R.draw_normals(6, 20000, 77, 3), z[:6]) and np.abs(z).max() < 5.9
    # bit layout: block 0 of (seed 0, step 0) is the Random123 known answer 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    x, y, zz, w = 0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8
    a = [x >> 8, ((y << 2) | (zz >> 30)) & 0xFFFFFF, ((zz << 12) | (w >> 20)) & 0xFFFFFF]
    t = [((x << 10) | (y >> 22)) & 0x3FFFF, (zz >> 12) & 0x3FFFF, (w >> 2) & 0x3FFFF]
    exp = []
    for aj, tj in zip(a, t):
        rad = np.sqrt(-2 * np.log(float(np.float32((aj + 0.5) * 2.0 ** -24))))
        th = 2 * np.pi * (tj + 0.5) * 2.0 ** -18
        exp += [rad * np.cos(th), rad * np.sin(th)]
    np.testing.assert_allclose(R.draw_normals(6, 1, 0, 0)[:, 0], exp, rtol=1e-12)


def test_philox_normals_moments():
    z, _ = R.philox_normals(200000, seed=1234, step=7)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    z2, _ = R.philox_normals(1000, seed=1234, step=7, elem_offset=500)
    assert np.array_equal(z2[:100], z[500:600])


# ---------------------------------------------------------------------------------------------------------
# oracle/port_torch.py (the timed CPU baseline) pinned to the same fixtures
@pytest.mark.parametrize("case", CASES)
def test_port_optimizer_matches_reference(case):
    import torch
    from oracle.port_torch import PortOptimSGHMC
    g = _npz("sgmcmc_step.npz")
    lr0, mom, wd, n_train, noise = g[case + "/hyper"]
    sizes = [int(s) for s in g["sizes"]]
    flat = torch.from_numpy(g[case + "/init"].copy())
    params = [t.clone().requires_grad_(True) for t in torch.split(flat, sizes)]
    opt = PortOptimSGHMC(params, lr0, mom, wd, int(n_train))
    for t in range(4):
        opt.lr = float(g["%s/lr%d" % (case, t)])
        gs = torch.split(torch.from_numpy(g["%s/g%d" % (case, t)].copy()), sizes)
        zs = list(torch.split(torch.from_numpy(g["%s/z%d" % (case, t)].copy()), sizes))
        for p, gg in zip(params, gs):
            p.grad = gg.clone()
        opt.step(add_langevin_noise=bool(noise), noise=zs)
        got = torch.cat([p.detach() for p in params]).numpy()
        assert np.array_equal(got, g["%s/p%d" % (case, t)])


def test_port_prediction_matches_reference():
    import torch
    from oracle.port_torch import port_metrics, port_prediction_update
    from ursabench_b200.models import MLP
    g = _npz("prediction.npz")
    ref = _json("prediction_metrics.json")["mlp"]
    hidden, in_dim, C = (int(v) for v in g["mlp/arch"])
    ms = []
    for row in g["mlp/bank"]:
        m = MLP(hidden, in_dim, C)
        torch.nn.utils.vector_to_parameters(torch.from_numpy(row.copy()), m.parameters())
        ms.append(m)
    x, y = torch.from_numpy(g["mlp/x"]), torch.from_numpy(g["mlp/y"])
    batches = [(x[i:i + 128], y[i:i + 128]) for i in range(0, len(x), 128)]
    P, E = port_prediction_update(ms, batches, C)
    assert np.array_equal(P.numpy(), g["mlp/ensemble_proba"])
    np.testing.assert_allclose(E.numpy(), g["mlp/entropy"], rtol=1e-6)
    m = port_metrics(P, len(ms), y)
    for k in ("error_rate", "nll", "brier_score", "ece"):
        assert m[k] == pytest.approx(ref[k], rel=1e-7, abs=1e-9)


def test_port_swa_collect_matches_reference():
    import torch
    from oracle.port_torch import port_swa_collect
    g = _npz("swa_collect.npz")
    ws = g["textbook/w"]
    mean, sq = torch.zeros(ws.shape[1]), torch.zeros(ws.shape[1])
    for k in range(ws.shape[0]):
        port_swa_collect(torch.from_numpy(ws[k].copy()), mean, sq, k)
        assert np.array_equal(mean.numpy(), g["textbook/mean%d" % k])
        assert np.array_equal(sq.numpy(), g["textbook/sq%d" % k])


def test_model_layouts_match_reference():
    """our models.py keeps the reference's parameter order / shapes / buffer order (flat layout = util.flatten)."""
    import warnings
    from ursabench_b200 import models
    lay = _json("layouts.json")
    mk = {"MLP400_c10": lambda: models.MLP(400, 784, 10),
          "PreResNet20_c10": lambda: models.PreResNet(num_classes=10, depth=20),
          "PreResNet8_c10": lambda: models.PreResNet(num_classes=10, depth=8),
          "WRN28x10_c100": lambda: models.WideResNet(num_classes=100, depth=28, widen_factor=10)}
    for k, f in mk.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = f()
        assert [[n, list(p.shape)] for n, p in m.named_parameters()] == lay[k]["params"], k
        assert [[n, list(b.shape), str(b.dtype)] for n, b in m.named_buffers()] == lay[k]["buffers"], k
        assert sum(p.numel() for p in m.parameters()) == lay[k]["D"]


def test_model_forward_matches_reference_golden():
    """same weights -> same logits as the reference's PreResNet / MLP (pins models.py and the flat layout)."""
    import torch
    from ursabench_b200 import models
    g = _npz("prediction.npz")
    m = models.PreResNet(num_classes=10, depth=8).eval()
    x = torch.from_numpy(g["preresnet8/x"].astype(np.float32))
    for s in range(2):
        torch.nn.utils.vector_to_parameters(torch.from_numpy(g["preresnet8/bank"][s].copy()), m.parameters())
        off = 0
        buf = torch.from_numpy(g["preresnet8/buffers"][s].copy())
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.copy_(buf[off:off + b.numel()].view(b.shape))
                off += b.numel()
        with torch.no_grad():
            out = m(x).numpy()
        np.testing.assert_allclose(out, g["preresnet8/logits"][s], rtol=1e-5, atol=1e-5)


# --------------------------------------------------------------------------- HMC host logic (a15)
def test_hmc_kept_iterations_equals_reference_slice():
    """`kept_iterations` must select exactly what the reference wrapper's samples[burn*L::L] (inference/hmc.py:80)
    selects from hamiltorch's list [init] + L positions per iteration."""
    from ursabench_b200.inference.hmc import kept_iterations
    for n in (1, 3, 10):
        for L in (1, 4, 7):
            lst = [("init", 0, 0)] + [("pos", it, k) for it in range(1, n + 1) for k in range(L)]
            for burn in (-12, -3, -1, 0, 1, 2, n, n + 1):
                want = lst[burn * L::L]
                first, use_first = kept_iterations(n, L, burn)
                got = []
                if first == 0 and not use_first:
                    got.append(("init", 0, 0))
                got += [("pos", it, 0 if use_first else L - 1) for it in range(max(first, 1), n + 1)]
                assert got == want, (n, L, burn)


def test_hmc_restatement_samples_gaussian_prior():
    """With no data term the target is the prior N(0, 1/tau): the restated chain must reproduce its variance."""
    import math
    rng = np.random.RandomState(0)
    D, tau, n_it = 40, 4.0, 300
    zs = [rng.randn(D).astype(np.float32) for _ in range(n_it)]
    lu = [math.log(rng.rand()) for _ in range(n_it)]
    ret, acc = R.hmc_chain(np.zeros(D, np.float32), lambda t: (0.0, np.zeros_like(t)), zs, lu, 0.15, 10, tau, 1.0)
    assert len(ret) == n_it * 10 + 1 and np.mean(acc) > 0.9
    s = np.stack(ret[50 * 10::10])
    assert abs(s.var() - 1 / tau) < 0.03


def test_decision_cost_matrices_match_reference():
    """tasks/decision_making.py:12-51: the three cost matrices, dumped from the live reference."""
    from ursabench_b200.tasks import decision_making as dm
    g = np.load(os.path.join(GOLD, "ood_decision.npz"))
    for name, fn, C in (("MNIST", dm.MNIST_cost, 10), ("CIFAR10", dm.CIFAR10_cost, 10), ("CIFAR100", dm.CIFAR100_cost, 100)):
        assert np.array_equal(fn(C).numpy(), g["cost/" + name]), name
    y, D = torch.tensor([3, 0, 7]), torch.tensor([3, 1, 0])
    assert float(dm.decision_cost(D, y, dm.MNIST_cost(10))) == pytest.approx(100.1)


def test_ood_golden_is_self_consistent():
    """The smoothed-probability sum the reference accumulates is the affine image of the unsmoothed one -- the identity
    OODDetection / Decision rely on (sum_s p~_s = (1-g) sum_s p_s + S g / C): rows sum to S."""
    g = np.load(os.path.join(GOLD, "ood_decision.npz"))
    np.testing.assert_allclose(g["ood/in_distribution_ensemble_proba"].sum(1), 4.0, rtol=2e-6)
    np.testing.assert_allclose(g["decision/ensemble_proba"].sum(1), 3.0, rtol=2e-6)
    np.testing.assert_allclose(g["decision/risk"], g["decision/ensemble_proba"] @ g["decision/cost_mat"], rtol=2e-5, atol=1e-4)
