"""GPU tests of the drop-in class API (inference/*, tasks/Prediction) against the reference's golden fixtures, the
torch-CPU port (oracle/port_torch.py) and the call shapes of the reference's drivers
(hyperopt/hyper_optimization.py:51-73, experiment.py:172-179, time_script.py:100-113)."""
import copy
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import port_torch as PT
from oracle import restate as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = torch.device("cuda")


def _npz(name):
    return np.load(os.path.join(GOLD, name))


def _json(name):
    return json.load(open(os.path.join(GOLD, name)))


@pytest.fixture(scope="module")
def U():
    import ursabench_b200
    return ursabench_b200


def _flat(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def _toy(n=256, d=20, c=3, bs=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 1, 4, 5, generator=g)
    w = torch.randn(d, c, generator=g)
    y = (x.view(n, -1) @ w).argmax(1)
    ds = torch.utils.data.TensorDataset(x, y)
    return ds, torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=False)


# ------------------------------------------------------------------------------------------ optimizer class
CASES = ["sgld_wd_noise", "sgld_nowd_nonoise", "sghmc_wd_noise", "sghmc_nowd_noise", "sghmc_wd_nonoise"]


@pytest.mark.parametrize("case", CASES)
def test_optimizer_class_matches_reference_golden(U, case):
    """the reference protocol verbatim: ragged parameter tensors, ``p.grad`` assigned directly, lr moved through
    ``param_groups`` between steps, identical noise -> bit-identical parameters and momentum buffers."""
    g = _npz("sgmcmc_step.npz")
    lr0, mom, wd, n_train, noise = g[case + "/hyper"]
    shapes = [(7,), (13, 5), (64,), (3, 3, 3, 4), (1,), (257,), (31, 33)]
    sizes = [int(s) for s in g["sizes"]]
    init = torch.from_numpy(g[case + "/init"].copy())
    params = [torch.nn.Parameter(t.clone().view(s).to(DEV)) for t, s in zip(torch.split(init, sizes), shapes)]
    opt = U.inference.optimSGHMC(params, lr=lr0, momentum=mom, num_training_samples=int(n_train), weight_decay=wd)
    for t in range(4):
        for grp in opt.param_groups:
            grp["lr"] = float(g["%s/lr%d" % (case, t)])
        gt = torch.from_numpy(g["%s/g%d" % (case, t)].copy())
        for p, gg in zip(params, torch.split(gt, sizes)):
            p.grad = gg.clone().view_as(p).to(DEV)
        z = torch.zeros(opt.flat.ld, device=DEV)
        z[:opt.flat.D] = torch.from_numpy(g["%s/z%d" % (case, t)].copy()).to(DEV)
        opt.step(add_langevin_noise=bool(noise), noise=z if noise else None)
        got = torch.cat([p.detach().reshape(-1) for p in params]).cpu().numpy()
        assert np.array_equal(got, g["%s/p%d" % (case, t)])
        if mom != 0:
            v = torch.cat([opt.state[p]["momentum_buffer"].reshape(-1) for p in params]).cpu().numpy()
            assert np.array_equal(v, g["%s/v%d" % (case, t)])
    assert opt.launches == 4                       # ONE kernel launch per step for all 7 tensors


def test_optimizer_argument_errors(U):
    p = [torch.nn.Parameter(torch.zeros(4, device=DEV))]
    with pytest.raises(ValueError, match="Invalid learning rate"):
        U.inference.optimSGHMC(p, lr=-1.0)
    with pytest.raises(ValueError, match="Invalid momentum"):
        U.inference.optimSGHMC(p, lr=0.1, momentum=-0.1)
    with pytest.raises(ValueError, match="Invalid weight_decay"):
        U.inference.optimSGHMC(p, lr=0.1, weight_decay=-1.0)
    with pytest.raises(ValueError):
        U.inference.optimSGHMC([torch.nn.Parameter(torch.zeros(4))], lr=0.1)       # CPU parameter: no CPU path


# ------------------------------------------------------------------------------------------ samplers
def _deterministic_trajectory(U, cls_name, hyp, epochs_to_run):
    """GPU sampler vs. the torch-CPU port with the Langevin noise gated off: weights after the same number of
    steps must agree (fp32 fwd/bwd on different hardware -> 1e-4)."""
    ds, loader = _toy()
    torch.manual_seed(1)
    model = U.models.MLP(16, 20, 3)
    ref_model = copy.deepcopy(model)
    inf = getattr(U.inference, cls_name)(dict(hyp), model.to(DEV), loader, device=DEV)
    return inf, ref_model, loader, ds


def test_csghmc_schedule_and_trajectory_vs_port(U, capsys):
    hyp = {"lr_0": 0.3, "prior_std": 1.0, "num_samples_per_cycle": 2, "cycle_length": 5, "burn_in_epochs": 1,
           "num_cycles": 2, "alpha": 0.3}
    inf, ref_model, loader, ds = _deterministic_trajectory(U, "cSGHMC", hyp, 2)
    # epochs 0,1 of the cycle are noise-free: (e % 5) + 1 > 5 - 1 - 2 = 2  is False for e = 0, 1
    noise_flags = []
    orig_step = inf.optimizer.step

    def spy(add_langevin_noise=True, **kw):
        noise_flags.append(add_langevin_noise)
        return orig_step(add_langevin_noise=add_langevin_noise, **kw)
    inf.optimizer.step = spy
    inf._run_epoch(lambda b: inf._noise_gate(), lr_for_batch=lambda b: inf._adjust_learning_rate(inf.optimizer, inf.epochs_run, b))
    inf.epochs_run += 1
    inf._run_epoch(lambda b: inf._noise_gate(), lr_for_batch=lambda b: inf._adjust_learning_rate(inf.optimizer, inf.epochs_run, b))
    inf.epochs_run += 1
    assert not any(noise_flags)
    batches = [(x, y) for x, y in loader]
    nb = R.csghmc_num_batch(len(ds), 32)
    opt = PT.PortOptimSGHMC(ref_model.parameters(), hyp["lr_0"], 1 - hyp["alpha"], 1.0, len(ds))
    ref = PT.port_sgmcmc_epochs(ref_model, batches, opt, 2, noise_fn=lambda e, b: False,
                                lr_fn=lambda e, b: PT.port_csghmc_lr(hyp["lr_0"], e, b, nb, 5, 2))
    got = _flat(inf.model).cpu()
    assert (got - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())
    # schedule: host float64, identical to the reference formula
    assert inf.num_batch == nb
    assert inf._adjust_learning_rate(inf.optimizer, 3, 4) == R.csghmc_lr(hyp["lr_0"], 3, 4, nb, 5, 2)


def test_csghmc_sample_gates_and_bank(U, capsys):
    hyp = {"lr_0": 0.05, "prior_std": 1.0, "num_samples_per_cycle": 2, "cycle_length": 4, "burn_in_epochs": 1,
           "num_cycles": 2, "alpha": 0.5}
    ds, loader = _toy()
    torch.manual_seed(2)
    inf = U.inference.cSGHMC(hyp, U.models.MLP(16, 20, 3).to(DEV), loader, device=DEV)
    samples = inf.sample()
    assert len(samples) == 4 and inf.epochs_run == 8           # samples at epochs 3,4 of each 4-epoch cycle
    out = capsys.readouterr().out
    assert out.count("Epoch: ") == 8
    assert inf.optimizer.launches == 8 * len(loader)           # exactly one K1 launch per step
    # the last handle equals the live weights; handles are CPU modules like the reference's deep copies
    last = samples[-1]
    assert isinstance(last, torch.nn.Module)
    assert torch.equal(_flat(last), _flat(inf.model).cpu())
    assert next(last.parameters()).device.type == "cpu"
    assert not torch.equal(_flat(samples[0]), _flat(samples[1]))
    x = torch.randn(5, 1, 4, 5)
    live = inf.model.eval()(x.to(DEV)).cpu()
    assert torch.allclose(last.eval()(x), live, atol=1e-5)
    with pytest.raises(AssertionError):
        U.inference.cSGHMC({**hyp, "cycle_length": 3}, U.models.MLP(16, 20, 3).to(DEV), loader, device=DEV)


@pytest.mark.parametrize("cls_name", ["SGLD", "SGHMC"])
def test_sghmc_sgld_api_and_scheduler(U, cls_name):
    hyp = {"lr": 0.05, "prior_std": 1.0, "num_samples": 3, "alpha": 0.4, "burn_in_epochs": 2}
    ds, loader = _toy()
    torch.manual_seed(3)
    inf = getattr(U.inference, cls_name)(hyp, U.models.MLP(16, 20, 3).to(DEV), loader, device=DEV)
    if cls_name == "SGLD":
        assert hyp["alpha"] == 1.0 and inf.optimizer.param_groups[0]["momentum"] == 0      # sgld.py:22
    samples = inf.sample()
    assert len(samples) == 3 and inf.burnt_in is True
    assert inf.optimizer.launches == (2 + 1 + 2) * len(loader)       # burn-in + 1, then 1 epoch per sample
    # CosineAnnealingLR(T_max = burn_in + num_samples) stepped once per epoch (sghmc.py:44-45,87)
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=hyp["lr"])
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(ref_opt, T_max=5)
    for _ in range(5):
        ref_opt.step()
        sch.step()
    assert inf.optimizer.param_groups[0]["lr"] == pytest.approx(ref_opt.param_groups[0]["lr"], rel=1e-9, abs=1e-12)
    # update_hyp re-initialises the model in place and keeps the flat views attached (hyper_optimization.py:55)
    before = _flat(inf.model).clone()
    inf.update_hyp({"lr": 0.01, "prior_std": 2.0, "num_samples": 2, "alpha": 0.4, "burn_in_epochs": 0})
    assert not torch.equal(before, _flat(inf.model)) and inf.flat.is_attached(inf.model)
    assert inf.burnt_in is False and len(inf.sample()) == 2


def test_sgld_learns_and_noise_has_reference_scale(U):
    """SGLD on a separable toy problem: the loss goes down, and the per-step noise std is sqrt(2 lr)/N (Q4)."""
    ds, loader = _toy(n=512)
    torch.manual_seed(4)
    model = U.models.MLP(16, 20, 3).to(DEV)
    hyp = {"lr": 0.5, "prior_std": 10.0, "num_samples": 1, "alpha": 1.0, "burn_in_epochs": 15}
    inf = U.inference.SGLD(hyp, model, loader, device=DEV)
    l0 = inf.compute_val_loss(loader)
    inf.sample()
    assert inf.compute_val_loss(loader) < 0.6 * l0
    opt = inf.optimizer
    p0 = opt.flat.p.clone()
    opt.flat.g.zero_()
    opt.param_groups[0]["weight_decay"] = 0.0
    opt.param_groups[0]["lr"] = 0.1                            # the cosine schedule has annealed lr to ~0 by now
    opt.step(add_langevin_noise=True)
    delta = (opt.flat.p - p0)[:opt.flat.D]
    expect = math.sqrt(2 * opt.param_groups[0]["lr"]) / len(ds)
    assert abs(delta.std().item() / expect - 1) < 0.1


# ------------------------------------------------------------------------------------------ SWAG
def _swag_hyp(**kw):
    h = {"swag_lr": 0.02, "swag_wd": 0.0, "lr_init": 0.05, "num_samples": 6, "momentum": 0.5, "burn_in_epochs": 2,
         "num_iterates": 5, "subspace_type": "covariance"}
    h.update(kw)
    return h


def test_swag_reference_compat_reproduces_quirks(U):
    facts = _json("swag_compat_facts.json")
    ds, loader = _toy(n=64, bs=16)
    torch.manual_seed(0)
    swag = U.inference.SWAG(_swag_hyp(reference_compat=True, num_samples=3), U.models.MLP(8, 20, 3).to(DEV), loader,
                            device=DEV)
    samples = swag.sample()
    flats = [_flat(m) for m in samples]
    assert len(samples) == facts["num_returned"]
    assert int(swag.num_models_collected.item()) == facts["num_models_collected"] == 0
    assert all(torch.equal(flats[0], f) for f in flats) == facts["all_samples_identical"]
    assert torch.equal(flats[0], _flat(swag.model).cpu()) == facts["sample_equals_last_iterate"]
    assert bool((swag.weight_variance == 1e-30).all()) == facts["variance_all_clamp"]
    assert bool((swag.subspace.cov_mat_sqrt == 0).all()) == facts["ring_all_zero"]
    with pytest.raises(AttributeError, match="subspace"):
        swag.sample_iterative(full_cov=True)


def test_swag_textbook_moments_and_draws(U):
    ds, loader = _toy(n=128, bs=32)
    torch.manual_seed(5)
    swag = U.inference.SWAG(_swag_hyp(num_samples=32), U.models.MLP(8, 20, 3).to(DEV), loader, device=DEV, max_rank=4)
    iterates = []
    orig = swag._collect_model

    def spy():
        iterates.append(_flat(swag.model).clone())
        orig()
    swag._collect_model = spy
    from ursabench_b200 import _C
    seen = {}
    orig_draw = _C.swag_draw

    def draw_spy(out, mean_, var_, D_, **kw):
        seen.update(kw, ld=out.shape[1])
        return orig_draw(out, mean_, var_, D_, **kw)
    _C.swag_draw = draw_spy
    try:
        samples = swag.sample(full_cov=True)
    finally:
        _C.swag_draw = orig_draw
    W = torch.stack(iterates)                                   # [5, D]
    assert int(swag.num_models_collected.item()) == 5 and len(samples) == 32
    # moments == CPU port run on the same iterates (bit-exact: same op order)
    mean, sq = torch.zeros(W.shape[1]), torch.zeros(W.shape[1])
    devs = []
    for k in range(5):
        devs.append(PT.port_swa_collect(W[k].cpu(), mean, sq, k).clone())
    assert torch.equal(swag.weight_mean.cpu(), mean) and torch.equal(swag.sq_mean.cpu(), sq)
    assert torch.equal(swag.subspace.cov_mat_sqrt.cpu(), torch.stack(devs[-4:]))       # ring: last max_rank rows
    assert int(swag.subspace.rank.item()) == 4
    var = torch.clamp(sq - mean ** 2, 1e-30)
    assert torch.equal(swag.weight_variance.cpu(), var)
    # draws == mean + sqrt(var) z1 + D^T z2 / sqrt(max_rank - 1) with the z2 the class drew and the documented
    # Philox z1 stream (swag.py:88-97 without the :98 overwrite)
    draws = torch.stack([_flat(m) for m in samples])            # [32, D] (CPU handles)
    D = W.shape[1]
    z1 = _C.swag_draw_noise(32, D, seen["seed"], seen["step"], DEV)[:, :D].cpu().numpy()
    expect = R.swag_draw(mean.numpy(), var.numpy(), z1, seen["ring"][:, :D].cpu().numpy(), seen["z2"].cpu().numpy(),
                         max_rank=4)
    np.testing.assert_allclose(draws.numpy(), expect, rtol=2e-5, atol=2e-6)
    assert len({tuple(d[:8].tolist()) for d in draws}) == 32    # all distinct


def test_swag_batched_draw_uses_one_pass(U, monkeypatch):
    ds, loader = _toy(n=64, bs=32)
    swag = U.inference.SWAG(_swag_hyp(num_samples=30, burn_in_epochs=1, num_iterates=2), U.models.MLP(8, 20, 3).to(DEV),
                            loader, device=DEV)
    calls = []
    from ursabench_b200 import _C
    orig = _C.swag_draw
    monkeypatch.setattr(_C, "swag_draw", lambda *a, **k: (calls.append(a[0].shape[0]), orig(*a, **k))[1])
    swag.sample(full_cov=True)
    assert calls == [30]                                        # 30 draws, ONE kernel launch over the ring


# ------------------------------------------------------------------------------------------ Prediction
def _mlp_models_from_golden(U, name):
    g = _npz("prediction.npz")
    hidden, in_dim, C = (int(v) for v in g[name + "/arch"])
    ms = []
    for row in g[name + "/bank"]:
        m = U.models.MLP(hidden, in_dim, C)
        torch.nn.utils.vector_to_parameters(torch.from_numpy(row.copy()), m.parameters())
        ms.append(m)
    x, y = torch.from_numpy(g[name + "/x"]), torch.from_numpy(g[name + "/y"])
    return g, ms, x, y, C


@pytest.mark.parametrize("engine", ["auto", "ffma", "generic"])
def test_prediction_matches_reference_golden_mlp(U, engine):
    g, ms, x, y, C = _mlp_models_from_golden(U, "mlp")
    ref = _json("prediction_metrics.json")["mlp"]
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=128, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, C, DEV, "ALL", engine=engine)
    assert torch.equal(task.targets, y)
    task.update_statistics(ms[:3], output_performance=False)            # accumulates across calls (:38-48)
    task.update_statistics(ms[3:], output_performance=False)
    assert task.num_samples_collected == 5
    assert task.last_engine == ("generic" if engine == "generic" else "fused_mlp")
    if engine != "generic":                                             # the product path is the tcgen05 kernel
        assert task.last_algo == (U._C.ALGO_TCGEN05_F16 if engine == "auto" else U._C.ALGO_FFMA)
    np.testing.assert_allclose(task.ensemble_proba.numpy(), g["mlp/ensemble_proba"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(task.expected_data_uncertainty.numpy(), g["mlp/entropy"], atol=1e-5, rtol=1e-5)
    m = task.get_performance_metrics()
    assert set(m) == set(ref)
    for k, v in ref.items():
        tol = 1e-9 if k == "error_rate" else 2e-5
        assert m[k] == pytest.approx(v, abs=tol, rel=2e-5), k
    assert all(next(mm.parameters()).device.type == "cpu" for mm in ms)   # callers' modules end where they started


def test_prediction_single_module_and_output_performance(U):
    g, ms, x, y, C = _mlp_models_from_golden(U, "mlp")
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=128, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, C, DEV, ["ll"])
    ll = task.update_statistics(ms[0], output_performance=True)
    assert isinstance(ll, float)
    assert ll == pytest.approx(_json("prediction_metrics.json")["mlp_single_ll"], rel=2e-5)
    task2 = U.tasks.Prediction({"in_distribution_test": loader}, C, DEV, ["ll", "ece"])
    with pytest.raises(RuntimeError, match="Multiple metrics"):
        task2.update_statistics(ms[0], output_performance=True)
    with pytest.raises(NotImplementedError):
        task2.update_statistics([1, 2, 3])
    with pytest.raises(AssertionError):
        U.tasks.Prediction({"in_distribution_test": loader}, C, DEV, ["accuracy"])
    task.reset()
    assert task.num_samples_collected == 0 and not task.ensemble_proba.any()
    assert task.expected_data_uncertainty.any()                            # reference reset() keeps it (Q10)


def test_prediction_preresnet8_generic_matches_reference_golden(U):
    g = _npz("prediction.npz")
    ref = _json("prediction_metrics.json")["preresnet8"]
    ms = []
    for s in range(2):
        m = U.models.PreResNet(num_classes=10, depth=8)
        torch.nn.utils.vector_to_parameters(torch.from_numpy(g["preresnet8/bank"][s].copy()), m.parameters())
        off, buf = 0, torch.from_numpy(g["preresnet8/buffers"][s].copy())
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.copy_(buf[off:off + b.numel()].view(b.shape))
                off += b.numel()
        ms.append(m)
    x, y = torch.from_numpy(g["preresnet8/x"].astype(np.float32)), torch.from_numpy(g["preresnet8/y"])
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 10, DEV, ["error_rate", "nll", "brier_score", "ece"])
    task.update_statistics(ms, output_performance=False)
    np.testing.assert_allclose(task.ensemble_proba.numpy(), g["preresnet8/ensemble_proba"], atol=1e-5, rtol=0)
    m = task.get_performance_metrics()
    for k in m:
        assert m[k] == pytest.approx(ref[k], abs=2e-5, rel=2e-5), k


def test_hyperopt_shaped_call_sequence(U, capsys):
    """hyperopt/hyper_optimization.py:51-73: update_hyp -> reset -> sample -> update_statistics(output_performance)."""
    ds, loader = _toy(n=256)
    tds, tloader = _toy(n=100, seed=9)
    torch.manual_seed(6)
    inf = U.inference.cSGLD(None, U.models.MLP(16, 20, 3).to(DEV), loader, device=DEV)
    task = U.tasks.Prediction({"in_distribution_test": tloader}, 3, DEV, ["ll"])
    objs = []
    for lr in (0.2, 0.4):
        hyp = {"lr_0": lr, "prior_std": 5.0, "num_samples_per_cycle": 2, "cycle_length": 4, "burn_in_epochs": 1,
               "num_cycles": 1}
        inf.update_hyp(hyp)
        task.reset()
        samples = U.util.silent(inf.sample)()
        obj = task.update_statistics(samples, output_performance=True)
        assert isinstance(obj, float) and obj < 0 and task.last_engine == "fused_mlp"
        assert task.num_samples_collected == 2
        objs.append(obj)
    assert capsys.readouterr().out.count("Epoch") == 0              # util.silent swallowed the per-epoch prints
    # the bank fast path and the reference-style path (materialised CPU modules) agree
    task.reset()
    a = task.update_statistics(samples, output_performance=True)
    plain = [copy.deepcopy(s.materialize()) for s in samples]
    task.reset()
    b = task.update_statistics(plain, output_performance=True)
    assert a == pytest.approx(b, rel=1e-6)


def test_prediction_preresnet8_fused_matches_reference_golden(U):
    g = _npz("prediction.npz")
    ref = _json("prediction_metrics.json")["preresnet8"]
    ms = []
    for s in range(2):
        m = U.models.PreResNet(num_classes=10, depth=8)
        torch.nn.utils.vector_to_parameters(torch.from_numpy(g["preresnet8/bank"][s].copy()), m.parameters())
        off, buf = 0, torch.from_numpy(g["preresnet8/buffers"][s].copy())
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.copy_(buf[off:off + b.numel()].view(b.shape))
                off += b.numel()
        ms.append(m)
    x, y = torch.from_numpy(g["preresnet8/x"].astype(np.float32)), torch.from_numpy(g["preresnet8/y"])
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 10, DEV, ["error_rate", "nll", "brier_score", "ece"])
    task.update_statistics(ms, output_performance=False)
    assert task.last_engine == "fused_preresnet" and task.last_algo == U._C.ALGO_TCGEN05_FUSED_F16
    np.testing.assert_allclose(task.ensemble_proba.numpy(), g["preresnet8/ensemble_proba"], atol=1e-5, rtol=0)
    m = task.get_performance_metrics()
    for k in m:
        assert m[k] == pytest.approx(ref[k], abs=2e-5, rel=2e-5), k


def test_prediction_fp16_range_overflow_is_loud_and_falls_back(U):
    """Activations beyond the FP16-split engine's range (|act| / 16 > 65504) must never produce finite wrong numbers:
    the kernel's logits go NaN and Prediction redoes the call on the TF32 engine."""
    torch.manual_seed(3)
    ms = [U.models.PreResNet(num_classes=10, depth=8).eval() for _ in range(2)]
    x = torch.randn(40, 3, 32, 32) * 2e7
    y = torch.randint(0, 10, (40,))
    # the raw kernel: NaN, not garbage
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms]).to(DEV)
    bufs = torch.stack([torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]) for m in ms]).to(DEV)
    P, E = torch.zeros(40, 10, device=DEV), torch.zeros(40, device=DEV)
    U._C.bma_preresnet_forward(bank, bufs, 2, x.to(DEV), 8, 10, P, E, algo=U._C.ALGO_TCGEN05_FUSED_F16)
    assert not bool(torch.isfinite(P).all())
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 10, DEV, ["error_rate"])
    task.update_statistics(ms, output_performance=False)
    assert task.last_algo == U._C.ALGO_TCGEN05_FUSED
    proba = task.ensemble_proba
    assert bool(torch.isfinite(proba).all())
    with torch.no_grad():
        ref = torch.stack([torch.softmax(m.double()(x.double()), -1) for m in ms]).sum(0)
    assert (proba.argmax(1) == ref.argmax(1)).float().mean().item() >= 0.95


def test_prediction_mlp_fp16_range_overflow_falls_back(U):
    """The MLP product path (2xFP16-split) sees an input beyond fp16's range: NaN, never finite-but-wrong, and Prediction
    redoes the evaluation on the 3xTF32 engine."""
    torch.manual_seed(5)
    ms = [U.models.MLP(200, 784, 10).eval() for _ in range(3)]
    x = torch.randn(64, 1, 28, 28)
    x[3, 0, 2, 2] = 3e5
    y = torch.randint(0, 10, (64,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=16, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 10, DEV, ["error_rate"])
    task.update_statistics(ms, output_performance=False)
    assert task.last_engine == "fused_mlp" and task.last_algo == U._C.ALGO_TCGEN05
    with torch.no_grad():
        ref = torch.stack([torch.softmax(m.double()(x.double().view(64, -1)), -1) for m in ms]).sum(0)
    assert (task.ensemble_proba.double() - ref).abs().max().item() < 1e-5
    # in range: the FP16-split engine answers
    x[3, 0, 2, 2] = 1.0
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=16, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 10, DEV, ["error_rate"])
    task.update_statistics(ms, output_performance=False)
    assert task.last_algo == U._C.ALGO_TCGEN05_F16


def test_cuda_graph_step_matches_eager(U, capsys):
    """forward + backward + K1 captured in one CUDA graph follows the same trajectory as eager launches
    (noise gated off through the device-resident scalars; lr changes every step)."""
    hyp = {"lr_0": 0.3, "prior_std": 1.0, "num_samples_per_cycle": 1, "cycle_length": 6, "burn_in_epochs": 0,
           "num_cycles": 1, "alpha": 0.3}
    ds, loader = _toy(n=256, bs=32)
    flats = []
    for graphed in (False, True):
        torch.manual_seed(11)
        inf = U.inference.cSGHMC(dict(hyp), U.models.MLP(16, 20, 3).to(DEV), loader, device=DEV)
        if graphed:
            x0, y0 = next(iter(loader))
            g = inf.enable_cuda_graph(x0, y0)
        for e in range(2):
            for b, (x, y) in enumerate(loader):
                inf._adjust_learning_rate(inf.optimizer, e, b)
                inf.train_step(x, y, add_langevin_noise=False)
        flats.append(_flat(inf.model).clone())
        if graphed:
            assert g.replays == 2 * len(loader) - 1            # all but the momentum-initialising first step
    assert (flats[0] - flats[1]).abs().max().item() < 1e-6
    # with noise on, the graphed sampler draws the same Philox stream as the eager one (same seed / step counter)
    outs = []
    for graphed in (False, True):
        torch.manual_seed(12)
        inf = U.inference.cSGHMC(dict(hyp), U.models.MLP(16, 20, 3).to(DEV), loader, device=DEV)
        if graphed:
            x0, y0 = next(iter(loader))
            inf.enable_cuda_graph(x0, y0)
        for b, (x, y) in enumerate(loader):
            inf._adjust_learning_rate(inf.optimizer, 0, b)
            inf.train_step(x, y, add_langevin_noise=True)
        outs.append(_flat(inf.model).clone())
    assert (outs[0] - outs[1]).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------ OODDetection / Decision
def _mlp_bank_models(U, g, key):
    hidden, in_dim, C = (int(v) for v in g[key + "/arch"])
    ms = []
    for row in g[key + "/bank"]:
        m = U.models.MLP(hidden, in_dim, C)
        torch.nn.utils.vector_to_parameters(torch.from_numpy(row.copy()), m.parameters())
        ms.append(m)
    return ms, C


@pytest.mark.parametrize("engine", ["auto", "generic"])
def test_ood_detection_matches_reference_golden(U, engine):
    """tasks/ood_detection.py:39-130 run by the live reference on the same weights (oracle/gen_golden.py::gen_ood_decision)."""
    g, ref = _npz("ood_decision.npz"), _json("ood_decision.json")
    ms, C = _mlp_bank_models(U, g, "ood")
    xin, xout = torch.from_numpy(g["ood/x_in"]), torch.from_numpy(g["ood/x_out"])
    lin = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xin, torch.zeros(len(xin), dtype=torch.long)), batch_size=128)
    lout = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xout, torch.zeros(len(xout), dtype=torch.long)), batch_size=64)
    task = U.tasks.OODDetection({"in_distribution_test": lin, "out_distribution_test": lout}, C, DEV, engine=engine)
    task.update_statistics(ms[:1], output_performance=False)            # accumulates across calls
    m = task.update_statistics(ms[1:], output_performance=True)
    assert task.num_samples_collected == 4
    assert task.last_engine == ("fused_mlp" if engine == "auto" else "generic")
    for k in ("in_distribution_ensemble_proba", "out_distribution_ensemble_proba"):
        np.testing.assert_allclose(getattr(task, k).numpy(), g["ood/" + k], atol=1e-5, rtol=0)
    for k in ("in_distribution_data_uncertainty", "out_distribution_data_uncertainty", "in_distribution_total_uncertainty",
              "out_distribution_total_uncertainty"):
        np.testing.assert_allclose(getattr(task, k).numpy(), g["ood/" + k], atol=2e-5, rtol=1e-5)
    for k in ("in_distribution_model_uncertainty", "out_distribution_model_uncertainty"):
        # total - data uncertainty: the difference of two O(1) fp32 entropies, each good to ~1e-5 (the fp32 forward's summation
        # order -- cuBLAS heuristics in the generic engine -- moves single elements by 2e-5)
        np.testing.assert_allclose(getattr(task, k).numpy(), g["ood/" + k], atol=5e-5, rtol=0)
    assert set(m) == set(ref["ood"])
    for k, v in ref["ood"].items():
        assert m[k] == pytest.approx(v, abs=2e-3), k                    # rank statistic of N = 470 scores
    task.reset()
    assert task.num_samples_collected == 0 and not task.in_distribution_data_uncertainty.any()
    m1 = task.update_statistics(ms[0])                                  # single module, default output_performance=True
    for k, v in ref["ood_single"].items():
        assert m1[k] == pytest.approx(v, abs=2e-3), k
    with pytest.raises(NotImplementedError):
        task.update_statistics([1, 2])


def test_decision_matches_reference_golden(U):
    """tasks/decision_making.py:83-152 run by the live reference on a torchvision MNIST object around synthetic images."""
    import torchvision
    g, ref = _npz("ood_decision.npz"), _json("ood_decision.json")
    ms, C = _mlp_bank_models(U, g, "decision")
    ds = torchvision.datasets.MNIST.__new__(torchvision.datasets.MNIST)
    ds.data, ds.targets = torch.from_numpy(g["decision/data_u8"]), torch.from_numpy(g["decision/targets"])
    ds.transform, ds.target_transform = torchvision.transforms.ToTensor(), None
    loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False)
    task = U.tasks.Decision({"decision_data_test": loader}, C, DEV)
    assert np.array_equal(task.cost_mat.numpy(), g["decision/cost_mat"])
    assert torch.equal(task.targets, ds.targets)
    task.update_statistics(ms[:2], output_performance=False)
    res = task.update_statistics(ms[2:], output_performance=True)
    assert task.num_samples_collected == ref["decision"]["num_samples_collected"] and task.last_engine == "fused_mlp"
    np.testing.assert_allclose(task.ensemble_proba.numpy(), g["decision/ensemble_proba"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(res["Pred_cost"].numpy(), g["decision/risk"], atol=2e-4, rtol=2e-5)
    assert np.array_equal(res["Decision"].numpy(), g["decision/D"])
    assert float(res["True_Cost"]) == pytest.approx(ref["decision"]["True_Cost"], rel=1e-6)
    task.reset()
    assert task.num_samples_collected == 0 and not task.ensemble_proba.any()
    # datasets the reference has no cost matrix for raise like the reference; cost_mat= is the documented extension
    tl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.randn(8, 784), torch.zeros(8, dtype=torch.long)), batch_size=4)
    with pytest.raises(NotImplementedError):
        U.tasks.Decision({"decision_data_test": tl}, C, DEV)
    t2 = U.tasks.Decision({"decision_data_test": tl}, C, DEV, cost_mat=U.tasks.decision_making.CIFAR10_cost(C))
    out = t2.update_statistics(ms[0])
    assert out["Decision"].shape == (8,) and out["Pred_cost"].shape == (8, C)


# ------------------------------------------------------------------------------------------ sharded evaluation
def _preresnet8_ensemble(U, n_models=3, n=150, seed=5):
    torch.manual_seed(seed)
    ms = [U.models.PreResNet(num_classes=10, depth=8).eval() for _ in range(n_models)]
    g = torch.Generator().manual_seed(seed)
    x, y = torch.randn(n, 3, 32, 32, generator=g), torch.randint(0, 10, (n,), generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)
    return ms, {"in_distribution_test": loader}


def test_prediction_pair_shards_sum_to_the_unsharded_result(U):
    """dist.shard_pairs splits the (sample, image) grid between ranks; summing the ranks' accumulators (what the single
    all-reduce does) must reproduce the unsharded evaluation: same kernels, same per-image sample order."""
    ms, loaders = _preresnet8_ensemble(U)
    full = U.tasks.Prediction(loaders, 10, DEV, ["error_rate", "nll", "brier_score", "ece"])
    full.update_statistics(ms, output_performance=False)
    world = 4
    P = torch.zeros_like(full._proba)
    E = torch.zeros_like(full._entropy)
    for r in range(world):
        part = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
        part.accumulate(ms, U.dist.shard_pairs(len(ms), 150, r, world, quantum=16))
        assert part.last_engine == "fused_preresnet"
        P += part._proba
        E += part._entropy
    assert float((P - full._proba).abs().max()) <= 2e-6
    assert float((E - full._entropy).abs().max()) <= 2e-5
    oi_a, of_a, _, _ = U._C.bma_metrics(P, 3, full._y)
    oi_b, of_b, _, _ = U._C.bma_metrics(full._proba, 3, full._y)
    assert torch.equal(oi_a, oi_b)


def test_host_module_lists_are_packed_in_overlapped_sub_batches(U, monkeypatch):
    """A list of 5+ conv-net modules goes down as sub-batches (2 first, then up to 8, never a one-module tail) so that packing
    overlaps the device; the result must not depend on the split, on whole lists or on a rank's partial image range."""
    from ursabench_b200.tasks import _engine
    ms, loaders = _preresnet8_ensemble(U, n_models=11, n=150)
    sizes = []
    orig = _engine.SampleBank.from_modules
    monkeypatch.setattr(_engine.SampleBank, "from_modules", classmethod(lambda cls, models, device: (sizes.append(len(models)), orig(models, device))[1]))
    split = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
    split.update_statistics(ms, output_performance=False)
    assert sizes == [2, 7, 2] and split.last_engine == "fused_preresnet"
    del sizes[:]
    part = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
    part.accumulate(ms, [(i, 32, 128) for i in range(5)])
    assert sizes == [2, 3]
    monkeypatch.setattr(_engine, "_PACK_SPLIT_MIN", 99)
    del sizes[:]
    whole = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
    whole.update_statistics(ms, output_performance=False)
    assert sizes == [11]
    assert float((split._proba - whole._proba).abs().max()) <= 2e-6
    assert float((split._entropy - whole._entropy).abs().max()) <= 2e-5
    whole5 = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
    whole5.accumulate(ms, [(i, 32, 128) for i in range(5)])
    assert float((part._proba - whole5._proba).abs().max()) <= 2e-6 and float(part._proba[:32].abs().max()) == 0.0


def test_tensor_shaped_loader_fast_path_equals_iteration(U):
    """A sequential DataLoader over a TensorDataset is ingested without re-collating; a shuffled-off custom dataset is
    iterated -- both must give the same resident inputs and targets."""
    ms, loaders = _preresnet8_ensemble(U, n_models=1, n=70)
    a = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
    ds = loaders["in_distribution_test"].dataset

    class Wrapped(torch.utils.data.Dataset):
        def __len__(self):
            return len(ds)

        def __getitem__(self, i):
            return ds[i]
    b = U.tasks.Prediction({"in_distribution_test": torch.utils.data.DataLoader(Wrapped(), batch_size=32)}, 10, DEV, ["error_rate"])
    assert torch.equal(a._x, b._x) and torch.equal(a.targets, b.targets) and a._batch_sizes == b._batch_sizes


def test_non_cifar_inputs_do_not_reach_the_fused_conv_forward(U):
    """The fused conv forwards only take (pointer, N): a loader whose images are not [3, 32, 32] must go through the
    module's own forward and raise the reference's shape error instead of reading out of bounds."""
    torch.manual_seed(0)
    m = U.models.PreResNet(num_classes=10, depth=8).eval()
    x, y = torch.randn(16, 1, 32, 32), torch.randint(0, 10, (16,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 10, DEV, ["error_rate"])
    with pytest.raises(RuntimeError):
        task.update_statistics([m], output_performance=False)
    assert task.last_engine != "fused_preresnet"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_classes_run_on_a_non_current_device(U):
    """``device=cuda:1`` while the current device stays 0 (the reference never calls set_device): the C-ABI launches
    must follow the tensors' device."""
    assert torch.cuda.current_device() == 0
    dev1 = torch.device("cuda", 1)
    ms, loaders = _preresnet8_ensemble(U, n_models=2, n=64)
    t0 = U.tasks.Prediction(loaders, 10, DEV, ["error_rate"])
    t1 = U.tasks.Prediction(loaders, 10, dev1, ["error_rate"])
    t0.update_statistics(ms, output_performance=False)
    t1.update_statistics(ms, output_performance=False)
    assert t1._proba.device == dev1 and torch.cuda.current_device() == 0
    assert torch.allclose(t0.ensemble_proba, t1.ensemble_proba, atol=1e-6)
    ds, loader = _toy()
    torch.manual_seed(0)
    inf = U.inference.SGHMC({"lr": 0.1, "prior_std": 1.0, "num_samples": 2, "alpha": 0.3, "burn_in_epochs": 1},
                            U.models.MLP(16, 20, 3), loader, device=dev1)
    out = inf.sample()
    assert len(out) == 2 and inf.flat.p.device == dev1


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import ursabench_b200 as U
    U.dist.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    try:
        ms, loaders = _preresnet8_ensemble(U, n_models=3, n=150)
        task = U.tasks.Prediction(loaders, 10, dev, ["error_rate", "nll", "brier_score", "ece"], distributed=True,
                                  replicated_samples=True)
        task.update_statistics(ms, output_performance=False)
        got = task.get_performance_metrics()
        proba, _, n_s = task._reduced()
        solo = U.tasks.Prediction(loaders, 10, dev, ["error_rate", "nll", "brier_score", "ece"])
        solo.update_statistics(ms, output_performance=False)
        want = solo.get_performance_metrics()
        assert n_s == 3
        assert float((proba - solo._proba).abs().max()) <= 2e-6
        assert task.get_counters()["correct"] == solo.get_counters()["correct"]
        for k in want:
            assert got[k] == pytest.approx(want[k], abs=1e-6), k
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_gpu_equals_single_gpu_on_nccl(tmp_path):
    """Real NCCL, two ranks, S = 3 samples (S mod world != 0 -> hybrid sample x image sharding): reduced probabilities,
    counters and metrics equal the single-GPU evaluation."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(2))


def test_handles_survive_update_hyp(U, capsys):
    """``sample()`` handles must stay independent of later runs (the reference returns deep copies, sghmc.py:99): after
    ``update_hyp`` + a second ``sample()`` the first run's handles still hold the first run's weights."""
    ds, loader = _toy()
    torch.manual_seed(0)
    hyp = {"lr": 0.1, "prior_std": 1.0, "num_samples": 2, "alpha": 0.3, "burn_in_epochs": 1}
    inf = U.inference.SGHMC(dict(hyp), U.models.MLP(16, 20, 3), loader, device=DEV)
    first = inf.sample()
    rows = [inf.bank.w[h._ursa_row, :inf.flat.D].clone() for h in first]
    old_bank = inf.bank
    inf.update_hyp(dict(hyp, lr=0.05))
    assert inf.bank is not old_bank and getattr(inf, "_graph", None) is None
    second = inf.sample()
    for h, w in zip(first, rows):
        assert h.is_pristine()
        got = torch.cat([p.detach().reshape(-1) for p in h.parameters()])
        assert torch.equal(got, w.cpu())
    assert not torch.equal(_flat(second[0]), _flat(first[0]))


def _swag_shard_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import ursabench_b200 as U
    U.dist.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    try:
        ds, loader = _toy()
        torch.manual_seed(5 + rank)                           # ranks start from different weights: only rank 0's fit counts
        hyp = {"swag_lr": 0.02, "swag_wd": 1e-4, "lr_init": 0.05, "num_samples": 5, "momentum": 0.5, "burn_in_epochs": 2,
               "num_iterates": 6, "subspace_type": "covariance", "shard_draws": True}
        inf = U.inference.SWAG(hyp, U.models.MLP(16, 20, 3), loader, device=dev, max_rank=4)
        out = inf.sample(full_cov=True)
        assert len(out) == len(range(rank, 5, world))         # draws s = rank (mod world)
        mean = inf.weight_mean.clone()
        gathered = [torch.zeros_like(mean) for _ in range(world)]
        dist.all_gather(gathered, mean)
        assert all(torch.equal(gathered[0], g) for g in gathered)          # one fit, broadcast
        assert int(inf.num_models_collected.item()) == 6 and inf.subspace.collected == 6
        mine = torch.stack([inf.bank.w[h._ursa_row, :inf.flat.D] for h in out])
        first = [torch.zeros_like(mine[0]) for _ in range(world)]
        dist.all_gather(first, mine[0].contiguous())
        assert not torch.equal(first[0], first[1])                          # different substreams on different ranks
        assert float((mine - mean[None]).abs().max()) > 0 and bool(torch.isfinite(mine).all())
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_swag_draws_sharded_over_ranks_after_one_broadcast(tmp_path):
    """SURVEY 8e row 2 on real NCCL: rank 0 fits, (mean, second moment, ring) are broadcast once, every rank draws its
    s = rank (mod world) share from its own substream."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_swag_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(2))
