"""Generate tests/golden/* from the LIVE, unmodified reference.

Runs only in the build container (needs /root/reference).  The fixtures it
writes are committed; the GPU box never sees the reference.

    python -m oracle.gen_golden

Protocols:
* identical gradients: ``p.grad`` is assigned directly (optim_sghmc.py:46 reads it);
* identical noise: ``torch.randn_like`` is swapped for a closure that replays
  pre-generated tensors while ``optimSGHMC.step`` runs (optim_sghmc.py:63-64).
"""
import copy
import json
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _ref():
    from oracle import stubs
    if not stubs.reference_available():
        sys.exit("reference not available: goldens can only be regenerated in the build container")
    stubs.install()
    import URSABench  # noqa: F401
    from URSABench import inference, models, tasks, util
    from URSABench.inference import optim_sghmc, subspaces
    return dict(inference=inference, models=models, tasks=tasks, util=util,
                optim_sghmc=optim_sghmc, subspaces=subspaces)


class _ReplayNoise:
    """Context manager: torch.randn_like returns queued tensors."""

    def __init__(self, tensors):
        self.q = list(tensors)

    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda t, **k: self.q.pop(0).view_as(t)
        return self

    def __exit__(self, *a):
        torch.randn_like = self.orig


def gen_sgmcmc_step(R):
    """optimSGHMC.step over ragged tensors, 4 steps, all branches."""
    shapes = [(7,), (13, 5), (64,), (3, 3, 3, 4), (1,), (257,), (31, 33)]
    sizes = [int(np.prod(s)) for s in shapes]
    D = sum(sizes)
    out = {"sizes": np.array(sizes)}
    cases = [
        ("sgld_wd_noise", dict(momentum=0.0, wd=1 / 0.1664 ** 2, lr=0.0999, noise=True)),
        ("sgld_nowd_nonoise", dict(momentum=0.0, wd=0.0, lr=0.01, noise=False)),
        ("sghmc_wd_noise", dict(momentum=1 - 0.10199674218893051, wd=1 / 0.14046818 ** 2, lr=0.03134895861148834, noise=True)),
        ("sghmc_nowd_noise", dict(momentum=0.5, wd=0.0, lr=0.5, noise=True)),
        ("sghmc_wd_nonoise", dict(momentum=0.7, wd=4.0, lr=0.06825362145900726, noise=False)),
    ]
    n_train = 60000
    steps = 4
    for ci, (name, c) in enumerate(cases):
        gen = torch.Generator().manual_seed(100 + ci)
        params = [torch.nn.Parameter(torch.randn(s, generator=gen)) for s in shapes]
        opt = R["optim_sghmc"].optimSGHMC(params, lr=c["lr"], momentum=c["momentum"],
                                          num_training_samples=n_train, weight_decay=c["wd"])
        out[name + "/init"] = torch.cat([p.detach().reshape(-1) for p in params]).numpy().copy()
        out[name + "/hyper"] = np.array([c["lr"], c["momentum"], c["wd"], n_train, float(c["noise"])], np.float64)
        for t in range(steps):
            g = torch.randn(D, generator=gen) * (10.0 if t == 2 else 1.0)
            z = torch.randn(D, generator=gen)
            lr_t = c["lr"] * (1.0 - 0.2 * t)            # the schedule moves lr between steps
            for grp in opt.param_groups:
                grp["lr"] = lr_t
            off = 0
            zs = []
            for p, n in zip(params, sizes):
                p.grad = g[off:off + n].view_as(p).clone()
                zs.append(z[off:off + n].clone())
                off += n
            with _ReplayNoise(zs):
                opt.step(add_langevin_noise=c["noise"])
            out["%s/g%d" % (name, t)] = g.numpy().copy()
            out["%s/z%d" % (name, t)] = z.numpy().copy()
            out["%s/lr%d" % (name, t)] = np.float64(lr_t)
            out["%s/p%d" % (name, t)] = torch.cat([p.detach().reshape(-1) for p in params]).numpy().copy()
            if c["momentum"] != 0:
                out["%s/v%d" % (name, t)] = torch.cat(
                    [opt.state[p]["momentum_buffer"].reshape(-1) for p in params]).numpy().copy()
    np.savez_compressed(os.path.join(OUT, "sgmcmc_step.npz"), **out)
    print("sgmcmc_step.npz", D)


def gen_csghmc_schedule(R):
    """Notebook known-answer trace + live _adjust_learning_rate / gates."""
    nb = open(os.path.join(os.path.dirname(R["inference"].__file__), "..", "examples",
                           "URSABench_MNIST_demo.ipynb")).read()
    trace = [(int(e), float(v)) for e, v in re.findall(r'"Epoch:\s+(\d+)\s+lr:\s+([0-9.e+-]+)\\n"', nb)]
    gold = {"notebook": {"lr_0": 0.06825362145900726, "cycle_length": 22, "num_cycles": 2, "n_train": 60000,
                         "batch_size": 100, "last_batch_idx": 599, "trace": trace}}

    class _DS(torch.utils.data.Dataset):
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return torch.zeros(4), 0

    live = []
    for (n, bs, hyp) in [
        (50000, 128, {"lr_0": 0.5, "prior_std": 0.5, "num_samples_per_cycle": 3, "cycle_length": 50,
                      "burn_in_epochs": 0, "num_cycles": 17, "alpha": 0.5}),
        (1000, 64, {"lr_0": 0.0586, "prior_std": 0.18, "num_samples_per_cycle": 3, "cycle_length": 21,
                    "burn_in_epochs": 1, "num_cycles": 10, "alpha": 1.0}),
        (37, 50, {"lr_0": 0.1, "prior_std": 1.0, "num_samples_per_cycle": 1, "cycle_length": 4,
                  "burn_in_epochs": 1, "num_cycles": 3, "alpha": 0.3}),
    ]:
        loader = torch.utils.data.DataLoader(_DS(n), batch_size=bs)
        obj = R["inference"].cSGHMC(dict(hyp), torch.nn.Linear(4, 2), loader)
        nb_iter = len(loader)
        lrs, noise_gate, sample_gate = [], [], []
        for epoch in range(min(hyp["cycle_length"] * 2 + 3, 60)):
            for b in (0, nb_iter // 2, nb_iter - 1):
                lrs.append([epoch, b, float(obj._adjust_learning_rate(obj.optimizer, epoch, b))])
            noise_gate.append(bool((epoch % obj.cycle_length) + 1 >
                                   (obj.cycle_length - obj.burn_in_epochs - obj.num_samples_per_cycle)))
            sample_gate.append(bool((epoch % obj.cycle_length) >= (obj.cycle_length - obj.num_samples_per_cycle)))
        live.append({"n_train": n, "batch_size": bs, "hyper": hyp, "num_batch": float(obj.num_batch),
                     "total_iterations": float(obj.total_iterations), "lrs": lrs,
                     "noise_gate_by_epochs_run": noise_gate, "sample_gate_by_epoch_index": sample_gate})
    gold["live"] = live
    # SWA schedule (swa.py:92-101)
    hyp = {"swag_lr": 0.01, "swag_wd": 5e-4, "lr_init": 0.05, "num_samples": 2, "momentum": 0.9,
           "burn_in_epochs": 10, "num_iterates": 3}
    loader = torch.utils.data.DataLoader(_DS(10), batch_size=5)
    swa = R["inference"].SWA(hyp, torch.nn.Linear(4, 2), loader)
    gold["swa_schedule"] = {"hyper": hyp, "lr": [float(swa._schedule(e)) for e in range(14)]}
    json.dump(gold, open(os.path.join(OUT, "schedules.json"), "w"), indent=0)
    print("schedules.json", len(trace))


def gen_swa_collect(R):
    """SWA._collect_model (moments + ring) with n incremented by the caller, ring wrap at max_rank=3,
    and the compat trajectory (n pinned at 0: SURVEY Q6)."""

    class _DS(torch.utils.data.Dataset):
        def __len__(self):
            return 8

        def __getitem__(self, i):
            return torch.zeros(6), 0

    loader = torch.utils.data.DataLoader(_DS(), batch_size=4)
    hyp = {"swag_lr": 0.01, "swag_wd": 5e-4, "lr_init": 0.05, "num_samples": 2, "momentum": 0.9,
           "burn_in_epochs": 2, "num_iterates": 3, "subspace_type": "covariance"}
    out = {}
    gen = torch.Generator().manual_seed(7)
    for mode in ("textbook", "swa_biased", "compat_n0"):
        model = torch.nn.Sequential(torch.nn.Linear(6, 37), torch.nn.Linear(37, 5))
        swa = R["inference"].SWA(dict(hyp), model, loader, max_rank=3)
        D = swa.num_parameters
        ws = []
        for k in range(6):
            w = torch.randn(D, generator=gen) * 0.05 + 0.3
            off = 0
            for p in model.parameters():
                p.data.copy_(w[off:off + p.numel()].view_as(p))
                off += p.numel()
            if mode == "swa_biased":
                swa.num_models_collected += 1          # swa.py:130 increments BEFORE collecting (Q8)
            swa._collect_model()
            if mode == "textbook":
                swa.num_models_collected += 1
            ws.append(w.numpy().copy())
            out["%s/mean%d" % (mode, k)] = swa.weight_mean.numpy().copy()
            out["%s/sq%d" % (mode, k)] = swa.sq_mean.numpy().copy()
            out["%s/ring%d" % (mode, k)] = swa.subspace.cov_mat_sqrt.numpy().copy()
            out["%s/rank%d" % (mode, k)] = swa.subspace.rank.numpy().copy()
        out[mode + "/w"] = np.stack(ws)
        mean, var = swa._get_mean_and_variance()
        out[mode + "/var"] = var.numpy().copy()
        out[mode + "/space"] = swa.subspace.get_space().numpy().copy()
    # diag draw formula (swag.py:86): torch.normal(mean, std) == mean + std * z for the same generator stream
    g1 = torch.Generator().manual_seed(3)
    mean = torch.randn(1000, generator=g1)
    std = torch.rand(1000, generator=g1)
    g2 = torch.Generator().manual_seed(11)
    draw = torch.normal(mean, std, generator=g2)
    g2 = torch.Generator().manual_seed(11)
    z = torch.randn(1000, generator=g2)
    out["normal/mean"], out["normal/std"], out["normal/z"], out["normal/draw"] = \
        mean.numpy(), std.numpy(), z.numpy(), draw.numpy()
    np.savez_compressed(os.path.join(OUT, "swa_collect.npz"), **out)
    print("swa_collect.npz")


def gen_swag_compat(R):
    """End-to-end reference SWAG.sample(): pins quirks Q5/Q6 (facts, not weights)."""
    torch.manual_seed(0)
    x = torch.randn(64, 1, 4, 5)
    y = torch.randint(0, 3, (64,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=16, shuffle=False)
    model = R["models"].mlp.MLP(8, 20, 3)
    hyp = {"swag_lr": 0.01, "swag_wd": 0.0, "lr_init": 0.05, "num_samples": 3, "momentum": 0.5,
           "burn_in_epochs": 2, "num_iterates": 3, "subspace_type": "covariance"}
    swag = R["inference"].SWAG(hyp, model, loader)
    samples = swag.sample()
    last = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    flat = [torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in samples]
    crash = None
    try:
        swag.sample_iterative(full_cov=True)
    except AttributeError as e:
        crash = str(e)
    facts = {
        "num_models_collected": int(swag.num_models_collected.item()),
        "all_samples_identical": bool(all(torch.equal(flat[0], f) for f in flat)),
        "sample_equals_last_iterate": bool(torch.equal(flat[0], last)),
        "variance_all_clamp": bool((swag.weight_variance == 1e-30).all().item()),
        "ring_all_zero": bool((swag.subspace.cov_mat_sqrt == 0).all().item()),
        "ring_rows": int(swag.subspace.cov_mat_sqrt.shape[0]),
        "full_cov_raises": crash,
        "num_returned": len(samples),
        "samples_on_cpu": all(next(m.parameters()).device.type == "cpu" for m in samples),
    }
    json.dump(facts, open(os.path.join(OUT, "swag_compat_facts.json"), "w"), indent=1)
    print("swag_compat_facts.json", facts)


def _flat_params(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy().copy()


def _flat_buffers(model):
    bufs = [b.detach().reshape(-1).float() for n, b in model.named_buffers() if b.dtype.is_floating_point]
    return torch.cat(bufs).numpy().copy() if bufs else np.zeros(0, np.float32)


def gen_prediction(R):
    """Prediction.update_statistics + get_performance_metrics on fixed weights."""
    out = {}
    metric_json = {}
    Prediction = R["tasks"].Prediction
    # -- MLP, C=10, ragged last batch, two update calls (accumulation across calls)
    torch.manual_seed(1)
    S, N, C = 5, 300, 10
    x = torch.randn(N, 1, 6, 8)
    y = torch.randint(0, C, (N,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=128, shuffle=False)
    ms = []
    for s in range(S):
        m = R["models"].mlp.MLP(24, 48, C)
        for p in m.parameters():
            p.data.mul_(3.0)                      # sharper logits
        ms.append(m)
    task = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), "ALL")
    task.update_statistics(ms[:3], output_performance=False)
    task.update_statistics(ms[3:], output_performance=False)
    metric_json["mlp"] = {k: float(v) for k, v in task.get_performance_metrics().items()}
    out["mlp/bank"] = np.stack([_flat_params(m) for m in ms])
    out["mlp/x"], out["mlp/y"] = x.numpy(), y.numpy()
    out["mlp/ensemble_proba"] = task.ensemble_proba.numpy().copy()
    out["mlp/entropy"] = task.expected_data_uncertainty.numpy().copy()
    out["mlp/logits"] = torch.stack([m(x) for m in ms]).detach().numpy()
    out["mlp/arch"] = np.array([24, 48, C])
    # single-module call + output_performance float
    t1 = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), ["ll"])
    metric_json["mlp_single_ll"] = float(t1.update_statistics(ms[0], output_performance=True))
    # -- MLP, C=100 (config 3's class count)
    torch.manual_seed(2)
    S, N, C = 3, 130, 100
    x = torch.randn(N, 20)
    y = torch.randint(0, C, (N,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=64, shuffle=False)
    ms = [R["models"].mlp.MLP(16, 20, C) for _ in range(S)]
    for m in ms:
        for p in m.parameters():
            p.data.mul_(4.0)
    task = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), "ALL")
    task.update_statistics(ms, output_performance=False)
    metric_json["mlp_c100"] = {k: float(v) for k, v in task.get_performance_metrics().items()}
    out["mlp_c100/bank"] = np.stack([_flat_params(m) for m in ms])
    out["mlp_c100/x"], out["mlp_c100/y"] = x.numpy(), y.numpy()
    out["mlp_c100/ensemble_proba"] = task.ensemble_proba.numpy().copy()
    out["mlp_c100/entropy"] = task.expected_data_uncertainty.numpy().copy()
    out["mlp_c100/arch"] = np.array([16, 20, C])
    # -- PreResNet depth 8 (all block kinds: identity, stride-2 + 1x1 downsample), eval-mode BN with
    #    non-trivial running stats and affine
    torch.manual_seed(3)
    S, N, C = 2, 20, 10
    x = torch.randn(N, 3, 32, 32)
    y = torch.randint(0, C, (N,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8, shuffle=False)
    ms = []
    for s in range(S):
        m = R["models"].preresnet.PreResNet(num_classes=C, depth=8)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data.uniform_(0.5, 1.5)
                mod.bias.data.normal_(0, 0.2)
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 2.0)
        m.fc.weight.data.mul_(6.0)
        ms.append(m)
    task = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), "ALL")
    task.update_statistics(ms, output_performance=False)
    metric_json["preresnet8"] = {k: float(v) for k, v in task.get_performance_metrics().items()}
    out["preresnet8/bank"] = np.stack([_flat_params(m) for m in ms])
    out["preresnet8/buffers"] = np.stack([_flat_buffers(m) for m in ms])
    out["preresnet8/x"], out["preresnet8/y"] = x.numpy().astype(np.float16).astype(np.float32), y.numpy()
    # x is stored at fp16 resolution to keep the fixture small: recompute the reference on the rounded x
    xr = torch.from_numpy(out["preresnet8/x"])
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xr, y), batch_size=8, shuffle=False)
    task = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), "ALL")
    task.update_statistics(ms, output_performance=False)
    metric_json["preresnet8"] = {k: float(v) for k, v in task.get_performance_metrics().items()}
    out["preresnet8/x"] = out["preresnet8/x"].astype(np.float16)
    out["preresnet8/ensemble_proba"] = task.ensemble_proba.numpy().copy()
    out["preresnet8/entropy"] = task.expected_data_uncertainty.numpy().copy()
    with torch.no_grad():
        out["preresnet8/logits"] = torch.stack([m.eval()(xr) for m in ms]).numpy()
    np.savez_compressed(os.path.join(OUT, "prediction.npz"), **out)
    json.dump(metric_json, open(os.path.join(OUT, "prediction_metrics.json"), "w"), indent=1)
    print("prediction.npz")


def gen_prediction_wrn(R):
    """Prediction.update_statistics on the reference WideResNet (WRN-10-2: every block is a transition block with a folded
    1x1 shortcut; WRN-16-2: identity blocks too).  Weights come from oracle/wrn_fill.py (seeded), so only x and the
    reference's outputs are stored."""
    import warnings
    from oracle.wrn_fill import wrn_fill
    Prediction = R["tasks"].Prediction
    out, metric_json = {}, {}
    for tag, depth, widen, C, S, N, seed, gain in (("wrn10x2", 10, 2, 10, 2, 11, 100, 1.0), ("wrn16x2", 16, 2, 100, 2, 5, 200, 0.25)):
        rng = np.random.RandomState(seed + 50)
        x16 = rng.randn(N, 3, 32, 32).astype(np.float16)
        x = torch.from_numpy(x16.astype(np.float32))
        y = torch.from_numpy(rng.randint(0, C, N))
        ms = []
        for s in range(S):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m = R["models"].wideresnet.WideResNet(num_classes=C, depth=depth, widen_factor=widen)
            ms.append(wrn_fill(m, seed + s, logit_gain=gain))
        loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=4, shuffle=False)
        task = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), "ALL")
        task.update_statistics(ms, output_performance=False)
        metric_json[tag] = {k: float(v) for k, v in task.get_performance_metrics().items()}
        out[tag + "/x"], out[tag + "/y"] = x16, y.numpy()
        out[tag + "/arch"] = np.array([depth, widen, C, S, seed])
        out[tag + "/gain"] = np.array([gain])
        out[tag + "/ensemble_proba"] = task.ensemble_proba.numpy().copy()
        out[tag + "/entropy"] = task.expected_data_uncertainty.numpy().copy()
        with torch.no_grad():
            out[tag + "/logits"] = torch.stack([m.eval()(x) for m in ms]).numpy()
        out[tag + "/D"] = np.array([sum(p.numel() for p in ms[0].parameters())])
    np.savez_compressed(os.path.join(OUT, "prediction_wrn.npz"), **out)
    json.dump(metric_json, open(os.path.join(OUT, "prediction_wrn_metrics.json"), "w"), indent=1)
    print("prediction_wrn.npz", {k: v for k, v in metric_json.items()})


def gen_prediction_preresnet20(R):
    """Prediction.update_statistics on the reference's PreResNet(depth=20) -- the north-star model -- with logits sharpened to
    the scale of a trained network (|logit| ~ 20), C = 10 and C = 100.  Weights come from oracle/wrn_fill.py (a seeded
    stream in flat-layout order), so only x and the reference's outputs are stored."""
    from oracle.wrn_fill import wrn_fill
    Prediction = R["tasks"].Prediction
    out, metric_json = {}, {}
    for tag, C, S, N, seed, gain in (("c10", 10, 2, 24, 300, 0.4), ("c100", 100, 2, 16, 400, 0.5)):
        rng = np.random.RandomState(seed + 50)
        x16 = rng.randn(N, 3, 32, 32).astype(np.float16)
        x = torch.from_numpy(x16.astype(np.float32))
        y = torch.from_numpy(rng.randint(0, C, N))
        ms = [wrn_fill(R["models"].preresnet.PreResNet(num_classes=C, depth=20), seed + s, logit_gain=gain) for s in range(S)]
        loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8, shuffle=False)
        task = Prediction({"in_distribution_test": loader}, C, torch.device("cpu"), "ALL")
        task.update_statistics(ms, output_performance=False)
        metric_json[tag] = {k: float(v) for k, v in task.get_performance_metrics().items()}
        out[tag + "/x"], out[tag + "/y"] = x16, y.numpy()
        out[tag + "/arch"] = np.array([20, C, S, seed])
        out[tag + "/gain"] = np.array([gain])
        out[tag + "/ensemble_proba"] = task.ensemble_proba.numpy().copy()
        out[tag + "/entropy"] = task.expected_data_uncertainty.numpy().copy()
        with torch.no_grad():
            logits = torch.stack([m.eval()(x) for m in ms])
            out[tag + "/logits"] = logits.numpy()
            l64 = torch.stack([copy.deepcopy(m).double().eval()(x.double()) for m in ms])
            # how far the reference's own fp32 forward is from the exact network on these inputs (context for the 1e-5 bar)
            out[tag + "/fp32_vs_fp64_proba"] = np.array([(torch.softmax(logits.double(), -1) - torch.softmax(l64, -1)).abs().max().item()])
        print(tag, "max |logit|", float(logits.abs().max()), "fp32 vs fp64 proba", float(out[tag + "/fp32_vs_fp64_proba"][0]))
    np.savez_compressed(os.path.join(OUT, "prediction_preresnet20.npz"), **out)
    json.dump(metric_json, open(os.path.join(OUT, "prediction_preresnet20_metrics.json"), "w"), indent=1)
    print("prediction_preresnet20.npz")


def gen_ess(R):
    """util.elliptical_slice (util.py:287-354) and util.log_pdf (:260-274) of the live reference: (a) 25 chained ESS updates
    of a 5-d correlated-Gaussian log density with seeded ``np.random`` (thetas, log densities, number of density calls),
    (b) ``log_pdf`` of three subspace points for an MLP(16, 20, 3) over a 3-batch loader at temperature 7."""
    util = R["util"]
    from URSABench.inference.projection_model import SubspaceModel
    out = {}
    rng = np.random.RandomState(5)
    A = rng.randn(5, 5)
    prec = A @ A.T + 0.5 * np.eye(5)
    mu = rng.randn(5)
    calls = [0]

    def lnpdf(th, subspace):
        calls[0] += 1
        d = th - mu
        return float(-0.5 * d @ prec @ d)

    np.random.seed(123)
    theta = np.zeros(5)
    thetas, lps, ncalls = [], [], []
    for _ in range(25):
        prior = np.random.normal(loc=0.0, scale=2.0, size=5)
        theta, lp = util.elliptical_slice(initial_theta=theta.copy(), prior=prior, lnpdf=lnpdf, subspace=None)
        thetas.append(theta.copy()); lps.append(lp); ncalls.append(calls[0])
    out["ess/prec"], out["ess/mu"] = prec, mu
    out["ess/thetas"], out["ess/lps"], out["ess/ncalls"] = np.stack(thetas), np.array(lps), np.array(ncalls)
    # (b) log_pdf
    torch.manual_seed(9)
    model = R["models"].mlp.MLP(16, 20, 3)
    D = sum(p.numel() for p in model.parameters())
    mean = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    factor = torch.randn(4, D) * 0.05
    x = torch.randn(70, 1, 4, 5)
    y = torch.randint(0, 3, (70,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)
    sub = SubspaceModel(mean, factor)
    ts = np.array([[0.0, 0.0, 0.0, 0.0], [0.5, -1.0, 0.25, 2.0], [-3.0, 1.5, 0.0, 0.7]])
    out["logpdf/mean"], out["logpdf/factor"] = mean.numpy(), factor.numpy()
    out["logpdf/x"], out["logpdf/y"], out["logpdf/t"] = x.numpy(), y.numpy(), ts
    out["logpdf/value"] = np.array([util.log_pdf(t, sub, model, loader, util.cross_entropy, 7.0, torch.device("cpu")) for t in ts])
    np.savez_compressed(os.path.join(OUT, "ess.npz"), **out)
    print("ess.npz", out["ess/ncalls"][-1], out["logpdf/value"])


def gen_pca_space(R):
    """PCASpace.collect_vector / get_space (inference/subspaces.py:103-156, integer pca_rank) and SubspaceModel.forward
    (inference/projection_model.py:6-14) on the live reference: ring wrap (11 collects into max_rank 8), pca_rank 5; a
    second case with fewer collects than max_rank (rank 3 < pca_rank)."""
    from URSABench.inference.projection_model import SubspaceModel
    out = {}
    for tag, D, max_rank, pca_rank, ncollect, seed in (("wrap", 503, 8, 5, 11, 0), ("short", 130, 20, 6, 3, 1)):
        rng = np.random.RandomState(seed)
        sp = R["subspaces"].PCASpace(num_parameters=D, pca_rank=pca_rank, max_rank=max_rank)
        scales = np.linspace(2.0, 0.2, ncollect)
        vecs = (rng.randn(ncollect, D) * scales[:, None]).astype(np.float32)
        for v in vecs:
            sp.collect_vector(torch.from_numpy(v))
        np.random.seed(seed)                              # randomized_svd draws its test matrix from the global numpy RNG
        space = sp.get_space()
        out[tag + "/vecs"] = vecs
        out[tag + "/cfg"] = np.array([D, max_rank, pca_rank, ncollect])
        out[tag + "/ring"] = sp.cov_mat_sqrt.numpy().copy()
        out[tag + "/rank"] = sp.rank.numpy().copy()
        out[tag + "/space"] = space.numpy().copy()
        mean = torch.from_numpy(rng.randn(D).astype(np.float32))
        t = torch.from_numpy(rng.randn(space.shape[0]).astype(np.float32))
        out[tag + "/mean"], out[tag + "/t"] = mean.numpy(), t.numpy()
        out[tag + "/projected"] = SubspaceModel(mean, space)(t).numpy().copy()
    # pca_rank = 'mle' (subspaces.py:123-153): the live reference's branch around the restated scikit-learn <= 0.22
    # ``_assess_dimension_`` (oracle/stubs.py binds oracle/restate.py::assess_dimension_sklearn022 in its place).  Three
    # spectra: a clear 3-direction signal in noise, a full ring with a 5-direction signal, pure noise.
    for tag, D, max_rank, ncollect, nsig, seed in (("mle3", 2000, 10, 10, 3, 2), ("mle5", 1500, 12, 17, 5, 3), ("mle0", 900, 8, 8, 0, 4)):
        rng = np.random.RandomState(seed)
        sp = R["subspaces"].PCASpace(num_parameters=D, pca_rank="mle", max_rank=max_rank)
        basis = rng.randn(max(nsig, 1), D).astype(np.float32)
        coef = rng.randn(ncollect, max(nsig, 1)).astype(np.float32) * (np.linspace(6.0, 3.0, max(nsig, 1))[None, :] if nsig else 0.0)
        vecs = (coef @ basis + 0.3 * rng.randn(ncollect, D)).astype(np.float32)
        for v in vecs:
            sp.collect_vector(torch.from_numpy(v))
        np.random.seed(seed)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints the chosen rank
            space = sp.get_space()
        out[tag + "/vecs"] = vecs
        out[tag + "/cfg"] = np.array([D, max_rank, ncollect])
        out[tag + "/ring"] = sp.cov_mat_sqrt.numpy().copy()
        out[tag + "/rank"] = sp.rank.numpy().copy()
        out[tag + "/space"] = space.numpy().copy()
        out[tag + "/ll"] = np.asarray(sp.ll, dtype=np.float64)
        out[tag + "/corrected_ll"] = np.asarray(sp.corrected_ll, dtype=np.float64)
        out[tag + "/chosen"] = np.array([int(sp.pca_rank)])
    np.savez_compressed(os.path.join(OUT, "pca_space.npz"), **out)
    print("pca_space.npz", {k: out[k].shape for k in out if k.endswith("space")}, {k: out[k] for k in out if k.endswith("chosen")})


def gen_bn_update(R):
    """util.bn_update (util.py:212-247) of the LIVE reference on its own WideResNet (WRN-10-2), CPU: the reference hard-codes
    ``input.cuda(non_blocking=True)`` (util.py:236), patched out harness-side for the call.  Weights from oracle/wrn_fill.py
    (seeded), ragged last batch (N = 44, batch 16)."""
    import warnings
    from oracle.wrn_fill import wrn_fill
    depth, widen, C, N, batch, seed = 10, 2, 10, 44, 16, 300
    rng = np.random.RandomState(seed + 1)
    x16 = rng.randn(N, 3, 32, 32).astype(np.float16)
    x = torch.from_numpy(x16.astype(np.float32))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.zeros(N, dtype=torch.long)), batch_size=batch,
                                         shuffle=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = wrn_fill(R["models"].wideresnet.WideResNet(num_classes=C, depth=depth, widen_factor=widen), seed)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # harness-side: no GPU in this container
    try:
        R["util"].bn_update(loader, m)
    finally:
        torch.Tensor.cuda = orig
    out = {"x": x16, "cfg": np.array([depth, widen, C, N, batch, seed]), "buffers": _flat_buffers(m),
           "momentum_after": np.array([mod.momentum for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)], np.float64)}
    np.savez_compressed(os.path.join(OUT, "bn_update.npz"), **out)
    print("bn_update.npz", out["buffers"].shape, out["momentum_after"][:3])


def gen_bn_update_preresnet(R):
    """util.bn_update (util.py:212-247) of the LIVE reference on its own PreResNet(depth=8) and PreResNet(depth=20), CPU (the
    hard-coded ``input.cuda()`` of util.py:236 patched out harness-side).  Weights from oracle/wrn_fill.py (seeded), ragged last
    batch."""
    from oracle.wrn_fill import wrn_fill
    out = {}
    for tag, depth, C, N, batch, seed in (("d8", 8, 10, 44, 16, 500), ("d20", 20, 10, 40, 16, 600)):
        rng = np.random.RandomState(seed + 1)
        x16 = rng.randn(N, 3, 32, 32).astype(np.float16)
        x = torch.from_numpy(x16.astype(np.float32))
        loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.zeros(N, dtype=torch.long)), batch_size=batch,
                                             shuffle=False)
        m = wrn_fill(R["models"].preresnet.PreResNet(num_classes=C, depth=depth), seed, logit_gain=0.4)
        orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self          # harness-side: no GPU in this container
        try:
            R["util"].bn_update(loader, m)
        finally:
            torch.Tensor.cuda = orig
        out[tag + "/x"] = x16
        out[tag + "/cfg"] = np.array([depth, C, N, batch, seed])
        out[tag + "/buffers"] = _flat_buffers(m)
    np.savez_compressed(os.path.join(OUT, "bn_update_preresnet.npz"), **out)
    print("bn_update_preresnet.npz", {k: v.shape for k, v in out.items() if k.endswith("buffers")})


def gen_metrics_edge(R):
    """get_performance_metrics / _get_ece / _get_brier on crafted probabilities: confidences exactly on bin
    edges, argmax ties, one-hot rows, C = 2..100."""
    from URSABench.tasks import prediction as P
    rng = np.random.RandomState(5)
    out = {}
    res = {}
    cases = {}
    # random dirichlet-like sums of S samples
    for name, (N, C, S, conc) in {"rand_c10": (1000, 10, 7, 0.3), "rand_c100": (500, 100, 30, 0.05),
                                  "rand_c2": (257, 2, 3, 1.0)}.items():
        p = rng.gamma(conc, size=(S, N, C)).astype(np.float32) + 1e-12
        p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
        cases[name] = (p.sum(0).astype(np.float32), S, rng.randint(0, C, N))
    # edge rows: S=1 so that pbar == stored value
    edges = np.arange(16, dtype=np.float64) / 15
    rows, tg = [], []
    for k in range(1, 16):
        for v in (np.float32(edges[k]), np.nextafter(np.float32(edges[k]), np.float32(2)),
                  np.nextafter(np.float32(edges[k]), np.float32(0))):
            if v > 1 or v < 0.5:
                continue
            r = np.zeros(4, np.float32)
            r[1] = v
            r[3] = np.float32(1) - v
            rows.append(r)
            tg.append(1 if len(rows) % 2 else 3)
    rows.append(np.array([0.25, 0.25, 0.25, 0.25], np.float32)); tg.append(0)   # 4-way tie -> index 0
    rows.append(np.array([0.1, 0.4, 0.4, 0.1], np.float32)); tg.append(2)       # tie -> first max (1) -> wrong
    rows.append(np.array([0.0, 0.0, 1.0, 0.0], np.float32)); tg.append(2)       # one-hot
    rows.append(np.array([0.0, 0.0, 0.0, 0.0], np.float32)); tg.append(0)       # conf 0 falls in no bin
    cases["edges_c4"] = (np.stack(rows), 1, np.array(tg))
    for name, (psum, S, y) in cases.items():
        N, C = psum.shape
        x = torch.zeros(N, 1)
        yt = torch.from_numpy(y.astype(np.int64))
        loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, yt), batch_size=64, shuffle=False)
        task = P.Prediction({"in_distribution_test": loader}, C, torch.device("cpu"),
                            ["error_rate", "nll", "ll", "brier_score", "ece"])
        task.ensemble_proba = torch.from_numpy(psum.copy())
        task.num_samples_collected = S
        m = task.get_performance_metrics()
        res[name] = {k: float(v) for k, v in m.items()}
        res[name]["nll_nosmooth"] = float(task.get_performance_metrics(smoothing=False)["nll"]) \
            if name != "edges_c4" else None
        pbar = (task.ensemble_proba / S).numpy()
        res[name]["ece_direct"] = float(P._get_ece(pbar, y))
        res[name]["brier_direct"] = float(P._get_brier(pbar, y))
        out[name + "/proba_sum"] = psum
        out[name + "/y"] = y.astype(np.int64)
        out[name + "/S"] = np.int64(S)
    # smoothing / entropy helpers
    pr = torch.from_numpy(cases["rand_c10"][0] / 7)
    out["smooth/in"] = pr.numpy()
    out["smooth/out"] = R["util"].central_smoothing(pr).numpy()
    out["smooth/entropy"] = R["util"].compute_predictive_entropy(R["util"].central_smoothing(pr)).numpy()
    np.savez_compressed(os.path.join(OUT, "metrics_edge.npz"), **out)
    json.dump(res, open(os.path.join(OUT, "metrics_edge.json"), "w"), indent=1)
    print("metrics_edge.npz")


def gen_layouts(R):
    """Flat layouts (name, shape) in model.parameters() order + buffer order."""
    lay = {}
    mk = {
        "MLP400_c10": lambda: R["models"].mlp.MLP(400, 784, 10),
        "PreResNet20_c10": lambda: R["models"].preresnet.PreResNet(num_classes=10, depth=20),
        "PreResNet8_c10": lambda: R["models"].preresnet.PreResNet(num_classes=10, depth=8),
        "WRN28x10_c100": lambda: R["models"].wideresnet.WideResNet(num_classes=100, depth=28, widen_factor=10),
    }
    import warnings
    for k, f in mk.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = f()
        lay[k] = {"params": [[n, list(p.shape)] for n, p in m.named_parameters()],
                  "buffers": [[n, list(b.shape), str(b.dtype)] for n, b in m.named_buffers()],
                  "D": int(sum(p.numel() for p in m.parameters()))}
    json.dump(lay, open(os.path.join(OUT, "layouts.json"), "w"))
    print("layouts.json", {k: v["D"] for k, v in lay.items()})


def gen_ood_decision(R):
    """OODDetection (tasks/ood_detection.py:39-130) and Decision (tasks/decision_making.py:83-152) on fixed weights."""
    import torchvision
    out, js = {}, {}
    # -- OOD: MLP 48-24-10, in-distribution N(0,1), out-of-distribution 3 N(0,1) + 1, two update calls
    torch.manual_seed(21)
    S, C = 4, 10
    xin, xout = torch.randn(300, 1, 6, 8), torch.randn(170, 1, 6, 8) * 3.0 + 1.0
    lin = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xin, torch.zeros(300, dtype=torch.long)), batch_size=128)
    lout = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xout, torch.zeros(170, dtype=torch.long)), batch_size=64)
    ms = []
    for s in range(S):
        m = R["models"].mlp.MLP(24, 48, C)
        for p in m.parameters():
            p.data.mul_(3.0)
        ms.append(m)
    task = R["tasks"].OODDetection({"in_distribution_test": lin, "out_distribution_test": lout}, C, torch.device("cpu"))
    task.update_statistics(ms[:1], output_performance=False)
    task.update_statistics(ms[1:], output_performance=False)
    js["ood"] = {k: float(v) for k, v in task.get_performance_metrics().items()}
    out["ood/bank"] = np.stack([_flat_params(m) for m in ms])
    out["ood/arch"] = np.array([24, 48, C])
    out["ood/x_in"], out["ood/x_out"] = xin.numpy(), xout.numpy()
    for k in ("in_distribution_ensemble_proba", "out_distribution_ensemble_proba", "in_distribution_data_uncertainty",
              "out_distribution_data_uncertainty", "in_distribution_total_uncertainty", "out_distribution_total_uncertainty",
              "in_distribution_model_uncertainty", "out_distribution_model_uncertainty"):
        out["ood/" + k] = getattr(task, k).numpy().copy()
    js["ood_single"] = {k: float(v) for k, v in R["tasks"].OODDetection(
        {"in_distribution_test": lin, "out_distribution_test": lout}, C, torch.device("cpu")).update_statistics(ms[0]).items()}
    # -- Decision: the cost matrix is chosen by the dataset CLASS (decision_making.py:95-102), so build a torchvision MNIST
    #    object around synthetic uint8 images without touching the disk
    torch.manual_seed(22)
    N = 90
    ds = torchvision.datasets.MNIST.__new__(torchvision.datasets.MNIST)
    ds.data = torch.randint(0, 256, (N, 28, 28), dtype=torch.uint8)
    ds.targets = torch.randint(0, 10, (N,))
    ds.transform = torchvision.transforms.ToTensor()
    ds.target_transform = None
    loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False)
    ms = []
    for s in range(3):
        m = R["models"].mlp.MLP(16, 784, 10)
        for p in m.parameters():
            p.data.mul_(2.0)
        ms.append(m)
    task = R["tasks"].Decision({"decision_data_test": loader}, 10, torch.device("cpu"))
    task.update_statistics(ms[:2], output_performance=False)
    res = task.update_statistics(ms[2:], output_performance=True)
    out["decision/bank"] = np.stack([_flat_params(m) for m in ms])
    out["decision/arch"] = np.array([16, 784, 10])
    out["decision/data_u8"], out["decision/targets"] = ds.data.numpy(), ds.targets.numpy()
    out["decision/cost_mat"] = task.cost_mat.numpy().copy()
    out["decision/ensemble_proba"] = task.ensemble_proba.numpy().copy()
    out["decision/risk"] = task.risk.numpy().copy()
    out["decision/D"] = res["Decision"].numpy().copy()
    js["decision"] = {"True_Cost": float(res["True_Cost"]), "num_samples_collected": int(task.num_samples_collected)}
    for name, fn in (("MNIST", "MNIST_cost"), ("CIFAR10", "CIFAR10_cost"), ("CIFAR100", "CIFAR100_cost")):
        import URSABench.tasks.decision_making as dm
        out["cost/" + name] = getattr(dm, fn)(100 if name == "CIFAR100" else 10).numpy()
    np.savez_compressed(os.path.join(OUT, "ood_decision.npz"), **out)
    json.dump(js, open(os.path.join(OUT, "ood_decision.json"), "w"), indent=1)
    print("ood_decision.npz")


def main():
    os.makedirs(OUT, exist_ok=True)
    R = _ref()
    gens = dict(sgmcmc_step=gen_sgmcmc_step, csghmc_schedule=gen_csghmc_schedule, swa_collect=gen_swa_collect,
                swag_compat=gen_swag_compat, prediction=gen_prediction, prediction_wrn=gen_prediction_wrn, prediction_preresnet20=gen_prediction_preresnet20, ess=gen_ess, pca_space=gen_pca_space, bn_update=gen_bn_update, bn_update_preresnet=gen_bn_update_preresnet, metrics_edge=gen_metrics_edge, layouts=gen_layouts,
                ood_decision=gen_ood_decision)
    for name in (sys.argv[1:] or list(gens)):        # `python -m oracle.gen_golden ood_decision` regenerates one fixture
        gens[name](R)


if __name__ == "__main__":
    main()
