"""PCA-subspace elliptical slice sampling (SURVEY 8(f).3 remainder) against goldens of the LIVE reference
(oracle/gen_golden.py::gen_ess: ``util.elliptical_slice``, ``util.log_pdf``)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ess.npz")


def test_elliptical_slice_reproduces_the_reference_chain():
    """Same ``np.random`` state, same log density -> the same 25 states, log densities and number of density calls
    (bracket shrinkage and RNG call order are the reference's, util.py:287-354).  No GPU involved."""
    from ursabench_b200.inference.pca_subspace import elliptical_slice
    g = np.load(GOLD)
    prec, mu = g["ess/prec"], g["ess/mu"]
    calls = [0]

    def lnpdf(th, subspace):
        calls[0] += 1
        d = th - mu
        return float(-0.5 * d @ prec @ d)

    np.random.seed(123)
    theta = np.zeros(5)
    for i in range(25):
        prior = np.random.normal(loc=0.0, scale=2.0, size=5)
        theta, lp = elliptical_slice(initial_theta=theta.copy(), prior=prior, lnpdf=lnpdf, subspace=None)
        assert np.array_equal(theta, g["ess/thetas"][i])
        assert lp == g["ess/lps"][i] and calls[0] == g["ess/ncalls"][i]
    with pytest.raises(IOError):
        elliptical_slice(np.zeros(3), np.zeros((2, 2)), lnpdf, cur_lnpdf=0.0)


def _toy_loader():
    g = np.load(GOLD)
    x, y = torch.from_numpy(g["logpdf/x"]), torch.from_numpy(g["logpdf/y"])
    return g, torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)


@pytest.mark.gpu
def test_log_pdf_of_subspace_points_matches_reference():
    import ursabench_b200 as U
    g, loader = _toy_loader()
    dev = torch.device("cuda")
    hyp = dict(U.inference.PCASubspaceSampler._defaults, rank=4, max_rank=4, temperature=7.0)
    inf = U.inference.PCASubspaceSampler(hyp, U.models.MLP(16, 20, 3), loader, device=dev)
    inf.subspace = U.inference.SubspaceModel(torch.from_numpy(g["logpdf/mean"]).to(dev), torch.from_numpy(g["logpdf/factor"]).to(dev))
    got = np.array([inf._oracle(t) for t in g["logpdf/t"]])
    np.testing.assert_allclose(got, g["logpdf/value"], rtol=2e-5, atol=2e-5)
    # the free function under the reference's name and signature (util.py:260-274) gives the same numbers
    model = U.models.MLP(16, 20, 3).to(dev)
    free = np.array([U.util.log_pdf(t, inf.subspace, model, loader, U.util.cross_entropy, 7.0, dev) for t in g["logpdf/t"]])
    np.testing.assert_allclose(free, g["logpdf/value"], rtol=2e-5, atol=2e-5)
    # the projection really went through the flat buffer the model's parameters are views of
    w = torch.from_numpy(g["logpdf/mean"] + g["logpdf/factor"].T @ g["logpdf/t"][-1].astype(np.float32)).float()
    now = torch.cat([p.detach().reshape(-1) for p in inf.model.parameters()]).cpu()
    assert torch.allclose(now, w, atol=1e-5)


@pytest.mark.gpu
def test_pca_subspace_sampler_end_to_end(capsys):
    """Class API as the reference's drivers use it: construct, ``sample()``, ``update_hyp``, ``sample()`` again; the chain
    moves, the samples are bank handles that ``Prediction`` evaluates in place."""
    import ursabench_b200 as U
    g, loader = _toy_loader()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    np.random.seed(0)
    hyp = {"swag_lr": 0.02, "swag_wd": 1e-4, "lr_init": 0.05, "num_samples": 4, "swag_momentum": 0.5, "swag_burn_in_epochs": 2,
           "num_swag_iterates": 5, "rank": 3, "max_rank": 5, "temperature": 50.0, "prior_std": 1.0}
    inf = U.inference.PCASubspaceSampler(hyp, U.models.MLP(16, 20, 3), loader, device=dev)
    out = inf.sample()
    assert len(out) == 4 and inf.subspace.rank == 3 and inf.lnpdf_evaluations >= 8
    flats = [torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in out]
    assert not torch.equal(flats[0], flats[-1])
    # every sample lies in mean + span(cov_factor)
    dev_vec = (flats[-1].to(dev) - inf.weight_mean)
    coef = torch.linalg.lstsq(inf.weight_covariance.t().double(), dev_vec.double()[:, None]).solution[:, 0]
    assert float((inf.weight_covariance.t().double() @ coef - dev_vec.double()).abs().max()) < 1e-4
    assert torch.allclose(coef.float().cpu(), inf.current_theta, atol=1e-3)
    task = U.tasks.Prediction({"in_distribution_test": loader}, 3, dev, ["error_rate", "nll"])
    task.update_statistics(out, output_performance=False)
    assert task.last_engine == "fused_mlp" and 0.0 <= task.get_performance_metrics()["error_rate"] <= 1.0
    inf.update_hyp(dict(hyp, num_samples=2))
    assert inf.subspace is None and inf.current_theta is None
    assert len(inf.sample()) == 2


def test_util_exports_the_reference_helpers_of_the_path():
    """``from ursabench_b200 import util`` must offer what the reference's samplers import from its util on this path
    (util.py:110-123, 185-199, 260-354)."""
    import ursabench_b200 as U
    from ursabench_b200.inference import pca_subspace
    for name in ("flatten", "set_weights", "unflatten_like", "check_bn", "reset_bn", "bn_update", "central_smoothing",
                 "compute_predictive_entropy", "cross_entropy", "log_pdf", "elliptical_slice", "convert_sample_to_net",
                 "get_loss_criterion", "reset_model", "adjust_learning_rate"):
        assert callable(getattr(U.util, name)), name
    assert pca_subspace.elliptical_slice is U.util.elliptical_slice
    m = U.models.MLP(8, 20, 3)
    n_par = sum(p.numel() for p in m.parameters())
    v = torch.arange(n_par, dtype=torch.float32)
    net = U.util.convert_sample_to_net(v, m)
    assert net is not m and torch.equal(torch.cat([p.reshape(-1) for p in net.parameters()]), v)
    assert not torch.equal(torch.cat([p.detach().reshape(-1) for p in m.parameters()]), v)          # the original is untouched
    with pytest.raises(ValueError):
        U.util.convert_sample_to_net(v[:-1].clone(), m)
    bn = torch.nn.BatchNorm1d(4)
    bn.running_mean.fill_(3.0), bn.running_var.fill_(5.0)
    torch.nn.Sequential(bn).apply(U.util.reset_bn)
    assert float(bn.running_mean.abs().max()) == 0.0 and float((bn.running_var - 1).abs().max()) == 0.0
    loss, out, extra = U.util.cross_entropy(torch.nn.Linear(5, 3), torch.zeros(2, 5), torch.tensor([0, 2]))
    assert out.shape == (2, 3) and extra == {} and loss.ndim == 0
