"""PreResNet-20 -- the north-star model -- against goldens produced by the LIVE reference
(oracle/gen_golden.py::gen_prediction_preresnet20: the reference's own ``PreResNet(depth=20)`` and ``Prediction`` on the CPU)
at the logit scale of a trained network (|logit| ~ 22 for C = 10, ~ 35 for C = 100).  Weights are re-created on both sides
from the seeded stream of oracle/wrn_fill.py in flat-layout order, so a layout mismatch shows up as a parity failure.

CPU: our module definition reproduces the reference's logits.  GPU: every engine, and ``tasks.Prediction`` with the
engine it picks by itself, meets the north star's 1e-5 on the BMA probabilities."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.wrn_fill import wrn_fill

GOLD = os.path.join(os.path.dirname(__file__), "golden", "prediction_preresnet20.npz")
TAGS = ["c10", "c100"]


def _models(g, tag):
    from ursabench_b200.models import PreResNet
    depth, C, S, seed = (int(v) for v in g[tag + "/arch"])
    gain = float(g[tag + "/gain"][0])
    return [wrn_fill(PreResNet(num_classes=C, depth=depth), seed + s, logit_gain=gain).eval() for s in range(S)], depth, C, S


@pytest.mark.parametrize("tag", TAGS)
def test_module_definition_reproduces_reference_logits(tag):
    g = np.load(GOLD)
    ms, depth, C, S = _models(g, tag)
    x = torch.from_numpy(g[tag + "/x"].astype(np.float32))
    with torch.no_grad():
        logits = torch.stack([m(x) for m in ms]).numpy()
    assert float(np.abs(g[tag + "/logits"]).max()) > 20.0                     # the fixture really is at trained-network scale
    np.testing.assert_allclose(logits, g[tag + "/logits"], rtol=2e-5, atol=2e-4)
    proba = torch.softmax(torch.from_numpy(logits), -1).sum(0).numpy()
    np.testing.assert_allclose(proba, g[tag + "/ensemble_proba"], atol=2e-6, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("algo_name", ["ffma", "tcgen05_fused", "tcgen05_fused_f16"])
def test_engines_meet_1e5_on_reference_golden(tag, algo_name):
    from ursabench_b200 import _C
    algo = {"ffma": _C.ALGO_FFMA, "tcgen05_fused": _C.ALGO_TCGEN05_FUSED, "tcgen05_fused_f16": _C.ALGO_TCGEN05_FUSED_F16}[algo_name]
    g = np.load(GOLD)
    ms, depth, C, S = _models(g, tag)
    dev = torch.device("cuda")
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms]).to(dev)
    bufs = torch.stack([torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]) for m in ms]).to(dev)
    x = torch.from_numpy(g[tag + "/x"].astype(np.float32)).to(dev)
    P, E = torch.zeros(x.shape[0], C, device=dev), torch.zeros(x.shape[0], device=dev)
    _C.bma_preresnet_forward(bank, bufs, S, x, depth, C, P, E, algo=algo)
    err = float(np.abs(P.cpu().numpy() - g[tag + "/ensemble_proba"]).max()) / S     # per-sample probability error
    print("%s %s: max |p - p_ref| = %.2e (reference fp32 vs fp64: %.1e)" % (tag, algo_name, err, float(g[tag + "/fp32_vs_fp64_proba"][0])))
    assert err <= 1e-5, err
    np.testing.assert_allclose(E.cpu().numpy(), g[tag + "/entropy"], atol=5e-5 * S, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_prediction_class_on_reference_golden(tag):
    import ursabench_b200 as U
    g = np.load(GOLD)
    ref = json.load(open(GOLD.replace(".npz", "_metrics.json")))[tag]
    ms, depth, C, S = _models(g, tag)
    x, y = torch.from_numpy(g[tag + "/x"].astype(np.float32)), torch.from_numpy(g[tag + "/y"])
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8, shuffle=False)
    task = U.tasks.Prediction({"in_distribution_test": loader}, C, torch.device("cuda"), ["error_rate", "nll", "brier_score", "ece"])
    task.update_statistics(ms, output_performance=False)
    assert task.last_engine == "fused_preresnet" and task.last_algo == U._C.ALGO_TCGEN05_FUSED_F16
    np.testing.assert_allclose(task.ensemble_proba.numpy(), g[tag + "/ensemble_proba"], atol=1e-5 * S, rtol=0)
    got = task.get_performance_metrics()
    assert got["error_rate"] == pytest.approx(ref["error_rate"], abs=1e-12)
    for k in ("nll", "brier_score", "ece"):
        assert got[k] == pytest.approx(ref[k], abs=2e-5, rel=2e-5), k
