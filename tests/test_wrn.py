"""WideResNet BMA forward (BASELINE.json configs[2]: SWAG on WRN-28-10, 30 draws + BMA eval).

CPU: the seeded fill on ``ursabench_b200.models.WideResNet`` reproduces the live reference's logits stored in
tests/golden/prediction_wrn.npz -- this pins parameter / buffer order (the flat bank layout) and the forward definition.
GPU: ``ursa_bma_wrn_forward`` (persistent 3xTF32 tcgen05 implicit GEMM, 1x1 shortcuts folded into conv2's K loop) against
that golden and against a plain PyTorch fp32 forward at widths with several output-channel tiles; ``Prediction`` routes
WideResNet banks through it."""
import os

import numpy as np
import pytest
import torch

from oracle.wrn_fill import wrn_fill
from ursabench_b200.models import WideResNet

GOLD = os.path.join(os.path.dirname(__file__), "golden", "prediction_wrn.npz")


def _models(g, tag):
    depth, widen, C, S, seed = (int(v) for v in g[tag + "/arch"])
    gain = float(g[tag + "/gain"][0])
    ms = [wrn_fill(WideResNet(num_classes=C, depth=depth, widen_factor=widen), seed + s, logit_gain=gain).eval() for s in range(S)]
    return ms, depth, widen, C, S


def _bank(ms, device):
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms]).to(device)
    bufs = torch.stack([torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]) for m in ms]).to(device)
    pad = (-bank.shape[1]) % 4
    if pad:
        bank = torch.nn.functional.pad(bank, (0, pad))
    return bank.contiguous(), bufs.contiguous()


@pytest.mark.parametrize("tag", ["wrn10x2", "wrn16x2"])
def test_seeded_fill_reproduces_reference_logits_cpu(tag):
    g = np.load(GOLD)
    ms, depth, widen, C, S = _models(g, tag)
    assert sum(p.numel() for p in ms[0].parameters()) == int(g[tag + "/D"][0])
    x = torch.from_numpy(g[tag + "/x"].astype(np.float32))
    with torch.no_grad():
        logits = torch.stack([m(x) for m in ms]).numpy()
    np.testing.assert_allclose(logits, g[tag + "/logits"], atol=2e-5, rtol=1e-5)
    p = torch.softmax(torch.from_numpy(logits), -1).sum(0).numpy()
    np.testing.assert_allclose(p, g[tag + "/ensemble_proba"], atol=1e-6)


def test_wrn_workspace_query_rejects_unsupported_shapes():
    from ursabench_b200 import _C
    lib = _C.lib()
    assert lib.ursa_bma_wrn_workspace(1, 16, 28, 10, 100, _C.ALGO_TCGEN05) > 0
    assert lib.ursa_bma_wrn_workspace(1, 16, 10, 2, 10, _C.ALGO_TCGEN05) > 0
    assert lib.ursa_bma_wrn_workspace(1, 16, 28, 10, 100, _C.ALGO_TCGEN05_F16) > 0
    assert lib.ursa_bma_wrn_workspace(1, 16, 28, 10, 100, _C.ALGO_FFMA) == 0        # tcgen05 engines only
    assert lib.ursa_bma_wrn_workspace(1, 16, 27, 10, 100, _C.ALGO_TCGEN05) == 0     # depth != 6n+4
    assert lib.ursa_bma_wrn_workspace(1, 16, 28, 3, 100, _C.ALGO_TCGEN05) == 0      # odd widen factor: widths % 32 != 0
    assert lib.ursa_bma_wrn_workspace(0, 16, 28, 10, 100, _C.ALGO_TCGEN05) == 0


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["ALGO_TCGEN05", "ALGO_TCGEN05_F16"])
@pytest.mark.parametrize("tag", ["wrn10x2", "wrn16x2"])
def test_k3_wrn_forward_matches_reference_golden(tag, engine):
    from ursabench_b200 import _C
    g = np.load(GOLD)
    ms, depth, widen, C, S = _models(g, tag)
    bank, bufs = _bank(ms, "cuda")
    x = torch.from_numpy(g[tag + "/x"].astype(np.float32)).cuda()
    N = x.shape[0]
    P, E = torch.zeros(N, C, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, C, device="cuda")
    _C.bma_wrn_forward(bank, bufs, S, x, depth, widen, C, P, E, logits_out=logits, algo=getattr(_C, engine))
    torch.cuda.synchronize()
    ref = g[tag + "/logits"]
    err = np.abs(logits.cpu().numpy() - ref).max()
    assert err < 1e-4 * max(1.0, np.abs(ref).max()), err
    np.testing.assert_allclose(P.cpu().numpy(), g[tag + "/ensemble_proba"], atol=1e-5, rtol=0)      # north star: 1e-5
    np.testing.assert_allclose(E.cpu().numpy(), g[tag + "/entropy"], atol=2e-5, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["ALGO_TCGEN05", "ALGO_TCGEN05_F16"])
@pytest.mark.parametrize("depth,widen,S,N,Cc", [(10, 10, 2, 9, 100), (16, 4, 2, 6, 10), (10, 6, 1, 515, 10), (10, 2, 3, 1, 10)])
def test_k3_wrn_forward_vs_torch_fp32(depth, widen, S, N, Cc, engine):
    """Widths with 1 / 2 / 4 output-channel tiles (widen 10: 160 / 320 / 640), identity and transition blocks, an odd image
    count (the 8x8 tiles pair two images), N = 1, and N > 512 (image chunking) against PyTorch fp32 (TF32 off)."""
    from ursabench_b200 import _C
    ms = [wrn_fill(WideResNet(num_classes=Cc, depth=depth, widen_factor=widen), 7 * depth + widen + s, logit_gain=0.5).cuda().eval()
          for s in range(S)]
    bank, bufs = _bank(ms, "cuda")
    torch.manual_seed(N)
    x = torch.randn(N, 3, 32, 32, device="cuda")
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    logits = torch.empty(S, N, Cc, device="cuda")
    _C.bma_wrn_forward(bank, bufs, S, x, depth, widen, Cc, P, E, logits_out=logits, algo=getattr(_C, engine))
    torch.cuda.synchronize()
    mm = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
            ref = torch.cat([torch.stack([m(x[i:i + 128]) for m in ms]) for i in range(0, N, 128)], 1)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = mm
    scale = max(1.0, ref.abs().max().item())
    assert (logits - ref).abs().max().item() < 1e-4 * scale
    pbar = torch.softmax(ref.double(), -1).mean(0)
    assert (P.double() / S - pbar).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_prediction_routes_wideresnet_through_the_tcgen05_engine():
    from ursabench_b200.tasks import Prediction
    g = np.load(GOLD)
    ms, depth, widen, C, S = _models(g, "wrn10x2")
    x = torch.from_numpy(g["wrn10x2/x"].astype(np.float32))
    y = torch.from_numpy(g["wrn10x2/y"])
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=4, shuffle=False)
    task = Prediction({"in_distribution_test": loader}, C, torch.device("cuda"), "ALL")
    task.update_statistics(ms, output_performance=False)
    from ursabench_b200 import _C
    assert task.last_engine == "fused_wrn" and task.last_algo == _C.ALGO_TCGEN05_F16      # the FP16-split engine is the product path
    np.testing.assert_allclose(task.ensemble_proba.cpu().numpy(), g["wrn10x2/ensemble_proba"], atol=1e-5, rtol=0)
    # an image beyond fp16's range: NaN from the FP16-split engine, so the evaluation is redone on the 3xTF32 engine
    x2 = x.clone()
    x2[1] *= 3e6
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x2, y), batch_size=4, shuffle=False)
    task = Prediction({"in_distribution_test": loader}, C, torch.device("cuda"), "ALL")
    task.update_statistics(ms, output_performance=False)
    assert task.last_algo == _C.ALGO_TCGEN05 and bool(torch.isfinite(task.ensemble_proba).all())
    keep = [i for i in range(x.shape[0]) if i != 1]
    np.testing.assert_allclose(task.ensemble_proba.cpu().numpy()[keep], g["wrn10x2/ensemble_proba"][keep], atol=1e-5, rtol=0)


# ------------------------------------------------------------------------------------------------ BatchNorm re-estimation
def _flat_buffers(m):
    return torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32])


def _flat_buffers_any(m):
    return torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype.is_floating_point])


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["ALGO_TCGEN05", "ALGO_TCGEN05_F16"])
@pytest.mark.parametrize("depth,widen,N,batch", [(10, 2, 300, 128), (16, 4, 70, 32), (10, 10, 40, 16), (10, 2, 1030, 256)])
def test_wrn_bn_update_matches_torch_train_mode(depth, widen, N, batch, engine):
    """ursa_wrn_bn_update (train-mode pass on the tcgen05 conv kernel, fp64 batch sums) against util.bn_update's PyTorch pass
    (reference util.py:212-247): ragged last batch, several chunks (N > 512), identity and transition blocks."""
    from ursabench_b200 import _C
    from ursabench_b200.util import bn_update
    m = wrn_fill(WideResNet(num_classes=10, depth=depth, widen_factor=widen), 31 * depth + widen).cuda()
    torch.manual_seed(N)
    x = torch.randn(N, 3, 32, 32)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.zeros(N, dtype=torch.long)), batch_size=batch,
                                         shuffle=False)
    row = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).contiguous()
    import copy
    m64 = copy.deepcopy(m).double()
    loader64 = [(xb.double(), yb) for xb, yb in loader]
    bn_update(loader64, m64, device=torch.device("cuda"))                 # the exact statistics of this network
    exact = _flat_buffers_any(m64)
    mm = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
            bn_update(loader, m, device=torch.device("cuda"))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = mm
    ref = _flat_buffers(m)
    buf = torch.full_like(ref, float("nan"))                       # the kernel must overwrite every statistic
    ws = _C.wrn_bn_update(row, buf, x.cuda(), batch, depth, widen, 10, algo=getattr(_C, engine))
    torch.cuda.synchronize()
    assert ws is not None
    assert torch.isfinite(buf).all()
    # errors in units that mean something: a running mean against the layer's standard deviation, a running variance
    # relatively.  The tensor core's truncated accumulation is a SYSTEMATIC ~1e-6 relative offset of every conv output (it does
    # not average out over the pixels like fp32 round-to-nearest does), so the means sit ~10x further from the exact value than
    # PyTorch's -- a few 1e-6 standard deviations.
    e_mean = e_var = t_mean = t_var = 0.0
    off = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            mu, var = exact[off:off + c], exact[off + c:off + 2 * c]
            sd = var.sqrt()
            e_mean = max(e_mean, ((buf[off:off + c].double() - mu).abs() / sd).max().item())
            e_var = max(e_var, ((buf[off + c:off + 2 * c].double() - var).abs() / var).max().item())
            t_mean = max(t_mean, ((ref[off:off + c].double() - mu).abs() / sd).max().item())
            t_var = max(t_var, ((ref[off + c:off + 2 * c].double() - var).abs() / var).max().item())
            off += 2 * c
    assert off == buf.numel()
    assert e_mean < 2e-5 and e_var < 5e-5, (e_mean, e_var, t_mean, t_var)


@pytest.mark.gpu
def test_swag_sample_uses_the_engine_bn_update_for_wideresnets():
    """SWAG.sample on a WideResNet: every returned sample carries BatchNorm statistics re-estimated by ursa_wrn_bn_update
    (no PyTorch pass), equal to what util.bn_update computes for the same weights."""
    from ursabench_b200 import inference
    from ursabench_b200.util import bn_update
    torch.manual_seed(0)
    N = 96
    x, y = torch.randn(N, 3, 32, 32), torch.randint(0, 10, (N,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)
    hyp = {"lr_init": 0.01, "swag_lr": 0.005, "swag_wd": 1e-4, "momentum": 0.9, "burn_in_epochs": 1, "num_iterates": 2,
           "num_samples": 2}
    model = WideResNet(num_classes=10, depth=10, widen_factor=2)
    sw = inference.SWAG(hyp, model=model, train_loader=loader, device=torch.device("cuda"))
    samples = sw.sample()
    assert len(samples) == 2 and getattr(sw, "_bn_x", None) is not None          # the engine path ran
    for i in range(2):
        ref_m = WideResNet(num_classes=10, depth=10, widen_factor=2).cuda()
        off = 0
        for p in ref_m.parameters():
            p.data.copy_(sw.bank.w[i, off:off + p.numel()].view(p.shape))
            off += p.numel()
        with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
            bn_update(loader, ref_m, device=torch.device("cuda"))
        ref = _flat_buffers(ref_m)
        got = sw.bank.b[i, :ref.numel()]
        assert ((got - ref).abs() / (ref.abs() + 1e-2)).max().item() < 5e-4


def test_arch_detection_accepts_the_reference_modules():
    """The engine recognises WideResNets structurally (class / attribute names), so the reference's OWN module objects -- what
    `inference.sample()` of the reference returns -- route through the tcgen05 forward too.  Needs the reference checkout
    (absent on the GPU box): skipped there."""
    from oracle import stubs
    if not stubs.reference_available():
        pytest.skip("reference checkout not present")
    stubs.install()
    import warnings
    from URSABench import models as ref_models
    from ursabench_b200.tasks._engine import _arch_of
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_models.wideresnet.WideResNet(num_classes=10, depth=16, widen_factor=4)
    assert _arch_of(ref) == ("wrn", 16, 4, 10)
    ours = WideResNet(num_classes=10, depth=16, widen_factor=4)
    assert _arch_of(ours) == ("wrn", 16, 4, 10)
    assert [tuple(p.shape) for p in ref.parameters()] == [tuple(p.shape) for p in ours.parameters()]
    assert [n for n, _ in ref.named_buffers()] == [n for n, _ in ours.named_buffers()]


# ---- bn_update pinned to the live reference (tests/golden/bn_update.npz, oracle/gen_golden.py::gen_bn_update) -----------------
BN_GOLD = os.path.join(os.path.dirname(__file__), "golden", "bn_update.npz")


def _bn_golden_problem():
    g = np.load(BN_GOLD)
    depth, widen, C, N, batch, seed = (int(v) for v in g["cfg"])
    m = wrn_fill(WideResNet(num_classes=C, depth=depth, widen_factor=widen), seed)
    x = torch.from_numpy(g["x"].astype(np.float32))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.zeros(N, dtype=torch.long)), batch_size=batch,
                                         shuffle=False)
    return g, m, x, loader, (depth, widen, C, N, batch)


def test_bn_update_port_reproduces_the_reference_cpu():
    """ursabench_b200.util.bn_update (the PyTorch pass, here on the CPU) against the running statistics the LIVE reference's
    util.bn_update produced for the same weights and batches; the modules' momenta are restored like the reference does."""
    from ursabench_b200.util import bn_update
    g, m, x, loader, _ = _bn_golden_problem()
    bn_update(loader, m, device=torch.device("cpu"))
    got = _flat_buffers(m).numpy()
    np.testing.assert_allclose(got, g["buffers"], rtol=2e-5, atol=2e-6)
    mom = [mod.momentum for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    np.testing.assert_array_equal(np.array(mom, np.float64), g["momentum_after"])


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["ALGO_TCGEN05", "ALGO_TCGEN05_F16"])
def test_wrn_bn_update_matches_reference_golden(engine):
    """ursa_wrn_bn_update against the same golden: means within 2e-5 of the layer's standard deviation, variances within 5e-5
    relative (the tensor core's systematic ~1e-6 offset, see the fp64-bracket test above)."""
    from ursabench_b200 import _C
    g, m, x, loader, (depth, widen, C, N, batch) = _bn_golden_problem()
    row = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).cuda().contiguous()
    ref = torch.from_numpy(g["buffers"]).cuda()
    buf = torch.full_like(ref, float("nan"))
    assert _C.wrn_bn_update(row, buf, x.cuda(), batch, depth, widen, C, algo=getattr(_C, engine)) is not None
    off = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            mu, var = ref[off:off + c].double(), ref[off + c:off + 2 * c].double()
            assert ((buf[off:off + c].double() - mu).abs() / var.sqrt()).max().item() < 2e-5
            assert ((buf[off + c:off + 2 * c].double() - var).abs() / var).max().item() < 5e-5
            off += 2 * c
    assert off == buf.numel()
