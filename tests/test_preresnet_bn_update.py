"""PreResNet BatchNorm re-estimation on the engine (``ursa_preresnet_bn_update``) against goldens of the LIVE
reference's ``util.bn_update`` on its own PreResNet (oracle/gen_golden.py::gen_bn_update_preresnet) and against PyTorch's
train-mode pass in fp64; SWAG.sample uses it, sample-batched, for all draws at once."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle.wrn_fill import wrn_fill

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bn_update_preresnet.npz")


def _flat_buffers(m):
    return torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype.is_floating_point])


def _errors(m, buf, exact):
    """(max |mean error| / sigma, max relative variance error) over all BatchNorm layers."""
    e_mean = e_var = 0.0
    off = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            mu, var = exact[off:off + c].double(), exact[off + c:off + 2 * c].double()
            e_mean = max(e_mean, ((buf[off:off + c].double() - mu).abs() / var.sqrt()).max().item())
            e_var = max(e_var, ((buf[off + c:off + 2 * c].double() - var).abs() / var).max().item())
            off += 2 * c
    assert off == buf.numel()
    return e_mean, e_var


@pytest.mark.parametrize("tag", ["d8", "d20"])
def test_port_bn_update_reproduces_reference_golden(tag):
    """CPU: our ``util.bn_update`` on our PreResNet definition reproduces the live reference's running statistics."""
    from ursabench_b200.models import PreResNet
    from ursabench_b200.util import bn_update
    g = np.load(GOLD)
    depth, C, N, batch, seed = (int(v) for v in g[tag + "/cfg"])
    m = wrn_fill(PreResNet(num_classes=C, depth=depth), seed, logit_gain=0.4)
    x = torch.from_numpy(g[tag + "/x"].astype(np.float32))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.zeros(N, dtype=torch.long)), batch_size=batch)
    bn_update(loader, m, device=torch.device("cpu"))
    e_mean, e_var = _errors(m, _flat_buffers(m), torch.from_numpy(g[tag + "/buffers"]))
    assert e_mean < 2e-5 and e_var < 5e-5, (e_mean, e_var)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["d8", "d20"])
def test_engine_bn_update_matches_reference_golden(tag):
    from ursabench_b200 import _C
    from ursabench_b200.models import PreResNet
    g = np.load(GOLD)
    depth, C, N, batch, seed = (int(v) for v in g[tag + "/cfg"])
    m = wrn_fill(PreResNet(num_classes=C, depth=depth), seed, logit_gain=0.4)
    row = torch.cat([p.detach().reshape(-1) for p in m.parameters()])[None].cuda().contiguous()
    buf = torch.full((1, g[tag + "/buffers"].size), float("nan"), device="cuda")
    ws = _C.preresnet_bn_update(row, buf, torch.from_numpy(g[tag + "/x"].astype(np.float32)).cuda(), batch, depth, C)
    assert ws is not None and bool(torch.isfinite(buf).all())               # every statistic overwritten
    e_mean, e_var = _errors(m, buf[0].cpu(), torch.from_numpy(g[tag + "/buffers"]))
    print(tag, "mean error / sigma %.2e, variance rel. error %.2e" % (e_mean, e_var))
    assert e_mean < 2e-5 and e_var < 5e-5, (e_mean, e_var)


@pytest.mark.gpu
@pytest.mark.parametrize("depth,C,N,batch,S", [(8, 10, 300, 128, 3), (20, 100, 70, 32, 2), (14, 10, 1030, 256, 9)])
def test_engine_bn_update_sample_batched_vs_fp64_train_mode(depth, C, N, batch, S):
    """S samples in one call (9 > the 8-sample launch chunk), several image chunks (N > 512), ragged last batch: every sample's
    statistics against the exact (fp64) train-mode pass of util.bn_update on that sample."""
    from ursabench_b200 import _C
    from ursabench_b200.models import PreResNet
    from ursabench_b200.util import bn_update
    ms = [wrn_fill(PreResNet(num_classes=C, depth=depth), 77 * depth + s, logit_gain=0.4) for s in range(S)]
    torch.manual_seed(N)
    x = torch.randn(N, 3, 32, 32)
    loader64 = [(xb.double().cuda(), yb) for xb, yb in
                torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.zeros(N, dtype=torch.long)), batch_size=batch)]
    rows = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms]).cuda().contiguous()
    nbuf = _flat_buffers(ms[0]).numel()
    bufs = torch.full((S, (nbuf + 3) // 4 * 4), float("nan"), device="cuda")
    ws = _C.preresnet_bn_update(rows, bufs, x.cuda(), batch, depth, C)
    assert ws is not None
    for s, m in enumerate(ms):
        m64 = copy.deepcopy(m).double().cuda()
        bn_update(loader64, m64, device=torch.device("cuda"))
        e_mean, e_var = _errors(m, bufs[s, :nbuf].cpu(), _flat_buffers(m64).cpu())
        assert e_mean < 2e-5 and e_var < 5e-5, (s, e_mean, e_var)
    assert _C.lib().ursa_preresnet_bn_update_workspace(2, 100, 127, 20, 10) == 0      # odd batch: not covered


@pytest.mark.gpu
def test_swag_sample_uses_the_engine_bn_update_for_preresnets():
    """SWAG.sample on the north-star architecture: all draws get their BatchNorm statistics from ONE sample-batched engine pass
    (no PyTorch pass), equal to what util.bn_update computes for the same weights."""
    from ursabench_b200 import inference
    from ursabench_b200.models import PreResNet
    from ursabench_b200.util import bn_update
    torch.manual_seed(0)
    N = 96
    x, y = torch.randn(N, 3, 32, 32), torch.randint(0, 10, (N,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=32, shuffle=False)
    hyp = {"lr_init": 0.01, "swag_lr": 0.005, "swag_wd": 1e-4, "momentum": 0.9, "burn_in_epochs": 1, "num_iterates": 2,
           "num_samples": 3}
    sw = inference.SWAG(hyp, model=PreResNet(num_classes=10, depth=8), train_loader=loader, device=torch.device("cuda"))
    calls = []
    from ursabench_b200 import _C
    orig = _C.preresnet_bn_update
    _C.preresnet_bn_update = lambda *a, **k: (calls.append(a[0].shape[0]), orig(*a, **k))[1]
    try:
        samples = sw.sample()
    finally:
        _C.preresnet_bn_update = orig
    assert len(samples) == 3 and calls == [3]                                # one sample-batched call for the three draws
    for h in samples:
        ref_m = PreResNet(num_classes=10, depth=8).cuda()
        torch.nn.utils.vector_to_parameters(sw.bank.w[h._ursa_row, :sw.flat.D].clone(), ref_m.parameters())
        ref_m = ref_m.double()                               # the exact pass (cuDNN's default fp32 convs run on TF32)
        bn_update([(xb.double(), yb) for xb, yb in loader], ref_m, device=torch.device("cuda"))
        got = sw.bank.b[h._ursa_row, :sw.flat.nb].cpu()
        e_mean, e_var = _errors(ref_m, got, _flat_buffers(ref_m).cpu())
        assert e_mean < 2e-5 and e_var < 5e-5, (e_mean, e_var)
