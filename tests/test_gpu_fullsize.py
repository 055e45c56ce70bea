"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle only runs at small sizes):
WideResNet-28-10 flat vectors (D = 36 546 980, configs[2]) for K1 / K2 and S = 100 PreResNet-20 samples on N = 10 000
images (configs[4]) for K3 / K4.  Everything goes through the C ABI."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
D_WRN = 36_546_980


@pytest.fixture(scope="module")
def C():
    from ursabench_b200 import _C
    _C.lib()
    return _C


def test_k1_full_size_identity_linearity_and_noise_moments(C):
    ld = (D_WRN + 3) // 4 * 4
    torch.manual_seed(0)
    p0 = torch.randn(ld, device="cuda")
    g = torch.randn(ld, device="cuda")
    # (1) zero gradient, no weight decay, no noise: parameters come back bit-identical, momentum stays zero
    p, v = p0.clone(), torch.zeros(ld, device="cuda")
    C.sgmcmc_step(p, torch.zeros_like(g), v, lr=0.1, momentum=0.5, wd_over_n=0.0, add_noise=False)
    assert torch.equal(p, p0) and not v.any()
    # (2) SGLD drift against the formula evaluated by torch in fp64 (optim_sghmc.py:47-65): p - lr (g + wd p)
    p = p0.clone()
    C.sgmcmc_step(p, g.clone(), None, lr=0.01, momentum=0.0, wd_over_n=1e-3, add_noise=False)
    ref = p0.double() - 0.01 * (g.double() + 1e-3 * p0.double())
    assert (p.double() - ref).abs().max().item() < 2e-6
    del ref
    # (3) linearity in the noise scale: (p(2s) - p(0)) == 2 (p(s) - p(0)) for the same Philox stream, and the injected
    #     noise has unit variance / zero mean / zero skew over 36.5 M draws
    outs = []
    for mul in (0.0, 1.0, 2.0):
        q = p0.clone()
        C.sgmcmc_step(q, torch.zeros_like(g), None, lr=1.0, momentum=0.0, wd_over_n=0.0, noise_mul=mul, noise_div=1.0,
                      add_noise=mul > 0, seed=7, step=3)
        outs.append(q)
    z1, z2 = outs[1] - outs[0], outs[2] - outs[0]
    assert (z2 - 2 * z1).abs().max().item() < 1e-5
    z = z1[:D_WRN].double()
    n = z.numel()
    assert abs(z.mean().item()) < 5 / math.sqrt(n)
    assert abs(z.var().item() - 1) < 2e-3
    assert abs((z ** 3).mean().item()) < 5e-3 and abs((z ** 4).mean().item() - 3) < 2e-2


def test_k2_full_size_collect_fixed_point_and_draw_contraction(C):
    D, K, S = D_WRN, 20, 30
    ld = (D + 3) // 4 * 4
    torch.manual_seed(1)
    w = torch.randn(ld, device="cuda") * 0.05
    mean, sq = torch.zeros(ld, device="cuda"), torch.zeros(ld, device="cuda")
    ring = torch.empty(K, ld, device="cuda")
    # collecting the same iterate n times is a fixed point: mean == w, sq_mean == w^2, every deviation row == 0
    # (n = 0: mean = w / 1 exactly; later mean * n/(n+1) + w/(n+1) rounds back to within 1 ulp)
    for n in range(3):
        C.swag_collect(w, mean, sq, ring[n], n)
    assert (mean - w).abs().max().item() <= 1e-8 + 2e-7 * w.abs().max().item()
    assert (sq - w * w).abs().max().item() <= 1e-9 + 5e-7 * (w * w).max().item()
    assert ring[0].abs().max().item() == 0.0 and ring[2].abs().max().item() < 1e-7
    var = torch.empty(ld, device="cuda")
    C.swag_variance(mean, sq, var, clamp=1e-30)
    assert var.min().item() >= 1e-30 and var.max().item() < 1e-7          # clamp(sq - mean^2): rounding noise only
    # draw with var = 0: out[s] = mean + ring^T z2[s] / sqrt(K-1) exactly; unit vectors z2 pick single ring rows, so the
    # tensor-core contraction (3xTF32) is checked against the ring itself at full size, ragged last tile included
    ring.normal_(0, 0.02)
    var.zero_()
    z2 = torch.zeros(S, K, device="cuda")
    for s in range(S):
        z2[s, s % K] = 1.0 + s                                            # scaled unit vectors
    out = torch.empty(S, ld, device="cuda")
    rd = math.sqrt(K - 1.0)
    C.swag_draw(out, mean, var, D, ring=ring, z2=z2, rank_div=rd, seed=3, step=1)
    worst = 0.0
    for s in (0, 7, 19, 20, 29):
        ref = mean[:D].double() + ring[s % K, :D].double() * ((1.0 + s) / rd)
        worst = max(worst, (out[s, :D].double() - ref).abs().max().item())
    assert worst < 1e-6, worst                                            # fp32-level: |term| <= 4 sigma * 0.02 * 30 / 4.4
    # Philox z1 with K = 0 and unit variance: out - mean is the documented N(0,1) stream (six normals per block, block
    # (s // 6) * D + d; test_gpu_kernels.py pins it against the oracle's restatement); moments over the full size
    var.fill_(1.0)
    out2 = torch.empty(8, ld, device="cuda")
    C.swag_draw(out2, mean, var, D, seed=9, step=4)
    zs = C.swag_draw_noise(8, D, 9, 4, "cuda")
    assert (out2[:, :D] - mean[None, :D] - zs[:, :D]).abs().max().item() < 1e-6
    zz = zs[:, :D].double()
    assert abs(zz.mean().item()) < 4 / math.sqrt(zz.numel()) and abs(zz.var().item() - 1) < 6 * math.sqrt(2 / zz.numel())
    assert abs((zz ** 4).mean().item() - 3) < 6 * math.sqrt(96 / zz.numel())
    c = torch.corrcoef(zz[:, :: 7])                                                          # draws of one block are uncorrelated
    assert (c - torch.eye(8, device="cuda", dtype=torch.float64)).abs().max().item() < 6 / math.sqrt(D / 7)


def test_k3_k4_full_size_bma_sharding_and_counters(C):
    """S = 100 PreResNet-20 samples on N = 10 000 images: evaluating two halves of the samples separately and adding the
    accumulators (what ranks do before the all-reduce) equals evaluating all of them; rows of the probability sum add up
    to S; the K4 counters agree with torch on the same probabilities."""
    from ursabench_b200 import models
    S, N = 100, 10_000
    torch.manual_seed(2)
    m = models.PreResNet(num_classes=10, depth=20)
    flat = torch.cat([q.detach().reshape(-1) for q in m.parameters()]).cuda()
    bank = (flat[None, :] + 0.02 * torch.randn(S, flat.numel(), device="cuda")).contiguous()
    nbuf = sum(b.numel() for b in m.buffers() if b.dtype == torch.float32)
    bufs = torch.zeros(S, (nbuf + 3) // 4 * 4, device="cuda")
    off = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            bufs[:, off + c:off + 2 * c] = 1.0
            off += 2 * c
    x = torch.randn(N, 3, 32, 32, device="cuda")
    y = torch.randint(0, 10, (N,), device="cuda")
    algo = C.ALGO_TCGEN05_FUSED_F16
    P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
    ws = C.bma_preresnet_forward(bank, bufs, S, x, 20, 10, P, E, algo=algo)
    Pa, Ea = torch.zeros_like(P), torch.zeros_like(E)
    ws = C.bma_preresnet_forward(bank[:37], bufs[:37], 37, x, 20, 10, Pa, Ea, algo=algo, workspace=ws)
    C.bma_preresnet_forward(bank[37:], bufs[37:], 63, x, 20, 10, Pa, Ea, algo=algo, workspace=ws)
    assert bool(torch.isfinite(P).all())
    assert (P - Pa).abs().max().item() / S < 1e-6 and (E - Ea).abs().max().item() / S < 1e-5
    assert (P.sum(1) - S).abs().max().item() < 2e-4
    # spot check against the plain PyTorch fp32 forward of 3 samples on 256 images
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        ref = torch.zeros(256, 10, device="cuda", dtype=torch.float64)
        for s in (0, 50, 99):
            mm = models.PreResNet(num_classes=10, depth=20).cuda().eval()
            torch.nn.utils.vector_to_parameters(bank[s], mm.parameters())
            ref += torch.softmax(mm(x[:256]).double(), -1)
    P3, E3 = torch.zeros(256, 10, device="cuda"), torch.zeros(256, device="cuda")
    idx = torch.tensor([0, 50, 99], device="cuda")
    C.bma_preresnet_forward(bank[idx].contiguous(), bufs[idx].contiguous(), 3, x[:256].contiguous(), 20, 10, P3, E3, algo=algo)
    assert (P3.double() - ref).abs().max().item() / 3 < 1e-5                         # north star
    # K4 on the full [N, C]: integer counters are exact
    oi, of, pred, conf = C.bma_metrics(P, S, y, want_rows=True)
    pbar = P / np.float32(S)
    assert torch.equal(pred.long(), pbar.argmax(1))
    assert int(oi[0]) == int((pbar.argmax(1) == y).sum())
    cnt = oi[1:16]
    assert int(cnt.sum()) == N                                                       # every confidence falls in one bin
    nll = -torch.log((1 - 1e-4) * pbar.double().gather(1, y[:, None]) + 1e-4 / 10).sum().item()
    assert float(of[0]) == pytest.approx(nll, rel=1e-6)


def test_k3_wrn28x10_full_size_accuracy_bracket_and_invariances(C):
    """WideResNet-28-10, C = 100 (configs[2]) at its real widths (160 / 320 / 640: 1 / 2 / 4 output-channel tiles, K up to
    5 760 + 320): (1) against an fp64 forward of the same network the BMA probabilities are no further off than twice what
    PyTorch's own fp32 forward (cuDNN, TF32 off -- the arithmetic the reference runs) is, i.e. the kernel sits inside the fp32
    noise of this ill-conditioned random network; (2) reversing the image order reverses the outputs bit for bit (a pixel's
    result does not depend on its tile neighbours, the 8 x 8 tiles pair different images); (3) the same sample twice doubles
    the accumulators exactly."""
    import copy
    from oracle.wrn_fill import wrn_fill
    from ursabench_b200.models import WideResNet
    Cc, N = 100, 33
    m = wrn_fill(WideResNet(num_classes=Cc, depth=28, widen_factor=10), 5, logit_gain=0.25).cuda().eval()
    row = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    bufs = torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32])
    assert row.numel() == D_WRN
    torch.manual_seed(1)
    x = torch.randn(N, 3, 32, 32, device="cuda")
    bank, bufbank = row[None].contiguous(), bufs[None].contiguous()
    P, E = torch.zeros(N, Cc, device="cuda"), torch.zeros(N, device="cuda")
    ws = C.bma_wrn_forward(bank, bufbank, 1, x, 28, 10, Cc, P, E)
    mm = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
            p32 = torch.softmax(m(x), -1).double()
            p64 = torch.softmax(copy.deepcopy(m).double()(x.double()), -1)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = mm
    err_ours = (P.double() - p64).abs().max().item()
    err_t32 = (p32 - p64).abs().max().item()
    assert err_ours <= max(2.0 * err_t32, 2e-5), (err_ours, err_t32)
    # (2) image permutation
    P2, E2 = torch.zeros_like(P), torch.zeros_like(E)
    C.bma_wrn_forward(bank, bufbank, 1, x.flip(0).contiguous(), 28, 10, Cc, P2, E2, workspace=ws)
    assert torch.equal(P2.flip(0), P) and torch.equal(E2.flip(0), E)
    # (3) the same sample twice
    P3, E3 = torch.zeros_like(P), torch.zeros_like(E)
    C.bma_wrn_forward(torch.cat([bank, bank]), torch.cat([bufbank, bufbank]), 2, x, 28, 10, Cc, P3, E3)
    assert torch.equal(P3, P + P) and torch.equal(E3, E + E)


def test_config3_pipeline_swag_wrn28x10_end_to_end(C):
    """BASELINE.json configs[2] at its real width, shortened in length: SWAG (rank-20 ring) on WideResNet-28-10 / C = 100 ->
    low-rank draws (K2b) -> BatchNorm re-estimation per draw on the conv engine (K3b) -> Prediction over the bank (K3 WideResNet
    forward + K4 metrics).  Checks that every stage ran on the engine and that the numbers are sane and self-consistent."""
    from ursabench_b200 import inference, tasks
    from ursabench_b200.models import WideResNet
    torch.manual_seed(0)
    Cc, n_train, n_test = 100, 256, 130
    xtr, ytr = torch.randn(n_train, 3, 32, 32), torch.randint(0, Cc, (n_train,))
    xte, yte = torch.randn(n_test, 3, 32, 32), torch.randint(0, Cc, (n_test,))
    train = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xtr, ytr), batch_size=128, shuffle=False)
    test = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xte, yte), batch_size=64, shuffle=False)
    hyp = {"lr_init": 0.01, "swag_lr": 0.005, "swag_wd": 5e-4, "momentum": 0.9, "burn_in_epochs": 1, "num_iterates": 3,
           "num_samples": 3, "subspace_type": "covariance"}
    sw = inference.SWAG(hyp, model=WideResNet(num_classes=Cc, depth=28, widen_factor=10), train_loader=train,
                        device=torch.device("cuda"))
    assert sw.num_parameters == D_WRN
    samples = sw.sample(full_cov=True)
    assert len(samples) == 3 and getattr(sw, "_bn_x", None) is not None            # K3b ran (no PyTorch bn_update pass)
    w = sw.bank.w[:3, :D_WRN]
    assert torch.isfinite(w).all() and (w[0] - w[1]).abs().max().item() > 0         # distinct low-rank draws
    b = sw.bank.b[:3]
    assert torch.isfinite(b).all() and (b[0] - b[1]).abs().max().item() > 0         # per-draw BatchNorm statistics
    task = tasks.Prediction({"in_distribution_test": test}, Cc, torch.device("cuda"), "ALL")
    task.update_statistics(samples, output_performance=False)
    assert task.last_engine == "fused_wrn"
    m = task.get_performance_metrics()
    P = task.ensemble_proba
    assert torch.allclose(P.sum(1).cpu(), torch.full((n_test,), 3.0), atol=1e-4)      # three softmax rows per image
    assert 0.0 <= m["error_rate"] <= 1.0 and math.isfinite(m["nll"]) and 0.0 <= m["ece"] <= 1.0 and 0.0 <= m["brier_score"] <= 2.0
    # the same samples through the model's own PyTorch forward (generic engine): the BMA probabilities agree
    ref = tasks.Prediction({"in_distribution_test": test}, Cc, torch.device("cuda"), "ALL", engine="generic")
    ref.update_statistics(samples, output_performance=False)
    assert ref.last_engine == "generic"
    assert (ref.ensemble_proba - P).abs().max().item() < 3e-5 * 3
