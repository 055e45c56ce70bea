"""Kernel micro-benchmarks (CUDA events, inputs larger than L2 unless noted).  Development aid; bench.py is the
contract benchmark.  Usage: python tools/microbench.py [k1|k2|k3|all] [--json out.json]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402

PEAK = 6549.8
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def k1(res):
    D = 36_546_980
    p, g, v, z = (torch.randn(D, device="cuda") for _ in range(4))
    snap = torch.empty_like(p)
    lr, N = 0.01, 50000.0
    for name, kw, bpp in [
        ("sghmc_philox", dict(v=v, momentum=0.5), 20),
        ("sghmc_nonoise", dict(v=v, momentum=0.5, add_noise=False), 20),
        ("sghmc_extnoise", dict(v=v, momentum=0.5, noise=z), 24),
        ("sghmc_philox_snapshot", dict(v=v, momentum=0.5, snapshot=snap), 24),
        ("sghmc_philox_zerograd", dict(v=v, momentum=0.5, zero_grad=True), 24),
        ("sgld_philox", dict(momentum=0.0), 12),
        ("sgld_nonoise", dict(momentum=0.0, add_noise=False), 12),
    ]:
        mom = kw["momentum"]
        args = dict(lr=lr, wd_over_n=1e-4, noise_mul=math.sqrt(2 * (1 - mom) * lr), noise_div=N, seed=1)
        args.update(kw)
        vv = args.pop("v", None)
        sn = args.pop("snapshot", None)
        nz = args.pop("noise", None)
        st = [0]

        def fn():
            st[0] += 1
            _C.sgmcmc_step(p, g, vv, sn, nz, step=st[0], **args)
        med, best = timeit(fn)
        gbs = D * bpp / med / 1e6
        res["k1_" + name] = dict(ms=med, ms_best=best, GBps=gbs, frac=gbs / PEAK, bytes_per_param=bpp, D=D)
        print("K1 %-24s %.3f ms  %.0f GB/s  (%.1f%% of %.0f)" % (name, med, gbs, 100 * gbs / PEAK, PEAK), flush=True)
    # copy-kernel yardstick on the same buffers
    med, _ = timeit(lambda: snap.copy_(p))
    print("torch copy_ D floats: %.3f ms  %.0f GB/s" % (med, D * 8 / med / 1e6))
    res["copy_yardstick_GBps"] = D * 8 / med / 1e6
    # small-D (PreResNet-20), L2 resident, and chain batched
    for D2, label in ((272_282, "preresnet20_1chain"), (272_284 * 128, "preresnet20_128chains")):
        p2, g2, v2 = (torch.randn(D2, device="cuda") for _ in range(3))
        st = [0]

        def fn2():
            st[0] += 1
            _C.sgmcmc_step(p2, g2, v2, lr=lr, momentum=0.5, wd_over_n=1e-4, noise_mul=0.1, noise_div=N, seed=1, step=st[0])
        med, best = timeit(fn2, iters=50)
        gbs = D2 * 20 / med / 1e6
        res["k1_" + label] = dict(ms=med, ms_best=best, GBps=gbs, frac=gbs / PEAK, D=D2)
        print("K1 %-24s %.4f ms  %.0f GB/s" % (label, med, gbs), flush=True)


def k2(res):
    D = 36_546_980
    ld = (D + 3) // 4 * 4
    w, mean, sq = (torch.randn(D, device="cuda") * 0.05 for _ in range(3))
    K, S = 20, 30
    ring = torch.randn(K, ld, device="cuda") * 0.01
    st = [0]

    def fc():
        st[0] += 1
        _C.swag_collect(w, mean, sq, ring[st[0] % K], st[0])
    med, best = timeit(fc)
    gbs = D * 24 / med / 1e6
    res["k2_collect"] = dict(ms=med, GBps=gbs, frac=gbs / PEAK)
    print("K2 collect  %.3f ms  %.0f GB/s (%.1f%%)" % (med, gbs, 100 * gbs / PEAK), flush=True)
    var = torch.empty(D, device="cuda")
    med, _ = timeit(lambda: _C.swag_variance(mean, sq.abs() + 1, var))
    _C.swag_variance(mean, mean * mean + 0.01, var)
    out = torch.empty(S, ld, device="cuda")
    z2 = torch.randn(S, K, device="cuda")
    for s_, k_ in ((30, 20), (30, 0), (8, 20), (1, 20)):
        o = out[:s_]
        r = ring[:k_] if k_ else None
        zz = z2[:s_, :k_].contiguous() if k_ else None
        med, best = timeit(lambda: _C.swag_draw(o, mean, var, D, ring=r, z2=zz, rank_div=math.sqrt(19.0), seed=3, step=1),
                           iters=10, warm=3)
        bytes_ = D * 4 * (k_ + 2 + s_)
        gbs = bytes_ / med / 1e6
        res["k2_draw_S%d_K%d" % (s_, k_)] = dict(ms=med, GBps=gbs, frac=gbs / PEAK)
        print("K2 draw S=%d K=%d  %.3f ms  %.0f GB/s (%.1f%%)" % (s_, k_, med, gbs, 100 * gbs / PEAK), flush=True)


def k3(res):
    from ursabench_b200.models import MLP
    S, N = 100, 10000
    torch.manual_seed(0)
    m = MLP(400, 784, 10)
    D = sum(p.numel() for p in m.parameters())
    bank = torch.randn(S, D, device="cuda") * 0.05
    x = torch.randn(N, 784, device="cuda")
    P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
    ws = [None]
    for algo, nm in ((_C.ALGO_FFMA, "ffma"), (_C.ALGO_TCGEN05, "tcgen05")):
        try:
            def f():
                ws[0] = _C.bma_mlp_forward(bank, S, x, 784, 400, 10, P, E, algo=algo, workspace=ws[0])
            med, best = timeit(f, iters=5, warm=2)
        except Exception as e:  # noqa: BLE001
            print("K3 mlp %s: %s" % (nm, e))
            ws[0] = None
            continue
        flops = 2 * (784 * 400 + 400 * 400 + 400 * 10) * S * N
        res["k3_mlp_" + nm] = dict(ms=med, TFLOPs=flops / med / 1e9, img_samples_per_s=S * N / med * 1e3)
        print("K3 mlp %-8s S=%d N=%d  %.2f ms  %.1f TFLOP/s  %.2f M img*samples/s" %
              (nm, S, N, med, flops / med / 1e9, S * N / med / 1e3), flush=True)
        ws[0] = None


def k3conv(res):
    from ursabench_b200.models import PreResNet
    S, N = 16, 2048
    torch.manual_seed(0)
    m = PreResNet(num_classes=10, depth=20)
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).cuda()
    bank = (flat[None, :] + 0.01 * torch.randn(S, flat.numel(), device="cuda")).contiguous()
    bufs = torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]).cuda()[None, :].repeat(S, 1).contiguous()
    x = torch.randn(N, 3, 32, 32, device="cuda")
    P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
    ws = [None]
    outs = {}
    for algo, nm in ((_C.ALGO_FFMA, "ffma"), (_C.ALGO_TCGEN05, "tcgen05")):
        try:
            lg = torch.empty(S, N, 10, device="cuda")

            def f():
                ws[0] = _C.bma_preresnet_forward(bank, bufs, S, x, 20, 10, P, E, logits_out=lg, algo=algo, workspace=ws[0])
            med, best = timeit(f, iters=3, warm=1)
            outs[nm] = lg
        except Exception as e:  # noqa: BLE001
            print("K3 preresnet20 %s: %s" % (nm, e))
            ws[0] = None
            continue
        flops = 81.63e6 * S * N
        res["k3_preresnet20_" + nm] = dict(ms=med, TFLOPs=flops / med / 1e9, img_samples_per_s=S * N / med * 1e3)
        print("K3 preresnet20 %-8s S=%d N=%d  %.2f ms  %.1f TFLOP/s  %.1f k img*samples/s" %
              (nm, S, N, med, flops / med / 1e9, S * N / med), flush=True)
        ws[0] = None
    if len(outs) == 2:
        print("max |logit diff| tcgen05 vs ffma: %.3e" % (outs["ffma"] - outs["tcgen05"]).abs().max().item())


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "all"
    res = {}
    print(torch.cuda.get_device_name(0), _C.device_info())
    if which in ("k1", "all"):
        k1(res)
    if which in ("k2", "all"):
        k2(res)
    if which in ("k3", "all"):
        k3(res)
    if which in ("k3conv", "all"):
        k3conv(res)
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
