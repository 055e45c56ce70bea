"""WRN BMA forward timing on one GPU: python tools/bench_bma_wrn.py [depth widen C S N] [--ref] [--f16]
Prints one JSON line: ms, img*samples/s, fp32-equivalent TFLOP/s (2*MAC of convs + linear), and with --ref the
per-sample PyTorch fp32 forward (cuDNN, TF32 off = the generic engine this kernel replaces) beside it."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402
from ursabench_b200.models import WideResNet  # noqa: E402


def wrn_fill(model, seed, logit_gain=4.0):
    """Seeded weights with O(1) activations and non-trivial BatchNorm statistics (a local twin of the test suite's filler:
    tools do not import the oracle package)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 4:
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / (p.shape[1] * p.shape[2] * p.shape[3])) ** 0.5)
            elif p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * (logit_gain / p.shape[1] ** 0.5))
            elif name.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.3)
            elif name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) * 1.5 + 0.5)
    return model


def wrn_flops(depth, widen):
    n = (depth - 4) // 6
    w = [16, 16 * widen, 32 * widen, 64 * widen]
    fl = 2 * 1024 * 27 * 16
    inp, hw = 16, 32
    for g, stride in enumerate((1, 2, 2)):
        for b in range(n):
            s = stride if b == 0 else 1
            cout = w[g + 1]
            fl += 2 * hw * hw * 9 * inp * cout
            ho = hw // s
            fl += 2 * ho * ho * 9 * cout * cout
            if s != 1 or inp != cout:
                fl += 2 * ho * ho * inp * cout
            inp, hw = cout, ho
    return fl


def bench_bn(depth, widen, C, N, batch=128):
    """ursa_wrn_bn_update vs. util.bn_update's PyTorch pass (cuDNN fp32) for one sample over N training images."""
    from ursabench_b200.util import bn_update
    m = wrn_fill(WideResNet(num_classes=C, depth=depth, widen_factor=widen), 0, logit_gain=0.25).cuda()
    row = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).contiguous()
    nbuf = sum(b.numel() for b in m.buffers() if b.dtype == torch.float32)
    buf = torch.zeros(nbuf, device="cuda")
    torch.manual_seed(0)
    x = torch.randn(N, 3, 32, 32, device="cuda")
    algo = _C.ALGO_TCGEN05_F16 if "--f16" in sys.argv else _C.ALGO_TCGEN05
    ws = _C.wrn_bn_update(row, buf, x[:batch * 2], batch, depth, widen, C, algo=algo)
    ws = _C.wrn_bn_update(row, buf, x, batch, depth, widen, C, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _C.wrn_bn_update(row, buf, x, batch, depth, widen, C, workspace=ws, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    loader = [(x[i:i + batch], None) for i in range(0, N, batch)]
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        bn_update(loader, m, device=torch.device("cuda"))
        torch.cuda.synchronize()
        e0.record()
        bn_update(loader, m, device=torch.device("cuda"))
        e1.record()
        torch.cuda.synchronize()
    tr = e0.elapsed_time(e1)
    ref = torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32])
    fl = wrn_flops(depth, widen)
    print(json.dumps({"bn_update": True, "depth": depth, "widen": widen, "N": N, "batch": batch, "ms": t, "img_per_s": N / t * 1e3,
                      "TFLOPs_fp32_equiv": fl * N / t / 1e9, "torch_fp32_ms": tr, "speedup": tr / t,
                      "max_rel_err_vs_torch": ((buf - ref).abs() / (ref.abs() + 1e-2)).max().item()}))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    depth, widen, C, S, N = (int(v) for v in (args + [28, 10, 100, 2, 1024][len(args):]))
    if "--bn" in sys.argv:
        return bench_bn(depth, widen, C, N)
    ms = [wrn_fill(WideResNet(num_classes=C, depth=depth, widen_factor=widen), s, logit_gain=0.25).cuda().eval() for s in range(S)]
    bank = torch.stack([torch.cat([p.detach().reshape(-1) for p in m.parameters()]) for m in ms])
    bufs = torch.stack([torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]) for m in ms])
    torch.manual_seed(0)
    x = torch.randn(N, 3, 32, 32, device="cuda")
    P, E = torch.zeros(N, C, device="cuda"), torch.zeros(N, device="cuda")
    algo = _C.ALGO_TCGEN05_F16 if "--f16" in sys.argv else _C.ALGO_TCGEN05
    ws = _C.bma_wrn_forward(bank, bufs, 1, x[:min(N, 64)], depth, widen, C, P[:min(N, 64)], E[:min(N, 64)], algo=algo)   # warm-up
    ws = _C.bma_wrn_forward(bank, bufs, S, x, depth, widen, C, P, E, algo=algo)
    P.zero_(), E.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _C.bma_wrn_forward(bank, bufs, S, x, depth, widen, C, P, E, workspace=ws, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    fl = wrn_flops(depth, widen) + 2 * 64 * widen * C
    out = {"engine": "f16" if "--f16" in sys.argv else "3xtf32", "depth": depth, "widen": widen, "C": C, "S": S, "N": N, "ms": t, "img_samples_per_s": S * N / t * 1e3,
           "TFLOPs_fp32_equiv": fl * S * N / t / 1e9, "MFLOP_per_img": fl / 1e6}
    if "--ref" in sys.argv:
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
            refs = [torch.cat([m(x[i:i + 128]) for i in range(0, N, 128)]) for m in ms]           # warm-up + reference
            torch.cuda.synchronize()
            e0.record()
            refs = [torch.cat([m(x[i:i + 128]) for i in range(0, N, 128)]) for m in ms]
            e1.record()
            torch.cuda.synchronize()
        tr = e0.elapsed_time(e1)
        pref = torch.softmax(torch.stack(refs).double(), -1).sum(0)
        import copy
        nb = min(N, 32)                                      # fp64 forward on a few images: who is closer to the exact network?
        with torch.no_grad():
            l64 = torch.stack([copy.deepcopy(m).double()(x[:nb].double()) for m in ms])
        p64 = torch.softmax(l64, -1).sum(0) / S
        out.update({"proba_err_vs_fp64_ours": (P[:nb].double() / S - p64).abs().max().item(),
                    "proba_err_vs_fp64_torch_fp32": (pref[:nb] / S - p64).abs().max().item(),
                    "logit_absmax": l64.abs().max().item()})
        out.update({"torch_fp32_ms": tr, "torch_fp32_TFLOPs": fl * S * N / tr / 1e9, "speedup": tr / t,
                    "max_abs_proba_err": (P.double() - pref).abs().max().item() / S})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
