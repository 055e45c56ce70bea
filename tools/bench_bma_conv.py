"""Time the sample-batched PreResNet-20 BMA forward (BASELINE.json configs[4] shape) for each algo.
Usage: python tools/bench_bma_conv.py [S] [N] [algos...]   (development aid; bench.py is the contract benchmark)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C, models  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
algos = sys.argv[3:] or ["tcgen05", "fused", "fused16"]
ALGO = {"ffma": _C.ALGO_FFMA, "tcgen05": _C.ALGO_TCGEN05, "fused": _C.ALGO_TCGEN05_FUSED,
        "fused16": _C.ALGO_TCGEN05_FUSED_F16}
dev = torch.device("cuda")
torch.manual_seed(0)
m = models.PreResNet(num_classes=10, depth=20)
flat = torch.cat([q.detach().reshape(-1) for q in m.parameters()]).to(dev)
Dp = flat.numel()
nbuf = sum(b.numel() for b in m.buffers() if b.dtype == torch.float32)
bank = (flat[None, :] + 0.01 * torch.randn(S, Dp, device=dev)).contiguous()
buf = torch.zeros(S, (nbuf + 3) // 4 * 4, device=dev)
off = 0
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        c = mod.num_features
        buf[:, off + c:off + 2 * c] = 1.0
        off += 2 * c
x = torch.randn(N, 3, 32, 32, device=dev)
ref = None
for name in algos:
    P, E = torch.zeros(N, 10, device=dev), torch.zeros(N, device=dev)
    ws = _C.bma_preresnet_forward(bank[:1], buf[:1], 1, x[:512], 20, 10, P[:512], E[:512], algo=ALGO[name])
    ws = None
    P.zero_(), E.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ws = _C.bma_preresnet_forward(bank, buf, S, x, 20, 10, P, E, algo=ALGO[name], workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if ref is None:
        ref = P.clone()
    print("%-8s S=%d N=%d: %.1f ms  %.0f img*samples/s  %.1f TFLOP/s fp32-equivalent  max|dP/S| vs first = %.2e"
          % (name, S, N, ms, S * N / ms * 1e3, 81.63e6 * S * N / ms / 1e9, ((P - ref).abs().max() / S).item()), flush=True)
