"""One WRN-28-10 BMA forward (1 sample x 512 images) on the FP16-split engine and one chain-batched HMC gradient on the FP16-split
GEMM kernel, for `ncu --set full` (profiles/r2s30_*)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402
from ursabench_b200.models import WideResNet  # noqa: E402

dev = torch.device("cuda")
m = WideResNet(num_classes=100, depth=28, widen_factor=10)
bank = torch.cat([p.detach().reshape(-1) for p in m.parameters()])[None].to(dev).contiguous()
nb = sum(b.numel() for b in m.buffers() if b.dtype == torch.float32)
bufs = torch.zeros(1, (nb + 3) // 4 * 4, device=dev)
off = 0
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        c = mod.num_features
        bufs[:, off + c:off + 2 * c] = 1.0
        off += 2 * c
x = torch.randn(512, 3, 32, 32, device=dev)
P, E = torch.zeros(512, 100, device=dev), torch.zeros(512, device=dev)
ws = None
for _ in range(2):
    ws = _C.bma_wrn_forward(bank, bufs, 1, x, 28, 10, 100, P, E, workspace=ws, algo=_C.ALGO_TCGEN05_F16)
torch.cuda.synchronize()
Dh = 199_210
ldh = (Dh + 3) // 4 * 4
th, gh = (torch.randn(128, ldh, device=dev) * 0.05 for _ in range(2))
xh = torch.randn(1000, 784, device=dev)
yh = torch.randint(0, 10, (1000,), device=dev)
ce = torch.zeros(128, device=dev)
ws = None
for _ in range(2):
    ws = _C.hmc_mlp_grad(th, xh, yh, 784, 200, 10, gh, ce, workspace=ws, engine="f16")
torch.cuda.synchronize()
print("done")
