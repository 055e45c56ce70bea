"""Where do the PreResNet BMA engines sit relative to fp32's own noise?  max |p - p_fp64| per engine next to PyTorch fp32
(cuDNN, TF32 off) on the same inputs.  python tools/accuracy_preresnet.py [depth N]"""
import copy
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402
from ursabench_b200.models import PreResNet  # noqa: E402


def main():
    depth, N = (int(v) for v in (sys.argv[1:] + [20, 256][len(sys.argv) - 1:]))
    torch.manual_seed(0)
    m = PreResNet(num_classes=10, depth=depth).cuda().eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
    m.fc.weight.data.mul_(4.0)
    bank = torch.cat([p.detach().reshape(-1) for p in m.parameters()])[None].contiguous()
    bufs = torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32])[None].contiguous()
    x = torch.randn(N, 3, 32, 32, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        l32 = m(x)
        l64 = copy.deepcopy(m).double()(x.double())
    p64 = torch.softmax(l64, -1)
    out = {"depth": depth, "N": N, "logit_absmax": l64.abs().max().item(),
           "torch_fp32": (torch.softmax(l32, -1).double() - p64).abs().max().item()}
    for name, algo in (("ffma", _C.ALGO_FFMA), ("tcgen05", _C.ALGO_TCGEN05), ("fused", _C.ALGO_TCGEN05_FUSED),
                       ("fused16", _C.ALGO_TCGEN05_FUSED_F16)):
        P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
        _C.bma_preresnet_forward(bank, bufs, 1, x, depth, 10, P, E, algo=algo)
        out[name] = (P.double() - p64).abs().max().item()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
