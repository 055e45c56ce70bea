"""MLP 784-400-400-10 BMA engines against an fp64 forward (max |p - p_fp64|), next to PyTorch fp32 (TF32 off)."""
import copy
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402
from ursabench_b200.models import MLP  # noqa: E402


def main():
    gain = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    torch.manual_seed(0)
    N = 2048
    m = MLP(400, 784, 10).cuda()
    for p in m.parameters():
        p.data.mul_(gain)
    bank = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    pad = (-bank.numel()) % 4
    bank = torch.nn.functional.pad(bank, (0, pad))[None].contiguous()
    x = torch.randn(N, 784, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        l32 = m(x)
        l64 = copy.deepcopy(m).double()(x.double())
    p64 = torch.softmax(l64, -1)
    out = {"gain": gain, "logit_absmax": l64.abs().max().item(), "torch_fp32": (torch.softmax(l32, -1).double() - p64).abs().max().item()}
    for name, algo in (("ffma", _C.ALGO_FFMA), ("tcgen05", _C.ALGO_TCGEN05)):
        P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
        _C.bma_mlp_forward(bank, 1, x, 784, 400, 10, P, E, algo=algo)
        out[name] = (P.double() - p64).abs().max().item()
    # speed: the bench's BMA configuration (S = 100, N = 10 000)
    S, Nb = 100, 10_000
    bk = (bank[:, :bank.shape[1]] + 0.01 * torch.randn(S, bank.shape[1], device="cuda")).contiguous()
    xb = torch.randn(Nb, 784, device="cuda")
    P, E = torch.zeros(Nb, 10, device="cuda"), torch.zeros(Nb, device="cuda")
    ws = _C.bma_mlp_forward(bk, S, xb, 784, 400, 10, P, E, algo=_C.ALGO_TCGEN05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _C.bma_mlp_forward(bk, S, xb, 784, 400, 10, P, E, algo=_C.ALGO_TCGEN05, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out.update({"S100_N10k_ms": ms, "TFLOPs": 955_200 * S * Nb / ms / 1e9})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
