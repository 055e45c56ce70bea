"""K3 MLP probe: BMA forward of S x MLP 784-400-400-10 on N images with each engine; CUDA-event medians and error vs fp64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C, models  # noqa: E402

S, N = int(os.environ.get("S", "100")), int(os.environ.get("N", "10000"))
torch.manual_seed(0)
m = models.MLP(400, 784, 10)
D = sum(q.numel() for q in m.parameters())
bank = torch.randn(S, D, device="cuda") * 0.05
x = torch.randn(N, 784, device="cuda")
P, E = torch.zeros(N, 10, device="cuda"), torch.zeros(N, device="cuda")
res = {}
for name, algo in (("tcgen05", _C.ALGO_TCGEN05), ("f16", _C.ALGO_TCGEN05_F16)):
    ws = None
    for _ in range(2):
        P.zero_(); E.zero_()
        ws = _C.bma_mlp_forward(bank, S, x, 784, 400, 10, P, E, workspace=ws, algo=algo)
    torch.cuda.synchronize()
    ts = []
    for _ in range(int(os.environ.get("REPS", "10"))):
        P.zero_(); E.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ws = _C.bma_mlp_forward(bank, S, x, 784, 400, 10, P, E, workspace=ws, algo=algo)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res[name] = P.clone()
    if ts:
        t = ts[len(ts) // 2]
        print("%-8s %.3f ms  %.1f TFLOP/s fp32-equivalent  finite=%s" % (name, t, 955_200 * N * S / t / 1e9, bool(torch.isfinite(P).all())), flush=True)
# fp64 reference on a slice
ns, nn = min(S, 4), min(N, 512)
W = bank[:ns].double()
o = 0
def take(r, c):
    global o
    w = W[:, o:o + r * c].view(ns, r, c); o += r * c
    b = W[:, o:o + r]; o += r
    return w, b
w1, b1 = take(400, 784); w2, b2 = take(400, 400); w3, b3 = take(10, 400)
xd = x[:nn].double()
h = torch.relu(torch.einsum("nk,shk->snh", xd, w1) + b1[:, None, :])
h = torch.relu(torch.einsum("snk,shk->snh", h, w2) + b2[:, None, :])
lg = torch.einsum("snk,sck->snc", h, w3) + b3[:, None, :]
pref = torch.softmax(lg, -1).sum(0)
for name in res:
    Pn, En = torch.zeros(nn, 10, device="cuda"), torch.zeros(nn, device="cuda")
    _C.bma_mlp_forward(bank[:ns].contiguous(), ns, x[:nn].contiguous(), 784, 400, 10, Pn, En, algo=getattr(_C, "ALGO_TCGEN05" if name == "tcgen05" else "ALGO_TCGEN05_F16"))
    print("%-8s max |p - p_fp64| per sample = %.2e  (|logit| max %.1f)" % (name, (Pn.double() - pref).abs().max().item() / ns, lg.abs().max().item()))
