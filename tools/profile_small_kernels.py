"""One launch of each HBM-bound / GEMM kernel at its bench size, for `ncu --set full` (profiles/r2s14_*, r2s23_*):
K1 sgmcmc_step (WRN size), K2a swag_collect, K2b swag_draw (S = 30, K = 20 and the diagonal K = 0), K2c ring_gram_tc, K5 hmc_leapfrog,
the MLP GEMMs of both tensor-core engines (BMA forward S = 16 x N = 10 000) and the chain-batched HMC gradient GEMMs."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402

dev = torch.device("cuda")
D = 36_546_980
ld = (D + 3) // 4 * 4
p, g, v = (torch.randn(ld, device=dev) * 0.05 for _ in range(3))
for rep in range(2):
    _C.sgmcmc_step(p, g, v, lr=0.01, momentum=0.5, wd_over_n=1e-4, noise_mul=math.sqrt(0.01), noise_div=5e4, seed=3, step=rep)
mean, sq = torch.zeros(ld, device=dev), torch.zeros(ld, device=dev)
K, S = 20, 30
ring = torch.randn(K, ld, device=dev) * 0.01
for rep in range(2):
    _C.swag_collect(p, mean, sq, ring[rep], rep)
var = torch.empty(ld, device=dev)
_C.swag_variance(mean, mean * mean + 1e-4, var)
bank = torch.empty(S, ld, device=dev)
z2 = torch.randn(S, K, device=dev)
for rep in range(2):
    _C.swag_draw(bank, mean, var, D, ring=ring, z2=z2, rank_div=math.sqrt(K - 1.0), seed=5, step=rep)
    _C.swag_draw(bank, mean, var, D, seed=5, step=rep)
    _C.swag_gram(ring, D)
del bank, ring, p, g, v, mean, sq, var
Dh = 199_210
ldh = (Dh + 3) // 4 * 4
th, rh, gh = (torch.randn(128, ldh, device=dev) * 0.05 for _ in range(3))
for rep in range(2):
    _C.hmc_leapfrog(th, rh, gh, kick=1e-4, drift=5e-4, tau=100.0)
x = torch.randn(1000, 784, device=dev)
y = torch.randint(0, 10, (1000,), device=dev)
ce = torch.zeros(128, device=dev)
ws = None
for rep in range(2):
    ws = _C.hmc_mlp_grad(th, x, y, 784, 200, 10, gh, ce, workspace=ws)
Dm = 400 * 784 + 400 + 400 * 400 + 400 + 10 * 400 + 10
bankm = torch.randn(16, Dm, device=dev) * 0.05
xm = torch.randn(10_000, 784, device=dev)
P, E = torch.zeros(10_000, 10, device=dev), torch.zeros(10_000, device=dev)
ws = None
for rep in range(2):
    ws = _C.bma_mlp_forward(bankm, 16, xm, 784, 400, 10, P, E, algo=_C.ALGO_TCGEN05, workspace=ws)
ws = None
for rep in range(2):
    ws = _C.bma_mlp_forward(bankm, 16, xm, 784, 400, 10, P, E, algo=_C.ALGO_TCGEN05_F16, workspace=ws)
torch.cuda.synchronize()
print("done")
