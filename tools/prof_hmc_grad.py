import sys, torch
sys.path.insert(0, '/root/repo')
from ursabench_b200 import _C
C, N, i, h, k = 128, 1000, 784, 200, 10
D = h*i + h + h*h + h + k*h + k
ld = (D+3)//4*4
theta = torch.randn(C, ld, device='cuda')*0.05
x = torch.randn(N, i, device='cuda'); y = torch.randint(0, k, (N,), device='cuda')
g = torch.zeros(C, ld, device='cuda'); ce = torch.zeros(C, device='cuda')
ws = None
for _ in range(3):
    ws = _C.hmc_mlp_grad(theta, x, y, i, h, k, g, ce, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ws = _C.hmc_mlp_grad(theta, x, y, i, h, k, g, ce, workspace=ws)
e1.record(); torch.cuda.synchronize()
print("ms per gradient", e0.elapsed_time(e1)/10)
