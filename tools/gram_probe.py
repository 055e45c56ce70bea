"""K2c probe: Gram matrix of a [20, 36.5 M] ring; CUDA-event median and the fp64 error."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402

D, K = 36_546_980, 20
ld = (D + 3) // 4 * 4
ring = torch.randn(K, ld, device="cuda") * 0.01
for _ in range(3):
    g = _C.swag_gram(ring, D)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g = _C.swag_gram(ring, D)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
ref = torch.zeros(K, K, dtype=torch.float64, device="cuda")
for c0 in range(0, D, 1 << 22):
    blk = ring[:, c0:min(D, c0 + (1 << 22))].double()
    ref += blk @ blk.t()
print("gram K=%d  %.3f ms (min %.3f)  %.0f GB/s   max rel err %.2e" % (K, ts[len(ts) // 2], ts[0], K * 4 * D / ts[len(ts) // 2] / 1e6,
                                                                    ((g - ref).abs().max() / ref.abs().max()).item()))
