"""K2b probe: the rank-K draw at the bench size (D = 36.5 M) for S / K combinations; prints CUDA-event medians.  With
`ncu -k regex:swag_draw --set full` around it: one profiled launch per combination (tools/ncu_summary.py digests the report)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import _C  # noqa: E402

dev = torch.device("cuda")
D = 36_546_980
ld = (D + 3) // 4 * 4
reps = int(os.environ.get("REPS", "20"))
mean = torch.randn(ld, device=dev) * 0.05
var = torch.rand(ld, device=dev) * 1e-4 + 1e-6
ring = torch.randn(20, ld, device=dev) * 0.01
for S, K in [(30, 20), (30, 0), (24, 20), (16, 20), (100, 20)]:
    bank = torch.empty(S, ld, device=dev)
    z2 = torch.randn(S, max(K, 1), device=dev)

    def fn(step):
        _C.swag_draw(bank, mean, var, D, ring=ring[:K] if K else None, z2=z2 if K else None, rank_div=math.sqrt(19.0), seed=5, step=step)
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    if ts:
        byts = (K + 2 + S) * 4 * D if K else (2 + S) * 4 * D
        print("draw S=%d K=%d  %.3f ms (min %.3f)  %.0f GB/s" % (S, K, ts[len(ts) // 2], ts[0], byts / ts[len(ts) // 2] / 1e6), flush=True)
    del bank
