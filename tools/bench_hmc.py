"""Time one HMC iteration (BASELINE.json configs[3] shape: MLP 784-200-200-10, 1000 points, 128 chains, L = 10) per gradient engine.
Usage: python tools/bench_hmc.py [chains] [engines...]   (development aid; bench.py `extras.hmc_mlp200_N1000` is the recorded figure)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursabench_b200 import inference, models  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
engines = sys.argv[2:] or ["mlp_tcgen05_fused_f16", "mlp_tcgen05_fused", "mlp_gemm"]
dev = torch.device("cuda")
g = torch.Generator().manual_seed(5)
xs, ys = torch.randn(1000, 1, 28, 28, generator=g), torch.randint(0, 10, (1000,), generator=g)
loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xs, ys), batch_size=1000)
for eng in engines:
    for graph in (True, False):
        hyp = {"step_size": 2.09e-4, "num_samples": 8, "L": 10, "tau": 100.0, "burn": 0, "mass": 0.192, "num_chains": C}
        torch.manual_seed(0)
        hm = inference.HMC(hyp, models.MLP(200, 784, 10), loader, device=dev)
        hm.grad_engine_request = eng
        hm.use_cuda_graph = graph
        hm.sample()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hm.sample()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        flop = 6 * (784 * 200 + 200 * 200 + 200 * 10) * 1000 * C * 11
        print("%-18s graph=%d: %.2f ms / iteration whole call, %.2f ms steady state = %.1f TFLOP/s of gradient work, accept %.3f"
              % (hm.grad_engine, graph, ms, hm.ms_per_iteration, flop / hm.ms_per_iteration / 1e9,
                 float(hm.acceptance_rate.mean())), flush=True)
