"""Where does the end-to-end Prediction call spend its non-overlapped time?  (development aid)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ursabench_b200 import models, tasks  # noqa: E402

dev = torch.device("cuda", 0)
S = int(os.environ.get("S", "100"))
host_models = bench._samples(lambda: models.PreResNet(num_classes=10, depth=20), S)
x_host, y_host = bench._host_data(10_000)
loaders = {"in_distribution_test": bench._loader(x_host, y_host)}


def step(verbose):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t = tasks.Prediction(loaders, 10, dev, bench.METRICS)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t.update_statistics(host_models, output_performance=False)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    t.get_performance_metrics()
    t5 = time.perf_counter()
    if verbose:
        print("ctor host %.2f ms | upload wait %.2f | update_statistics host %.2f | device tail %.2f | metrics %.2f | total %.2f"
              % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t5 - t0) * 1e3), flush=True)


for i in range(5):
    step(i >= 2)
