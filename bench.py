"""bench.py -- contract benchmark of the URSABench hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-graph] [--skip-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): cyclical SGHMC on PreResNet-20 (D = 272 282), synthetic CIFAR-10-shaped
batches of 128, N_train = 50 000, ONE independent chain per GPU (no data-path collective; "scaling": "weak").
A step = forward + backward (PyTorch autograd, as the north star keeps them) + ONE fused K1 launch that applies the
cSGHMC update with in-register Philox noise and zeroes the gradients.

  value     whole-job steps/s (sum over chains), batches resident in HBM (a pool larger than L2), device-timed
  e2e       same metric through the public class API (`cSGHMC.train_step`) with pinned HOST batches: H2D of every
            batch and a D2H read of every step's loss inside the timed region
  roofline  the K1 kernel on the HBM-bound layout the north star names (chain-batched [128, D], 697 MB/launch,
            20 B/param) timed live with CUDA events; `in_step` is the same kernel at the single-chain size
  cpu_baseline / --impl reference   the reference's CPU path (oracle/port_torch.py, same op sequence as
            optimSGHMC.step + the sampler loop) on the box's host cores
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "cSGHMC PreResNet-20 (D=272282) CIFAR-10-shaped batch 128, 1 chain/GPU (BASELINE.json configs[1])"
HYP = {"lr_0": 0.5, "prior_std": 0.5, "num_samples_per_cycle": 3, "cycle_length": 50, "burn_in_epochs": 0,
       "num_cycles": 17, "alpha": 0.5}           # hyperparams/WideResNet28x10CIFAR10/csghmc_hyperparams.json shape
N_TRAIN, BATCH, NUM_CLASSES = 50_000, 128, 10
POOL_BATCHES = 96                                 # 96 x 1.57 MB = 151 MB of distinct inputs (> 126 MB L2)
K1_CHAINS = 128                                   # roofline layout: [128 chains, D] = 697 MB per launch at 20 B/param


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class _ClockSampler:
    """nvidia-smi sampled DURING the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, smax, power = [], set(), None, []
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def _make_model():
    from ursabench_b200 import models
    torch.manual_seed(0)
    return models.PreResNet(num_classes=NUM_CLASSES, depth=20)


class _ListLoader:
    """Minimal train_loader duck type: pre-batched tensors + the attributes the inference classes read."""

    class _DS:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

    def __init__(self, batches, n_train, batch_size):
        self.batches, self.dataset, self.batch_size = batches, self._DS(n_train), batch_size

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_steps_per_s(budget_s, max_steps, warm=1):
    """Reference CPU path: PreResNet-20 fwd/bwd + optimSGHMC.step op sequence (oracle/port_torch.py)."""
    from oracle import port_torch as PT
    from oracle import restate as R
    torch.set_num_threads(os.cpu_count() or 1)
    model = _make_model()
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(BATCH, 3, 32, 32, generator=g), torch.randint(0, NUM_CLASSES, (BATCH,), generator=g))
               for _ in range(4)]
    opt = PT.PortOptimSGHMC(model.parameters(), HYP["lr_0"], 1 - HYP["alpha"], 1 / HYP["prior_std"] ** 2, N_TRAIN)
    nb = R.csghmc_num_batch(N_TRAIN, BATCH)
    crit = torch.nn.CrossEntropyLoss()
    model.train()

    def one(i):
        x, y = batches[i % len(batches)]
        opt.lr = PT.port_csghmc_lr(HYP["lr_0"], 0, i, nb, HYP["cycle_length"], HYP["num_cycles"])
        logits = model(x)
        opt.zero_grad()
        loss = crit(logits, y)
        loss.backward()
        loss.item()
        opt.step(add_langevin_noise=True)

    for i in range(warm):
        one(i)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n < 2 or time.perf_counter() - t0 < budget_s):
        one(warm + n)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, n, dt, cores = cpu_reference_steps_per_s(budget_s=60.0, max_steps=max(args.steps, 2), warm=min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": "sgmcmc_steps_per_s", "value": sps, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": n, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 / sps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path (torch-CPU port of the reference's op sequence), "
                   "one chain on the host cores; /root/reference does not exist on the GPU box"},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": "%d cSGHMC steps (fwd+bwd+optimSGHMC.step), batch 128, %.1f s" % (n, dt)},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def _event_time_ms(fn, iters, stream_sync=True):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return sum(ts) / len(ts), ts


def run_ours(args):
    from ursabench_b200 import _C, dist as udist, inference
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU arm)")
    rank, world = udist.init_from_env("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _C.lib()
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    K, W = args.steps, max(args.warmup, 3)

    # ---- the chain of this rank --------------------------------------------------------------------------------
    model = _make_model().to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    host_x = torch.randn(POOL_BATCHES, BATCH, 3, 32, 32, generator=g).pin_memory()
    host_y = torch.randint(0, NUM_CLASSES, (POOL_BATCHES, BATCH), generator=g).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)
    loader = _ListLoader([(host_x[i], host_y[i]) for i in range(POOL_BATCHES)], N_TRAIN, BATCH)
    torch.manual_seed(1234 + rank)                         # Philox key of this chain
    inf = inference.cSGHMC(dict(HYP), model, loader, device=dev)
    inf._channels_last = bool(args.channels_last)          # default on: NHWC activations for cuDNN
    inf.optimizer.elem_offset = udist.chain_elem_offset(rank, inf.flat.D)
    inf.model.train()
    if not args.no_graph:
        inf.enable_cuda_graph(dev_x[0], dev_y[0])
    step_no = [0]

    def lr_and_gate():
        i = step_no[0]
        step_no[0] += 1
        inf._adjust_learning_rate(inf.optimizer, 0, i % 391)
        return True

    def resident_step():
        i = step_no[0] % POOL_BATCHES
        noise = lr_and_gate()
        return inf.train_step(dev_x[i], dev_y[i], noise)

    def host_step():
        i = step_no[0] % POOL_BATCHES
        noise = lr_and_gate()
        loss = inf.train_step(host_x[i], host_y[i], noise)
        return loss.item()                                 # D2H of the step's result, every step (sghmc.py:82)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: resident inputs ---------------------------------------------------------------------------------
    for _ in range(W):
        resident_step()
    launches0 = inf.optimizer.launches
    replays0 = inf._graph.replays if getattr(inf, "_graph", None) is not None else 0
    sampler = _ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        resident_step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    t_ms = udist.allreduce_max_scalar(e0.elapsed_time(e1), dev)
    graph_replays = (inf._graph.replays - replays0) if getattr(inf, "_graph", None) is not None else 0
    my_launches = (inf.optimizer.launches - launches0) + graph_replays   # set_dyn + K1-in-graph, or K1 eager
    value = world * K / (t_ms / 1e3)

    # ---- e2e: host batches through the public API, loss read back every step -------------------------------------
    for _ in range(W):
        host_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        host_step()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = udist.allreduce_max_scalar(max(e0.elapsed_time(e1), wall_ms), dev)
    e2e_value = world * K / (e2e_ms / 1e3)
    h2d = BATCH * 3 * 32 * 32 * 4 + BATCH * 8

    # ---- roofline: K1 on the HBM-bound layout, live CUDA events ---------------------------------------------------
    peak, peak_src = _peaks()
    D = inf.flat.D
    ld = (D + 3) // 4 * 4
    n_big = K1_CHAINS * ld
    pb, gb, vb = (torch.randn(n_big, device=dev) for _ in range(3))
    mom = 1 - HYP["alpha"]
    kw = dict(lr=0.1, momentum=mom, wd_over_n=(1 / HYP["prior_std"] ** 2) / N_TRAIN,
              noise_mul=math.sqrt(2 * (1 - mom) * 0.1), noise_div=float(N_TRAIN), seed=7)
    cnt = [0]

    def k1_big():
        cnt[0] += 1
        _C.sgmcmc_step(pb, gb, vb, step=cnt[0], **kw)
    for _ in range(5):
        k1_big()
    torch.cuda.synchronize()
    k1_ms, _ = _event_time_ms(k1_big, max(20, min(K, 100)))
    alg_bytes = 20 * n_big
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    # the same kernel at the single-chain size inside the step (L2 resident, launch bound): CUDA-graph replay of
    # 20 back-to-back launches so that host launch latency is excluded
    p1, g1, v1 = (torch.randn(ld, device=dev) for _ in range(3))
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        _C.sgmcmc_step(p1, g1, v1, step=1, **kw)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        for j in range(20):
            _C.sgmcmc_step(p1, g1, v1, step=2 + j, **kw)
    gr.replay()
    torch.cuda.synchronize()
    small_ms, _ = _event_time_ms(gr.replay, 20)
    small_us = small_ms * 1e3 / 20
    del pb, gb, vb
    roofline = {"bound": "hbm", "kernel": "sgmcmc_step_kernel<momentum, philox>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "layout": "[%d chains, D=%d] flat fp32, %d MB algorithmic per launch (20 B/param)" % (K1_CHAINS, D, alg_bytes >> 20),
                "ms_per_launch": k1_ms,
                "in_step": {"D": D, "bytes": 20 * D, "us_per_launch_graph_replay": small_us,
                            "achieved_GBps": 20 * D / (small_us * 1e-6) / 1e9,
                            "note": "single chain: 5.4 MB, L2 resident, launch-latency bound (SURVEY 7.3)"}}

    line = {
        "metric": "sgmcmc_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "chains": world, "parallelism": "chains x%d, no collective" % world,
                   "cuda_graph": not args.no_graph, "channels_last": bool(args.channels_last), "cudnn_benchmark": bool(args.cudnn_benchmark), "l2": "inputs cycle through a %d-batch pool (%d MB > 126 MB L2)" % (POOL_BATCHES, POOL_BATCHES * h2d >> 20),
                   "hyper": HYP},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / K, "api": "ursabench_b200.inference.cSGHMC.train_step(host pinned batch) + loss.item()"},
        "gpu_launches": my_launches,
        "roofline": roofline,
    }

    if not args.skip_extras:
        try:
            line["extras"] = run_extras(dev, rank, world, peak)
        except Exception as e:  # noqa: BLE001  (extras never invalidate the headline line)
            line["extras"] = {"error": repr(e)}
    if rank == 0 and world == 1:
        sps, n, dt, cores = cpu_reference_steps_per_s(budget_s=15.0, max_steps=60)
        line["cpu_baseline"] = {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                                "sample": "%d cSGHMC steps (PreResNet-20 fwd+bwd+optimSGHMC.step op sequence), batch 128, %.1f s"
                                          % (n, dt)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_extras(dev, rank, world, peak):
    """Secondary figures of the same path (not the contract metric): K1 at WRN size, SWAG collect / draw,
    BMA img/s with samples sharded over ranks + the single all-reduce."""
    from ursabench_b200 import _C, dist as udist, models
    out = {}
    D = 36_546_980
    ld = (D + 3) // 4 * 4
    p, g, v = (torch.randn(ld, device=dev) * 0.05 for _ in range(3))
    cnt = [0]

    def k1(mom):
        cnt[0] += 1
        _C.sgmcmc_step(p, g, v if mom else None, lr=0.01, momentum=mom, wd_over_n=1e-4,
                       noise_mul=math.sqrt(2 * (1 - mom) * 0.01), noise_div=5e4, seed=3, step=cnt[0])
    for mom, name, bpp in ((0.5, "k1_sghmc_wrn28x10", 20), (0.0, "k1_sgld_wrn28x10", 12)):
        for _ in range(3):
            k1(mom)
        ms, _ = _event_time_ms(lambda: k1(mom), 20)
        out[name] = {"D": D, "ms": ms, "steps_per_s": 1e3 / ms, "GBps": bpp * D / ms / 1e6, "frac": bpp * D / ms / 1e6 / peak}
    mean, sq = torch.zeros(ld, device=dev), torch.zeros(ld, device=dev)
    K, S = 20, 30
    ring = torch.randn(K, ld, device=dev) * 0.01

    def collect():
        cnt[0] += 1
        _C.swag_collect(p, mean, sq, ring[cnt[0] % K], cnt[0] % 50)
    for _ in range(3):
        collect()
    ms, _ = _event_time_ms(collect, 20)
    out["k2_collect_wrn28x10"] = {"ms": ms, "GBps": 24 * D / ms / 1e6, "frac": 24 * D / ms / 1e6 / peak}
    var = torch.empty(ld, device=dev)
    _C.swag_variance(mean, mean * mean + 1e-4, var)
    bank = torch.empty(S, ld, device=dev)
    z2 = torch.randn(S, K, device=dev)
    fn = lambda: _C.swag_draw(bank, mean, var, D, ring=ring, z2=z2, rank_div=math.sqrt(K - 1.0), seed=5, step=1)  # noqa: E731
    for _ in range(3):
        fn()
    ms, _ = _event_time_ms(fn, 10)
    out["k2_draw_S30_K20_wrn28x10"] = {"ms": ms, "GBps": (K + 2 + S) * 4 * D / ms / 1e6,
                                       "frac": (K + 2 + S) * 4 * D / ms / 1e6 / peak, "us_per_draw": ms * 1e3 / S}
    fn = lambda: _C.swag_gram(ring, D)  # noqa: E731
    for _ in range(3):
        fn()
    ms, _ = _event_time_ms(fn, 10)
    out["k2c_gram_K20_wrn28x10"] = {"ms": ms, "GBps": K * 4 * D / ms / 1e6, "frac": K * 4 * D / ms / 1e6 / peak}
    del p, g, v, mean, sq, ring, var, bank
    torch.cuda.empty_cache()
    # BMA: MLP 784-400-400-10, S = 100 posterior samples sharded over ranks, N = 10 000
    S_all, N = 100, 10_000
    lo, hi = udist.shard_range(S_all, rank, world)
    m = models.MLP(400, 784, 10)
    Dm = sum(q.numel() for q in m.parameters())
    bankm = torch.randn(hi - lo, Dm, device=dev) * 0.05
    x = torch.randn(N, 784, device=dev)
    y = torch.randint(0, 10, (N,), device=dev)
    P, E = torch.zeros(N, 10, device=dev), torch.zeros(N, device=dev)
    ws = [None]

    for algo_name, algo in (("ffma", _C.ALGO_FFMA), ("tcgen05", _C.ALGO_TCGEN05)):
        def bma():
            P.zero_()
            E.zero_()
            ws[0] = _C.bma_mlp_forward(bankm, hi - lo, x, 784, 400, 10, P, E, workspace=ws[0], algo=algo)
            Pr, Er, n = udist.allreduce_bma(P, E, hi - lo)
            return _C.bma_metrics(Pr, n, y)
        ws[0] = None
        bma()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms, _ = _event_time_ms(bma, 5)
        ms = udist.allreduce_max_scalar(ms, dev)
        out["bma_mlp400_S100_N10k_" + algo_name] = {"ms": ms, "img_per_s_over_S_samples": N / ms * 1e3,
                                                    "img_samples_per_s": N * S_all / ms * 1e3,
                                                    "TFLOPs": 955_200 * N * S_all / ms / 1e9, "n_gpus": world}
    del bankm, x, P, E
    ws[0] = None
    torch.cuda.empty_cache()
    # BMA: PreResNet-20 (BASELINE.json configs[4]), S = 100 samples sharded over ranks, N = 10 000 test images
    m = models.PreResNet(num_classes=10, depth=20)
    Dp = sum(q.numel() for q in m.parameters())
    nbuf = sum(b.numel() for b in m.buffers() if b.dtype == torch.float32)
    ns = hi - lo
    flat = torch.cat([q.detach().reshape(-1) for q in m.parameters()]).to(dev)
    bankp = (flat[None, :] + 0.01 * torch.randn(ns, Dp, device=dev)).contiguous()
    bufp = torch.zeros(ns, (nbuf + 3) // 4 * 4, device=dev)
    # running_mean = 0, running_var = 1 per BN layer: buffers are [mean(C), var(C)] per layer in order
    off = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            bufp[:, off:off + c] = 0.0
            bufp[:, off + c:off + 2 * c] = 1.0
            off += 2 * c
    xi = torch.randn(N, 3, 32, 32, device=dev)
    P, E = torch.zeros(N, 10, device=dev), torch.zeros(N, device=dev)

    for algo_name, algo in (("ffma", _C.ALGO_FFMA), ("tcgen05", _C.ALGO_TCGEN05), ("fused", _C.ALGO_TCGEN05_FUSED),
                            ("fused16", _C.ALGO_TCGEN05_FUSED_F16)):
        def bma_conv():
            P.zero_()
            E.zero_()
            ws[0] = _C.bma_preresnet_forward(bankp, bufp, ns, xi, 20, 10, P, E, workspace=ws[0], algo=algo)
            Pr, Er, n = udist.allreduce_bma(P, E, ns)
            return _C.bma_metrics(Pr, n, y)
        ws[0] = None
        ws[0] = _C.bma_preresnet_forward(bankp[:1], bufp[:1], 1, xi[:512], 20, 10, P[:512], E[:512], workspace=None,
                                         algo=algo)
        ws[0] = None
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms, _ = _event_time_ms(bma_conv, 2)
        ms = udist.allreduce_max_scalar(ms, dev)
        out["bma_preresnet20_S100_N10k_" + algo_name] = {"ms": ms, "img_per_s_over_S_samples": N / ms * 1e3,
                                                         "img_samples_per_s": N * S_all / ms * 1e3,
                                                         "TFLOPs": 81.63e6 * N * S_all / ms / 1e9, "n_gpus": world}
    # BASELINE.json configs[4]: the S = 10 / 100 / 1000 sweep on the product engine (S = 100 is the line above)
    for S_sw in (10, 1000):
        lo_s, hi_s = udist.shard_range(S_sw, rank, world)
        ns_s = hi_s - lo_s
        bank_s = (flat[None, :] + 0.01 * torch.randn(max(ns_s, 1), Dp, device=dev)).contiguous()
        buf_s = bufp[:1].expand(max(ns_s, 1), -1).contiguous()

        def bma_sweep():
            P.zero_()
            E.zero_()
            if ns_s > 0:
                ws[0] = _C.bma_preresnet_forward(bank_s, buf_s, ns_s, xi, 20, 10, P, E, workspace=ws[0],
                                                 algo=_C.ALGO_TCGEN05_FUSED_F16)
            Pr, Er, n = udist.allreduce_bma(P, E, ns_s)
            return _C.bma_metrics(Pr, n, y)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms, _ = _event_time_ms(bma_sweep, 1)
        ms = udist.allreduce_max_scalar(ms, dev)
        out["bma_preresnet20_S%d_N10k_fused16" % S_sw] = {"ms": ms, "img_per_s_over_S_samples": N / ms * 1e3,
                                                          "img_samples_per_s": N * S_sw / ms * 1e3,
                                                          "TFLOPs": 81.63e6 * N * S_sw / ms / 1e9, "n_gpus": world}
        del bank_s, buf_s
    del bankp, bufp, xi, P, E
    ws[0] = None
    torch.cuda.empty_cache()
    # BMA: WideResNet-28-10, C = 100 (BASELINE.json configs[2]): S = 30 SWAG draws sharded over ranks, N = 10 000 test images.
    # Every conv runs on the persistent 3xTF32 tcgen05 implicit-GEMM kernel (csrc/bma_wrn_tc.cu); beside it the engine it
    # replaces -- the per-sample PyTorch forward with cuDNN fp32 (TF32 off, as parity demands) -- on a bounded sample.
    S_w, C_w = 30, 100
    lo_w, hi_w = udist.shard_range(S_w, rank, world)
    ns = hi_w - lo_w
    m = models.WideResNet(num_classes=C_w, depth=28, widen_factor=10)
    Dw = sum(q.numel() for q in m.parameters())
    nbuf = sum(b.numel() for b in m.buffers() if b.dtype == torch.float32)
    flat = torch.cat([q.detach().reshape(-1) for q in m.parameters()]).to(dev)
    bankw = torch.empty(max(ns, 1), Dw, device=dev)
    for i in range(max(ns, 1)):                               # row by row: no [S, D] temporary next to the 4.4 GB bank
        bankw[i] = flat + 0.002 * torch.randn(Dw, device=dev)
    bufw = torch.zeros(max(ns, 1), (nbuf + 3) // 4 * 4, device=dev)
    off = 0
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            bufw[:, off + c:off + 2 * c] = 1.0
            off += 2 * c
    xi = torch.randn(N, 3, 32, 32, device=dev)
    yw = torch.randint(0, C_w, (N,), device=dev)
    P, E = torch.zeros(N, C_w, device=dev), torch.zeros(N, device=dev)
    wrn_flop = 11_902_350_336                                 # 2*MAC per image and sample (convs + shortcuts + linear)

    def bma_wrn():
        P.zero_()
        E.zero_()
        if ns > 0:
            ws[0] = _C.bma_wrn_forward(bankw, bufw, ns, xi, 28, 10, C_w, P, E, workspace=ws[0])
        Pr, Er, n = udist.allreduce_bma(P, E, ns)
        return _C.bma_metrics(Pr, n, yw)
    ws[0] = _C.bma_wrn_forward(bankw[:1], bufw[:1], 1, xi[:512], 28, 10, C_w, P[:512], E[:512], workspace=None)   # warm-up
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    wsampler = _ClockSampler(dev.index if dev.index is not None else 0)
    wsampler.start()
    ms, _ = _event_time_ms(bma_wrn, 1)
    wclocks = wsampler.stop()                               # a 5-17 s tensor-core run: the SM clock under its own power draw
    ms = udist.allreduce_max_scalar(ms, dev)
    out["bma_wrn28x10_S30_N10k_tcgen05"] = {"ms": ms, "img_per_s_over_S_samples": N / ms * 1e3,
                                            "img_samples_per_s": N * S_w / ms * 1e3,
                                            "TFLOPs": wrn_flop * N * S_w / ms / 1e9, "n_gpus": world, "clocks": wclocks}
    # K3b: BatchNorm re-estimation of one SWAG draw (util.bn_update, once per sample): bounded sample of the 50 000-image pass
    Nbn, Bbn = 2048, 128
    bn_fn = lambda: _C.wrn_bn_update(bankw[0], bufw[0], xi[:Nbn], Bbn, 28, 10, C_w, workspace=ws[0])  # noqa: E731
    ws[0] = None
    ws[0] = bn_fn()
    torch.cuda.synchronize()
    ms, _ = _event_time_ms(bn_fn, 2)
    out["k3b_bn_update_wrn28x10_N2048_b128"] = {"ms": ms, "img_per_s": Nbn / ms * 1e3, "TFLOPs": wrn_flop * Nbn / ms / 1e9,
                                                "full_pass_50k_images_s": ms * 50_000 / Nbn / 1e3}
    if rank == 0:
        worker = m.to(dev).eval()
        mm_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
                fwd = lambda: [worker(xi[i:i + 128]) for i in range(0, 512, 128)]  # noqa: E731
                fwd()
                tms, _ = _event_time_ms(fwd, 2)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = mm_tf32
        out["bma_wrn28x10_torch_cudnn_fp32"] = {"sample": "S=1, N=512, batch 128", "ms": tms, "img_samples_per_s": 512 / tms * 1e3,
                                                "TFLOPs": wrn_flop * 512 / tms / 1e9}
        del worker
    del bankw, bufw, xi, P, E
    ws[0] = None
    torch.cuda.empty_cache()
    # HMC (BASELINE.json configs[3]): MLP 784-200-10, full batch of 1000 points, 128 chains per GPU, L = 10
    from ursabench_b200 import inference
    Ch, Lh = 128, 10
    Dh = 199_210
    ldh = (Dh + 3) // 4 * 4
    th, rh, gh = (torch.randn(Ch, ldh, device=dev) * 0.05 for _ in range(3))
    fn = lambda: _C.hmc_leapfrog(th, rh, gh, kick=1e-4, drift=5e-4, tau=100.0)  # noqa: E731
    for _ in range(3):
        fn()
    ms, _ = _event_time_ms(fn, 20)
    out["k5_leapfrog_128x199210"] = {"ms": ms, "GBps": 20 * Ch * ldh / ms / 1e6, "frac": 20 * Ch * ldh / ms / 1e6 / peak}
    del th, rh, gh
    g = torch.Generator().manual_seed(5)
    xs, ys = torch.randn(1000, 1, 28, 28, generator=g), torch.randint(0, 10, (1000,), generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xs, ys), batch_size=1000)
    hyp = {"step_size": 2.09e-4, "num_samples": 2, "L": Lh, "tau": 100.0, "burn": 0, "mass": 0.192, "num_chains": Ch}
    hm = inference.HMC(hyp, models.MLP(200, 784, 10), loader, device=dev)
    hm.sample()                                            # warm-up (cuBLAS heuristics, vmap tracing)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hm.sample()
    e1.record()
    torch.cuda.synchronize()
    ms = udist.allreduce_max_scalar(e0.elapsed_time(e1), dev)
    steps = hyp["num_samples"] * (Lh + 1)                  # gradient evaluations per chain
    out["hmc_mlp200_N1000"] = {"chains_per_gpu": Ch, "n_gpus": world, "L": Lh, "ms_per_iteration": ms / hyp["num_samples"],
                               "chain_leapfrog_steps_per_s": world * Ch * hyp["num_samples"] * Lh / ms * 1e3,
                               "grad_TFLOPs": 6 * (784 * 200 + 200 * 200 + 200 * 10) * 1000 * Ch * steps / ms / 1e9,
                               "accept_rate": float(hm.acceptance_rate.mean())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--channels-last", type=int, default=1, help="NHWC activations in fwd/bwd (engine default)")
    ap.add_argument("--cudnn-benchmark", type=int, default=0, help="torch.backends.cudnn.benchmark for fwd/bwd")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
