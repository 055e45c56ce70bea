"""bench.py -- contract benchmark of the URSABench hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--skip-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], the "BMA img/s over S samples at 1/2/4/8 B200" half of the metric): Bayesian model
averaging of S = 100 PreResNet-20 posterior samples over N = 10 000 synthetic CIFAR-10-shaped test images through
``tasks.Prediction`` -- the sample-batched tcgen05 forward (stem, three fused stage kernels, two transitions, head),
the softmax-average / entropy accumulation, ONE NCCL all-reduce of the [N, C] + [N] sums when N_gpus > 1, and the K4
metric counters (accuracy, NLL, Brier, ECE bins).  A step = one complete evaluation; the unit of work is one
(image, sample) pair: value = S * N / t  [img*samples/s], whole job.  Total work is fixed ("scaling": "strong"): the
(sample, image) grid is split into balanced contiguous shares (``dist.shard_pairs``) -- whole samples when S divides
over the ranks, image blocks otherwise.

  value     inputs resident in HBM when the timed region starts (test images + replicated sample bank, 232 MB > L2),
            device-timed with CUDA events, max over ranks
  e2e       the call a reference user makes, from HOST data: ``Prediction(dataloader, ...)`` (upload of the pinned
            123 MB test set), ``update_statistics(list of CPU modules)`` (upload of every sample this rank touches),
            ``get_performance_metrics()`` (all-reduce, K4, counters read back into the metrics dict)
  roofline  the dominant kernel of the step, ``preresnet_stage16_kernel<16>`` (tensor bound): algorithmic FLOPs per
            launch / its average launch duration, measured live with CUDA events on the launching stream inside the
            timed steps (``ursa_profile_begin/end``); per-kernel times of the whole forward beside it
  cpu_baseline / --impl reference   the reference's own ``tasks.Prediction`` (unmodified, staged by oracle/make_ref.py
            into oracle/_ref; the torch-CPU port in oracle/port_torch.py if that is absent) on the box's host cores,
            each step a bounded sample of the same workload (S_ref samples x N_ref images), extrapolated linearly
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_ALL, N_TEST, NUM_CLASSES, DEPTH, BATCH = 100, 10_000, 10, 20, 128
METRICS = ["error_rate", "nll", "brier_score", "ece"]
FLOP_PER_PAIR = 81.63e6                           # 2*MAC over convs + fc of PreResNet-20
# per (image, sample): stage 1 = the 3 -> 16 stem conv (0.885 MFLOP) + 6 convs, stages 2 / 3 = 5 convs each
STAGE_FLOP = {"stage_c16": 0.884736e6 + 6 * 4.718592e6, "stage_c32": 5 * 4.718592e6, "stage_c64": 5 * 4.718592e6}
REF_S, REF_N = 2, N_TEST                          # bounded sample of the reference arm per step: 2 of the 100 samples over the
                                                  # whole test set (20 000 pairs, 1.5-2 s on 16 host cores); S enters linearly
# identical in both arms (the driver compares the dicts)
CONFIG = {
    "workload": "BMA evaluation (tasks.Prediction) of S=100 PreResNet-20 posterior samples on N=10000 synthetic "
                "CIFAR-10-shaped test images (BASELINE.json configs[4])",
    "S": S_ALL, "N": N_TEST, "num_classes": NUM_CLASSES, "batch_size": BATCH, "metrics": METRICS,
    "unit_of_work": "one (image, sample) forward + softmax-average; step = S*N pairs + metrics",
    "parallelism": "(sample, image-block) grid sharded over ranks, one all-reduce of N*C+N+1 floats",
    "l2": "inputs larger than L2 (123 MB test set + 109 MB sample bank vs 126 MB)",
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1350.0}, "fallback (B200_PROFILING.md)"


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



# ------------------------------------------------------------------------------------------------ synthetic workload
def _host_data(n=N_TEST, pin=True):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 3, 32, 32, generator=g)
    y = torch.randint(0, NUM_CLASSES, (n,), generator=g)
    if pin and torch.cuda.is_available():
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


def _loader(x, y):
    return torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BATCH, shuffle=False)


def _fill_sample(module, base_flat, s):
    """Sample s = the seed-0 initialisation + 0.01 * N(0, 1) (seed 1000 + s), BatchNorm statistics at their defaults."""
    g = torch.Generator().manual_seed(1000 + s)
    flat = base_flat + 0.01 * torch.randn(base_flat.numel(), generator=g)
    off = 0
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    return module


def _samples(model_ctor, n_samples):
    """``n_samples`` CPU modules built by ``model_ctor`` (ours or the reference's class: same architecture, same values)."""
    import copy
    torch.manual_seed(0)
    base = model_ctor()
    base_flat = torch.cat([p.detach().reshape(-1) for p in base.parameters()])
    return [_fill_sample(copy.deepcopy(base), base_flat, s) for s in range(n_samples)]


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _reference_modules():
    """(tasks.Prediction, PreResNet ctor, optimSGHMC, kind) from the UNMODIFIED reference when its staged copy (or
    /root/reference) is importable, else from the torch-CPU port."""
    from oracle import stubs
    try:
        if not os.path.isdir(os.path.join(stubs.REF_ROOT, "URSABench")):
            raise ImportError("no reference package under %s" % stubs.REF_ROOT)
        stubs.install()
        from URSABench import tasks as rtasks
        from URSABench.inference.optim_sghmc import optimSGHMC as ref_opt
        from URSABench.models import preresnet as rpre
        return rtasks.Prediction, (lambda: rpre.PreResNet(num_classes=NUM_CLASSES, depth=DEPTH)), ref_opt, "reference"
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("bench.py: unmodified reference not importable (%r); timing the torch-CPU port\n" % (e,))
        return None, None, None, "port"


class _PortPrediction:
    """oracle/port_torch.py behind the reference's Prediction call shape (used only when oracle/_ref is absent)."""

    def __init__(self, dataloader, num_classes, device, metric_list):
        from oracle import port_torch as PT
        self.PT, self.loader, self.C = PT, dataloader["in_distribution_test"], num_classes
        self.targets = torch.cat([yb for _, yb in self.loader])
        self.n = 0
        self.proba = None

    def update_statistics(self, models, output_performance=True, smoothing=True):
        self.proba, _ = self.PT.port_prediction_update(models, list(self.loader), self.C)
        self.n += len(models)

    def get_performance_metrics(self):
        return self.PT.port_metrics(self.proba, self.n, self.targets)


def cpu_reference_prediction(steps, warmup, s_ref=REF_S, n_ref=REF_N):
    """The reference's BMA evaluation on the host cores: each step = Prediction(...) + update_statistics(S_ref modules
    over N_ref images) + get_performance_metrics(), i.e. S_ref * N_ref pairs.  Returns img*samples/s and the sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    RefPrediction, ctor, _, kind = _reference_modules()
    if kind == "port":
        from ursabench_b200 import models as M
        ctor = lambda: M.PreResNet(num_classes=NUM_CLASSES, depth=DEPTH)      # noqa: E731
        RefPrediction = _PortPrediction
    models = _samples(ctor, s_ref)
    x, y = _host_data(n_ref, pin=False)
    loader = {"in_distribution_test": _loader(x, y)}

    def one():
        task = RefPrediction(loader, NUM_CLASSES, torch.device("cpu"), METRICS)
        task.update_statistics(models, output_performance=False)
        return task.get_performance_metrics()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = one()
    dt = time.perf_counter() - t0
    pairs = s_ref * n_ref
    return {"value": steps * pairs / dt, "unit": "img*samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%d steps x (S=%d samples x N=%d images, batch %d) of tasks.Prediction on the host CPU in %.1f s; "
                      "linear in S*N, so the S=100 x N=10000 figure is this rate (extrapolated, BASELINE.md 4)"
                      % (steps, s_ref, n_ref, BATCH, dt),
            "ms_per_step": dt / steps * 1e3, "metrics_of_sample": {k: float(v) for k, v in out.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_prediction(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "bma_img_samples_per_s", "value": cb["value"], "unit": "img*samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "img*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "img_per_s_over_S_samples": cb["value"] / S_ALL,
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
    from ursabench_b200 import _C, dist as udist, models, tasks
    from ursabench_b200.bank import SampleBank
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU arm)")
    rank, world = udist.init_from_env("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _C.lib()
    K, W = args.steps, max(args.warmup, 3)
    S, N = args.samples, args.images

    # ---- host side of the workload: pinned test set + S posterior samples as CPU modules (what sample() hands a user) ----
    ctor = lambda: models.PreResNet(num_classes=NUM_CLASSES, depth=DEPTH)      # noqa: E731
    host_models = _samples(ctor, S)
    x_host, y_host = _host_data(N)
    loaders = {"in_distribution_test": _loader(x_host, y_host)}
    dist_on = world > 1

    def barrier():
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: everything resident (test set uploaded by the task, samples in a device bank replicated on each rank) ----
    task = tasks.Prediction(loaders, NUM_CLASSES, dev, METRICS, distributed=dist_on, replicated_samples=True)
    bank = SampleBank.from_modules(host_models, dev)
    bank.skeleton = ctor()
    y_dev = task._y

    def resident_step():
        task.reset()
        task._entropy.zero_()
        task.update_from_bank(bank)
        proba, _, n_s = task._reduced()                      # ONE all-reduce when world > 1
        return _C.bma_metrics(proba, n_s, y_dev)             # K4 counters stay on the device

    for _ in range(W):
        resident_step()
    sampler = _ClockSampler(local)
    barrier()
    l0 = _C.launch_count()
    _C.profile_begin()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        oi, of, _, _ = resident_step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    prof = _C.profile_end()
    my_launches = _C.launch_count() - l0
    t_ms = udist.allreduce_max_scalar(e0.elapsed_time(e1), dev)
    value = K * S * N / (t_ms / 1e3)
    algo = task.last_algo
    counters_dev = (oi.clone(), of.clone())
    proba_dev = task._reduced()[0].clone()

    # ---- verification: the sharded + all-reduced result equals the single-rank evaluation ----------------------
    verify = None
    if dist_on:
        solo = tasks.Prediction(loaders, NUM_CLASSES, dev, METRICS, distributed=False)
        solo.update_from_bank(bank)
        so_i, so_f, _, _ = _C.bma_metrics(solo._proba, S, y_dev)
        diff = float((proba_dev / S - solo._proba / S).abs().max())
        verify = {"max_abs_diff_mean_proba_vs_single_rank": diff, "counters_equal": bool(torch.equal(so_i, counters_dev[0])),
                  "counters_max_abs_diff": int((so_i - counters_dev[0]).abs().max())}
        if diff > 1e-5:
            raise SystemExit("bench.py: sharded BMA differs from the single-rank result by %.3g (> 1e-5)" % diff)
        del solo

    # ---- e2e: from host data through the public API, metrics dict read back, every step --------------------------
    h2d = [0]

    def host_step():
        t = tasks.Prediction(loaders, NUM_CLASSES, dev, METRICS, distributed=dist_on, replicated_samples=True)
        t.update_statistics(host_models, output_performance=False)
        out = t.get_performance_metrics()
        h2d[0] = t.h2d_bytes + y_host.numel() * 8 + t.h2d_sample_bytes
        return out

    del task
    for _ in range(min(W, 3)):
        host_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        metrics = host_step()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = udist.allreduce_max_scalar(max(e0.elapsed_time(e1), wall_ms), dev)
    e2e_value = K * S * N / (e2e_ms / 1e3)
    d2h = (1 + 2 * 15) * 8 + (2 + 15) * 8

    # ---- roofline of the dominant kernel, from the event pairs recorded inside the timed steps ---------------------
    peaks, peak_src = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    pairs_rank = sum(hi - lo for _, lo, hi in udist.shard_pairs(S, N, rank, world))
    per_kernel = {}
    for kind, (ms, cnt) in prof.items():
        if cnt:
            per_kernel[kind] = {"launches": cnt, "ms_total": ms, "us_per_launch": ms * 1e3 / cnt,
                                "share_of_step": ms / t_ms}
    dom = "stage_c16"
    roofline = None
    if dom in per_kernel:
        ms, cnt = prof[dom]
        flops_per_launch = STAGE_FLOP[dom] * pairs_rank * K / cnt
        achieved = flops_per_launch / (ms / cnt * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "stage16_ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "preresnet_stage16_kernel<16> (stem + the 6 fused 3x3 convs of stage 1, 2xFP16-split tcgen05)",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                    "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peak_src,
                    "flops_per_launch": flops_per_launch, "ms_per_launch": ms / cnt, "launches": cnt,
                    "note": "algorithmic = fp32-equivalent 2*MAC (29.2 MFLOP per pair); the tensor pipe issues 3 FP16 products "
                            "per MAC, i.e. %.1f TFLOP/s of MMA work" % (3 * achieved)}
        stage_ms = sum(prof[k][0] for k in STAGE_FLOP)
        roofline["all_stage_kernels"] = {"achieved": sum(STAGE_FLOP.values()) * pairs_rank * K / (stage_ms * 1e-3) / 1e12,
                                         "ms_per_step": stage_ms / K}
        roofline["whole_forward"] = {"achieved": FLOP_PER_PAIR * S * N / (t_ms / K * 1e-3) / 1e12 / world,
                                     "note": "81.63 MFLOP x pairs / step time, per GPU"}
    line = {
        "metric": "bma_img_samples_per_s", "value": value, "unit": "img*samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": CONFIG,
        "img_per_s_over_S_samples": value / S,
        "engine": {"algo": int(algo) if algo is not None else None,
                   "name": "URSA_ALGO_TCGEN05_FUSED_F16 (2xFP16-split operands, fp32 accumulate in TMEM)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "img*samples/s", "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K,
                "api": "tasks.Prediction({'in_distribution_test': DataLoader(pinned host tensors)}, ...) + "
                       "update_statistics(list of CPU nn.Modules) + get_performance_metrics() -> dict"},
        "gpu_launches": my_launches,
        "roofline": roofline,
        "per_kernel": per_kernel,
        "metrics": {k: float(v) for k, v in metrics.items()},
    }
    if verify is not None:
        line["verify"] = verify

    if not args.skip_extras:
        try:
            line["extras"] = run_extras(dev, rank, world, float(peaks["hbm_gbs"]))
        except Exception as e:  # noqa: BLE001  (extras never invalidate the headline line)
            line["extras"] = {"error": repr(e)}
    if rank == 0 and world == 1:
        cb = cpu_reference_prediction(steps=5, warmup=1)          # ~10 s of host work
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if not args.skip_extras:
            try:
                line["extras"]["cpu"] = cpu_extras()
            except Exception as e:  # noqa: BLE001
                line["extras"]["cpu"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist_on:
        torch.distributed.destroy_process_group()


HYP_CSGHMC = {"lr_0": 0.5, "prior_std": 0.5, "num_samples_per_cycle": 3, "cycle_length": 50, "burn_in_epochs": 0,
              "num_cycles": 17, "alpha": 0.5}     # hyperparams/WideResNet28x10CIFAR10/csghmc_hyperparams.json shape


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


def sampler_extras(dev, rank, world, peak):
    """BASELINE.json configs[1] (last round's contract line, now an extra): one cSGHMC chain per GPU on PreResNet-20,
    batch 128 -- forward + backward on stock PyTorch / cuDNN + ONE fused K1 launch -- and, next to it, what the K1 launch
    replaces: the reference's own optimSGHMC.step (7-8 ATen launches per parameter tensor) run on the same GPU."""
    from ursabench_b200 import _C, dist as udist, inference, models
    out = {}
    torch.manual_seed(0)
    model = models.PreResNet(num_classes=NUM_CLASSES, depth=DEPTH).to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    pool = 96                                                        # 96 x 1.57 MB of distinct batches (> L2)
    hx = torch.randn(pool, BATCH, 3, 32, 32, generator=g).pin_memory()
    hy = torch.randint(0, NUM_CLASSES, (pool, BATCH), generator=g).pin_memory()
    dx, dy = hx.to(dev), hy.to(dev)
    loader = _ListLoader([(hx[i], hy[i]) for i in range(pool)], 50_000, BATCH)
    torch.manual_seed(1234 + rank)
    inf = inference.cSGHMC(dict(HYP_CSGHMC), model, loader, device=dev)
    inf.optimizer.elem_offset = udist.chain_elem_offset(rank, inf.flat.D)
    inf.model.train()
    inf.enable_cuda_graph(dx[0], dy[0])
    n = [0]

    def step(host):
        i = n[0] % pool
        inf._adjust_learning_rate(inf.optimizer, 0, n[0] % 391)
        n[0] += 1
        if host:
            return inf.train_step(hx[i], hy[i], True).item()
        return inf.train_step(dx[i], dy[i], True)
    for host, name in ((False, "resident"), (True, "e2e_host_batches_loss_item")):
        for _ in range(5):
            step(host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(60):
            step(host)
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 if host else 0.0)
        ms = udist.allreduce_max_scalar(ms, dev)
        out["csghmc_preresnet20_b128_steps_per_s_" + name] = {"value": world * 60 / ms * 1e3, "ms_per_step": ms / 60,
                                                              "chains": world, "note": "fwd/bwd = PyTorch autograd + cuDNN; K1 = 1 launch"}
    inf.disable_cuda_graph()
    del inf, dx, dy
    # K1 on the HBM-bound chain-batched layout [128 chains, D = 272 284]
    D = 272_282
    ld = (D + 3) // 4 * 4
    nb = 128 * ld
    pb, gb, vb = (torch.randn(nb, device=dev) for _ in range(3))
    kw = dict(lr=0.1, momentum=0.5, wd_over_n=4.0 / 50_000, noise_mul=math.sqrt(2 * 0.5 * 0.1), noise_div=5e4, seed=7)
    c = [0]

    def k1():
        c[0] += 1
        _C.sgmcmc_step(pb, gb, vb, step=c[0], **kw)
    for _ in range(5):
        k1()
    ms, _ = _event_time_ms(k1, 30)
    out["k1_sghmc_128chains_preresnet20"] = {"ms": ms, "GBps": 20 * nb / ms / 1e6, "frac": 20 * nb / ms / 1e6 / peak,
                                             "bytes_per_launch": 20 * nb}
    del pb, gb, vb
    # the reference's optimSGHMC.step itself on this GPU (stock ATen), WRN-28-10-sized parameter list, momentum 0.5, noise on
    if rank == 0:
        _, _, ref_opt, kind = _reference_modules()
        wrn = models.WideResNet(num_classes=100, depth=28, widen_factor=10).to(dev)
        for q in wrn.parameters():
            q.grad = torch.randn_like(q)
        if ref_opt is not None:
            opt = ref_opt(wrn.parameters(), lr=0.01, momentum=0.5, weight_decay=1e-4 * 5e4, num_training_samples=50_000)
            fn = lambda: opt.step(add_langevin_noise=True)      # noqa: E731
        else:
            from oracle import port_torch as PT
            opt = PT.PortOptimSGHMC(wrn.parameters(), 0.01, 0.5, 1e-4 * 5e4, 50_000)
            fn = lambda: opt.step(add_langevin_noise=True)      # noqa: E731
        for _ in range(3):
            fn()
        ms, _ = _event_time_ms(fn, 10)
        Dw = sum(q.numel() for q in wrn.parameters())
        out["reference_optimSGHMC_step_on_gpu_wrn28x10"] = {"ms": ms, "steps_per_s": 1e3 / ms, "kind": kind,
                                                            "GBps_at_20B_per_param": 20 * Dw / ms / 1e6,
                                                            "note": "stock ATen, 108 parameter tensors x 7-8 launches; compare k1_sghmc_wrn28x10"}
        del wrn, opt
    torch.cuda.empty_cache()
    return out


def cpu_extras():
    """Host-CPU timings of the reference's code for the other configs, each on a bounded sample; the
    matching GPU figures are the same-named entries of ``extras``.  Unmodified reference classes when staged
    (kind "reference"), else the torch-CPU port."""
    from oracle import port_torch as PT
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    RefPrediction, _, ref_opt, kind = _reference_modules()
    out = {"cores": cores, "kind": kind}

    def timeit(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    # config 3 sizes: optimSGHMC.step, SWA._collect_model and the (restated) rank-20 draw at D = 36.5 M
    D = 36_546_980
    p = torch.nn.Parameter(torch.randn(D) * 0.05)
    p.grad = torch.randn(D)
    if ref_opt is not None:
        opt = ref_opt([p], lr=0.01, momentum=0.5, weight_decay=5.0, num_training_samples=50_000)
    else:
        opt = PT.PortOptimSGHMC([p], 0.01, 0.5, 5.0, 50_000)
    t = timeit(lambda: opt.step(add_langevin_noise=True), 3)
    out["optimSGHMC_step_D36.5M"] = {"ms": t * 1e3, "steps_per_s": 1 / t, "GBps_at_20B_per_param": 20 * D / t / 1e9,
                                     "sample": "3 steps, one flat tensor of D = 36 546 980, momentum 0.5, noise on"}
    w = p.detach()
    mean, sq = torch.zeros(D), torch.zeros(D)
    t = timeit(lambda: PT.port_swa_collect(w, mean, sq, 3), 3)
    out["swa_collect_D36.5M"] = {"ms": t * 1e3, "GBps_at_24B_per_param": 24 * D / t / 1e9,
                                 "sample": "3 collects (inference/swa.py:79-90 op sequence), D = 36 546 980", "kind": "port"}
    K = 20
    ring = torch.randn(K, D) * 0.01
    var = torch.full((D,), 1e-4)
    t = timeit(lambda: PT.port_swag_draw(mean, var, ring, K, 1), 2)
    out["swag_draw_K20_D36.5M"] = {"ms_per_draw": t * 1e3, "ms_for_30_draws_extrapolated": 30 * t * 1e3,
                                   "sample": "2 single draws (inference/swag.py:88-97 restated; the shipped branch raises), K = 20",
                                   "kind": "port"}
    del ring, var, mean, sq, p, w
    # config 1: Prediction on MLP 784-400-400-10, bounded S = 5 x N = 2000
    from ursabench_b200 import models as M
    torch.manual_seed(0)
    mlps = [M.MLP(400, 784, 10) for _ in range(5)]
    g = torch.Generator().manual_seed(3)
    xm, ym = torch.randn(2000, 1, 28, 28, generator=g), torch.randint(0, 10, (2000,), generator=g)
    lm = {"in_distribution_test": torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xm, ym), batch_size=BATCH)}
    Pred = RefPrediction if RefPrediction is not None else _PortPrediction

    def mlp_eval():
        tk = Pred(lm, 10, torch.device("cpu"), METRICS)
        tk.update_statistics(mlps, output_performance=False)
        return tk.get_performance_metrics()
    t = timeit(mlp_eval, 3)
    out["bma_mlp400"] = {"img_samples_per_s": 5 * 2000 / t, "sample": "3 x (S = 5 x N = 2000) through tasks.Prediction"}
    # config 2: one cSGHMC step (PreResNet-20 fwd + bwd + optimSGHMC.step), batch 128
    from oracle import restate as R
    torch.manual_seed(0)
    model = M.PreResNet(num_classes=NUM_CLASSES, depth=DEPTH)
    if ref_opt is not None:
        opt = ref_opt(model.parameters(), lr=0.5, momentum=0.5, weight_decay=4.0, num_training_samples=50_000)
    else:
        opt = PT.PortOptimSGHMC(model.parameters(), 0.5, 0.5, 4.0, 50_000)
    xb, yb = torch.randn(BATCH, 3, 32, 32, generator=g), torch.randint(0, 10, (BATCH,), generator=g)
    crit = torch.nn.CrossEntropyLoss()
    nbatch = R.csghmc_num_batch(50_000, BATCH)
    model.train()
    i = [0]

    def sgmcmc_step():
        lr = PT.port_csghmc_lr(0.5, 0, i[0], nbatch, 50, 17)
        i[0] += 1
        if ref_opt is not None:
            for gq in opt.param_groups:
                gq["lr"] = lr
        else:
            opt.lr = lr
        logits = model(xb)
        opt.zero_grad()
        loss = crit(logits, yb)
        loss.backward()
        loss.item()
        opt.step(add_langevin_noise=True)
    t = timeit(sgmcmc_step, 20, warm=2)
    out["csghmc_preresnet20_b128"] = {"steps_per_s": 1 / t, "sample": "20 steps (csghmc.py:80-93 loop body), one chain"}
    # config 3: Prediction on WideResNet-28-10 (C = 100), bounded S = 1 x N = 64 (11.9 GFLOP per image)
    try:
        from URSABench.models import wideresnet as rwrn
        wrn, wkind = rwrn.WideResNet(num_classes=100, depth=28, widen_factor=10), kind
    except Exception:  # noqa: BLE001
        wrn, wkind = M.WideResNet(num_classes=100, depth=28, widen_factor=10), "port"
    wrn.eval()
    xw, yw = torch.randn(64, 3, 32, 32, generator=g), torch.randint(0, 100, (64,), generator=g)
    lw = {"in_distribution_test": torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xw, yw), batch_size=64)}

    def wrn_eval():
        tk = Pred(lw, 100, torch.device("cpu"), METRICS)
        tk.update_statistics([wrn], output_performance=False)
        return tk.get_performance_metrics()
    t = timeit(wrn_eval, 2)
    out["bma_wrn28x10"] = {"img_samples_per_s": 64 / t, "ms_for_S30_N10k_extrapolated": 30 * 10_000 / 64 * t * 1e3, "kind": wkind,
                           "sample": "2 x (S = 1 x N = 64) through tasks.Prediction"}
    # config 4: one HMC leapfrog step of ONE chain = autograd gradient of the MLP-200 log-likelihood over 1 000 points + the
    # kick / drift updates (what hamiltorch.sample does L times per iteration and chain, inference/hmc.py:71-75; hamiltorch is
    # not vendored, so this is the restated loop body on torch CPU: kind "port")
    mlp = M.MLP(200, 784, 10)
    xh, yh = torch.randn(1000, 784, generator=g), torch.randint(0, 10, (1000,), generator=g)
    params = list(mlp.parameters())
    mom = [torch.randn_like(q) for q in params]
    ce = torch.nn.CrossEntropyLoss(reduction="sum")

    def leapfrog():
        grads = torch.autograd.grad(ce(mlp(xh), yh), params)
        with torch.no_grad():
            for q, r, gq in zip(params, mom, grads):
                r.add_(gq + 100.0 * q, alpha=-2.09e-4)
                q.add_(r, alpha=2.09e-4 / 0.192)
    t = timeit(leapfrog, 50, warm=3)
    out["hmc_mlp200_N1000"] = {"chain_leapfrog_steps_per_s": 1 / t, "kind": "port",
                               "sample": "50 leapfrog steps of one chain (gradient over 1 000 points + kick / drift), torch CPU"}
    return out


def run_extras(dev, rank, world, peak):
    """Secondary figures of the same path (not the contract metric): K1 at WRN size, SWAG collect / draw,
    BMA img/s with samples sharded over ranks + the single all-reduce."""
    from ursabench_b200 import _C, dist as udist, models
    out = {}
    out.update(sampler_extras(dev, rank, world, peak))
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
    fn = lambda: _C.swag_draw(bank, mean, var, D, seed=5, step=1)  # noqa: E731     (SWAG-Diag: no low-rank term)
    for _ in range(3):
        fn()
    ms, _ = _event_time_ms(fn, 10)
    out["k2_draw_S30_diag_wrn28x10"] = {"ms": ms, "GBps": (2 + S) * 4 * D / ms / 1e6, "frac": (2 + S) * 4 * D / ms / 1e6 / peak,
                                        "us_per_draw": ms * 1e3 / S}
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

    for algo_name, algo in (("ffma", _C.ALGO_FFMA), ("tcgen05", _C.ALGO_TCGEN05), ("f16", _C.ALGO_TCGEN05_F16)):
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
        ms, _ = _event_time_ms(bma_conv, 1)
        ms = udist.allreduce_max_scalar(ms, dev)
        out["bma_preresnet20_S100_N10k_" + algo_name] = {"ms": ms, "img_per_s_over_S_samples": N / ms * 1e3,
                                                         "img_samples_per_s": N * S_all / ms * 1e3,
                                                         "TFLOPs": 81.63e6 * N * S_all / ms / 1e9, "n_gpus": world}
    # K3b for the north-star model: BatchNorm re-estimation (util.bn_update, once per SWAG sample) of 8 draws, sample-batched,
    # on a bounded sample of the 50 000-image pass; beside it the PyTorch / cuDNN fp32 train-mode pass it replaces
    Nbn, Bbn, Sbn = 2048, 128, min(8, ns) if ns > 0 else 0
    if Sbn > 0:
        from ursabench_b200.util import bn_update as torch_bn_update
        bn_ws = [None]
        bn_fn = lambda: bn_ws.__setitem__(0, _C.preresnet_bn_update(bankp[:Sbn], bufp[:Sbn].clone(), xi[:Nbn], Bbn, 20, 10, workspace=bn_ws[0]))  # noqa: E731
        bn_fn()
        torch.cuda.synchronize()
        ms, _ = _event_time_ms(bn_fn, 3)
        out["k3b_bn_update_preresnet20_S%d_N2048_b128" % Sbn] = {"ms": ms, "img_samples_per_s": Sbn * Nbn / ms * 1e3,
                                                                 "full_pass_50k_images_per_sample_s": ms * 50_000 / Nbn / Sbn / 1e3}
        if rank == 0:
            mt = m.to(dev)
            tf32c, tf32m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            try:
                ld_bn = [(xi[i:i + Bbn], None) for i in range(0, Nbn, Bbn)]
                torch_bn_update(ld_bn, mt, device=dev)
                tms, _ = _event_time_ms(lambda: torch_bn_update(ld_bn, mt, device=dev), 2)
            finally:
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32c, tf32m
            out["k3b_bn_update_preresnet20_torch_cudnn_fp32"] = {"sample": "1 sample x 2048 images, batch 128", "ms": tms,
                                                                 "img_samples_per_s": Nbn / tms * 1e3}
            m.cpu()
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

    for algo_name, algo in (("tcgen05", _C.ALGO_TCGEN05), ("f16", _C.ALGO_TCGEN05_F16)):
        def bma_wrn():
            P.zero_()
            E.zero_()
            if ns > 0:
                ws[0] = _C.bma_wrn_forward(bankw, bufw, ns, xi, 28, 10, C_w, P, E, workspace=ws[0], algo=algo)
            Pr, Er, n = udist.allreduce_bma(P, E, ns)
            return _C.bma_metrics(Pr, n, yw)
        ws[0] = _C.bma_wrn_forward(bankw[:1], bufw[:1], 1, xi[:512], 28, 10, C_w, P[:512], E[:512], workspace=None, algo=algo)   # warm-up
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        wsampler = _ClockSampler(dev.index if dev.index is not None else 0)
        wsampler.start()
        ms, _ = _event_time_ms(bma_wrn, 1)
        wclocks = wsampler.stop()                           # a 5-17 s tensor-core run: the SM clock under its own power draw
        ms = udist.allreduce_max_scalar(ms, dev)
        out["bma_wrn28x10_S30_N10k_" + algo_name] = {"ms": ms, "img_per_s_over_S_samples": N / ms * 1e3,
                                                     "img_samples_per_s": N * S_w / ms * 1e3,
                                                     "TFLOPs": wrn_flop * N * S_w / ms / 1e9, "n_gpus": world, "clocks": wclocks}
    # K3b: BatchNorm re-estimation of one SWAG draw (util.bn_update, once per sample): bounded sample of the 50 000-image pass
    Nbn, Bbn = 2048, 128
    for algo_name, algo in (("", _C.ALGO_TCGEN05), ("_f16", _C.ALGO_TCGEN05_F16)):
        bn_fn = lambda: _C.wrn_bn_update(bankw[0], bufw[0], xi[:Nbn], Bbn, 28, 10, C_w, workspace=ws[0], algo=algo)  # noqa: E731
        ws[0] = None
        ws[0] = bn_fn()
        torch.cuda.synchronize()
        ms, _ = _event_time_ms(bn_fn, 2)
        out["k3b_bn_update_wrn28x10_N2048_b128" + algo_name] = {"ms": ms, "img_per_s": Nbn / ms * 1e3,
                                                                "TFLOPs": wrn_flop * Nbn / ms / 1e9,
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
    hyp = {"step_size": 2.09e-4, "num_samples": 8, "L": Lh, "tau": 100.0, "burn": 0, "mass": 0.192, "num_chains": Ch}
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
    steady = udist.allreduce_max_scalar(hm.ms_per_iteration, dev)
    out["hmc_mlp200_N1000"] = {"chains_per_gpu": Ch, "n_gpus": world, "L": Lh, "ms_per_iteration": steady,
                               "ms_per_iteration_whole_call": ms / hyp["num_samples"],
                               "chain_leapfrog_steps_per_s": world * Ch * Lh / steady * 1e3,
                               "grad_TFLOPs": 6 * (784 * 200 + 200 * 200 + 200 * 10) * 1000 * Ch * (Lh + 1) / steady / 1e9,
                               "accept_rate": float(hm.acceptance_rate.mean()), "grad_engine": hm.grad_engine,
                               "graph_replays": hm.graph_replays,
                               "note": "ms_per_iteration = device time of iterations 3..8 (CUDA-graph replays of the whole leapfrog trajectory); the whole-call figure also "
                                       "carries chain initialisation on the host, the eager first iteration and the graph capture"}
    return out




def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--samples", type=int, default=S_ALL, help="posterior samples S (default: the contract's 100)")
    ap.add_argument("--images", type=int, default=N_TEST, help="test images N (default: the contract's 10000)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
