"""``Prediction``: S-sample Bayesian-model-averaging evaluation (reference tasks/prediction.py:12-267).

Reference hot loop (:52-75): for every test batch, for every sample: move the whole model to the device, forward,
softmax twice, two device->host copies, accumulate on the CPU, move the model back.  Here the test set is uploaded
once, the posterior samples are rows of a device bank, the forward is sample-batched (fused K3 kernels for MLP and
PreResNet; any other architecture runs its own PyTorch forward on the device), softmax-average and entropy are
accumulated on the device in sample order by ``ursa_bma_accumulate``, and the metric counters come from
``ursa_bma_metrics``.  With ``distributed=True`` the accumulators of all ranks are summed with one all-reduce.

Same constructor, attributes and return types as the reference; metric values follow its arithmetic
(fp32 ``p / S``, first-max argmax, (lo, hi] bins on float64 edges, smoothing gamma = 1e-4).
"""
import copy

import numpy as np
import torch

from .. import _C, dist as udist, util
from ..bank import BankedSample, SampleBank
from ..flat import FlatParams
from .task_base import _Task

__all__ = ["Prediction"]

_LOGIT_CHUNK_BYTES = 256 << 20


def _arch_of(module):
    """('mlp', in_dim, hidden, C) / ('preresnet', depth, C) / None -- structural match against models.py."""
    name = type(module).__name__
    if name == "MLP" and all(hasattr(module, a) for a in ("fc1", "fc2", "fc3")):
        f1, f2, f3 = module.fc1, module.fc2, module.fc3
        if all(isinstance(f, torch.nn.Linear) and f.bias is not None for f in (f1, f2, f3)) \
                and f2.in_features == f1.out_features == f2.out_features == f3.in_features \
                and len(list(module.parameters())) == 6:
            return ("mlp", f1.in_features, f1.out_features, f3.out_features)
    if name == "PreResNet" and hasattr(module, "layer1") and hasattr(module, "fc"):
        blocks = list(module.layer1)
        if blocks and type(blocks[0]).__name__ == "BasicBlock" and module.fc.in_features == 64:
            return ("preresnet", 6 * len(blocks) + 2, module.fc.out_features)
    return None


class Prediction(_Task):
    supported_metric_list = ["error_rate", "nll", "ll", "brier_score", "ece", "misclass_model_uncertainty_auroc",
                             "misclass_model_uncertainty_aucpr", "misclass_total_uncertainty_auroc",
                             "misclass_total_uncertainty_aucpr", "misclass_confidence_auroc",
                             "misclass_confidence_aucpr"]

    def __init__(self, dataloader, num_classes, device, metric_list, distributed=False, engine="auto"):
        super().__init__(dataloader, num_classes, device)
        _C.lib()
        self.data_loader = dataloader["in_distribution_test"]
        self.num_classes = num_classes
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Prediction: device must be a CUDA device -- ursabench_b200 has no CPU path")
        self.distributed = distributed
        self.engine = engine      # 'auto' | 'ffma' (pin the fp32 CUDA-core kernels) | 'generic' (per-sample PyTorch forward)
        self._ws = None           # K3 workspace, kept across calls
        self.last_algo = None
        self.num_samples_collected = 0
        self.required_metric_list = self.supported_metric_list if metric_list == "ALL" else metric_list
        assert all(metric in self.supported_metric_list for metric in self.required_metric_list)
        # one pass over the loader (the reference does the same to cache the targets, :28-31); the inputs are
        # uploaded once and stay resident -- the loader must not shuffle
        xs, ys = [], []
        for batch_data, batch_labels in self.data_loader:
            xs.append(batch_data)
            ys.append(batch_labels)
        self.targets = torch.cat(ys)
        self._n = len(self.data_loader.dataset)
        self._batch_sizes = [len(x) for x in xs]
        self._x = torch.cat(xs).to(self.device, non_blocking=True).float().contiguous()
        self._y = self.targets.to(self.device).long().contiguous()
        self._proba = torch.zeros(self._n, num_classes, device=self.device)
        self._entropy = torch.zeros(self._n, device=self.device)
        self._workers = {}
        self.last_engine = None
        self.kernel_launches = 0

    # -- reference-compatible attribute views (CPU tensors in the reference) ---------------------------------
    @property
    def ensemble_proba(self):
        return self._proba.cpu()

    @ensemble_proba.setter
    def ensemble_proba(self, value):
        self._proba = value.to(self.device, dtype=torch.float32).contiguous()

    @property
    def expected_data_uncertainty(self):
        return self._entropy.cpu()

    @expected_data_uncertainty.setter
    def expected_data_uncertainty(self, value):
        self._entropy = value.to(self.device, dtype=torch.float32).contiguous()

    def reset(self):
        """reference :33-35 -- NB it does not clear ``expected_data_uncertainty``; kept."""
        self.num_samples_collected = 0
        self._proba = torch.zeros(self._n, self.num_classes, device=self.device)

    # -- accumulation ------------------------------------------------------------------------------------------
    def update_statistics(self, models, output_performance=True, smoothing=True):
        if isinstance(models, list):
            if not all(isinstance(m, torch.nn.Module) for m in models):
                raise NotImplementedError
            model_list = models
        elif isinstance(models, torch.nn.Module):
            model_list = [models]
        else:
            raise NotImplementedError
        self.num_samples_collected += len(model_list)
        if model_list:
            with torch.no_grad():
                self._accumulate(model_list)
        if output_performance:
            return self.get_performance_metrics(output_performance, smoothing)

    def update_from_bank(self, bank, rows=None):
        """Fast path without module handles: evaluate bank rows directly (used by bench.py / multi-GPU drivers)."""
        rows = list(range(bank.count)) if rows is None else list(rows)
        self.num_samples_collected += len(rows)
        if rows:
            w, b = bank.rows(rows)
            arch = _arch_of(bank.skeleton) if bank.skeleton is not None else None
            with torch.no_grad():
                self._accumulate_rows(w, b, arch, bank.skeleton)

    def _accumulate(self, model_list):
        banked = all(isinstance(m, BankedSample) and m.is_pristine() for m in model_list)
        if banked and len({id(m._ursa_bank) for m in model_list}) == 1:
            bank = model_list[0]._ursa_bank
            w, b = bank.rows([m._ursa_row for m in model_list])
            self._accumulate_rows(w, b, _arch_of(bank.skeleton), bank.skeleton)
            return
        plain = [m.materialize() if isinstance(m, BankedSample) else m for m in model_list]
        arch = _arch_of(plain[0])
        if self.engine in ("auto", "ffma") and arch is not None and self._fused_available(arch) \
                and all(_arch_of(m) == arch for m in plain):
            bank = SampleBank.from_modules(plain, self.device)       # one H2D per sample instead of 2 per batch
            self._accumulate_rows(bank.w[:bank.count], bank.b[:bank.count], arch, None)
            return
        self._accumulate_generic_modules(plain)

    def _fused_available(self, arch):
        return self._pick_algo(arch) is not None

    def _pick_algo(self, arch):
        """Fastest engine whose workspace query accepts the shape: the tcgen05 paths (3xTF32, fp32-level accuracy,
        parity-tested at the same 1e-5 bar) first, the fp32 CUDA-core kernels for shapes they do not cover
        (e.g. an MLP width that is not a multiple of 4).  ``engine='ffma'`` pins the CUDA-core kernels."""
        lib = _C.lib()
        if arch[0] == "mlp":
            order = (_C.ALGO_FFMA,) if self.engine == "ffma" else (_C.ALGO_TCGEN05, _C.ALGO_FFMA)
            for algo in order:
                if lib.ursa_bma_mlp_workspace(1, 1, arch[1], arch[2], arch[3], algo) > 0:
                    return algo
        elif arch[0] == "preresnet":
            order = (_C.ALGO_FFMA,) if self.engine == "ffma" else (_C.ALGO_TCGEN05_FUSED, _C.ALGO_TCGEN05, _C.ALGO_FFMA)
            for algo in order:
                if lib.ursa_bma_preresnet_workspace(1, 1, arch[1], arch[2], algo) > 0:
                    return algo
        return None

    def _accumulate_rows(self, w, b, arch, skeleton):
        S = w.shape[0]
        if self.engine in ("auto", "ffma") and arch is not None and self._fused_available(arch):
            algo = self._pick_algo(arch)
            if arch[0] == "mlp":
                _, in_dim, hidden, C = arch
                x2 = self._x.view(self._n, -1)
                if x2.shape[1] != in_dim or C != self.num_classes:
                    raise ValueError("MLP input / class dimensions do not match the task")
                self._ws = _C.bma_mlp_forward(w, S, x2, in_dim, hidden, C, self._proba, self._entropy, algo=algo,
                                              workspace=self._ws)
                self.last_engine = "fused_mlp"
            else:
                _, depth, C = arch
                self._ws = _C.bma_preresnet_forward(w, b, S, self._x, depth, C, self._proba, self._entropy, algo=algo,
                                                    workspace=self._ws)
                self.last_engine = "fused_preresnet"
            self.last_algo = algo
            self.kernel_launches += 1
            return
        if skeleton is None:
            raise RuntimeError("no module skeleton available for the generic forward")
        worker = self._worker_for(skeleton)
        flat = worker._ursa_worker_flat

        def load(i):
            flat.load_vector(w[i])
            flat.load_buffers(b[i])
            return worker

        self._accumulate_generic(S, load)

    def _worker_for(self, skeleton):
        key = id(skeleton)
        if key not in self._workers:
            worker = copy.deepcopy(skeleton).to(self.device)
            worker._ursa_worker_flat = FlatParams.from_model(worker, self.device)
            worker.eval()
            self._workers[key] = worker
        return self._workers[key]

    def _accumulate_generic_modules(self, plain):
        homes = [next(m.parameters()).device for m in plain]

        def load(i):
            plain[i].to(self.device)
            plain[i].eval()
            return plain[i]

        def unload(i):
            plain[i].to(homes[i])

        self._accumulate_generic(len(plain), load, unload)

    def _accumulate_generic(self, S, load, unload=None):
        """Per-sample PyTorch forward on the device over the resident test set; logits are gathered per chunk of
        samples and reduced by ONE ``ursa_bma_accumulate`` launch per chunk (sample order preserved)."""
        N, C = self._n, self.num_classes
        chunk = max(1, min(S, _LOGIT_CHUNK_BYTES // max(1, N * C * 4)))
        logits = torch.empty(chunk, N, C, device=self.device)
        tf32_matmul = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False          # fp32 parity with the reference's CPU forward
        try:
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
                self._generic_chunks(S, chunk, logits, load, unload)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_matmul
        self.last_engine = "generic"

    def _generic_chunks(self, S, chunk, logits, load, unload):
        for s0 in range(0, S, chunk):
            ns = min(chunk, S - s0)
            for j in range(ns):
                model = load(s0 + j)
                off = 0
                for bs in self._batch_sizes:
                    out = model(self._x[off:off + bs])
                    logits[j, off:off + bs] = out.float()
                    off += bs
                if unload is not None:
                    unload(s0 + j)
            _C.bma_accumulate(logits[:ns], self._proba, self._entropy)
            self.kernel_launches += 1

    # -- metrics --------------------------------------------------------------------------------------------------
    def _reduced(self):
        if self.distributed:
            return udist.allreduce_bma(self._proba, self._entropy, self.num_samples_collected)
        return self._proba, self._entropy, self.num_samples_collected

    def get_counters(self, smoothing=True, n_bins=15):
        """Device metric counters (K4) as numpy: dict(correct, bin_count, bin_correct, bin_conf_sum, nll_sum,
        brier_sum, n)."""
        proba, _, S = self._reduced()
        oi, of, _, _ = _C.bma_metrics(proba, S, self._y, gamma=1e-4 if smoothing else 0.0, n_bins=n_bins)
        self.kernel_launches += 2
        oi, of = oi.cpu().numpy(), of.cpu().numpy()
        return dict(correct=int(oi[0]), bin_count=oi[1:1 + n_bins], bin_correct=oi[1 + n_bins:],
                    bin_conf_sum=of[2:], nll_sum=float(of[0]), brier_sum=float(of[1]), n=self._n)

    def get_performance_metrics(self, output_performance=False, smoothing=True):
        want = self.required_metric_list
        out = {}
        counters = None
        if any(m in want for m in ("error_rate", "nll", "ll", "brier_score", "ece")):
            counters = self.get_counters(smoothing)
        n = self._n
        host = None
        for metric in want:
            if metric == "error_rate":
                out[metric] = 1 - counters["correct"] / n                                   # reference :82-85
            elif metric in ("nll", "ll"):
                nll = counters["nll_sum"] / n                                               # :86-96
                out[metric] = -nll if metric == "ll" else nll
            elif metric == "brier_score":
                out[metric] = counters["brier_sum"] / n                                     # :97-99, :185-194
            elif metric == "ece":
                ece = 0.0                                                                   # :152-182
                for cnt, csum, acc in zip(counters["bin_count"], counters["bin_conf_sum"], counters["bin_correct"]):
                    if cnt > 0:
                        ece += abs(csum / cnt - acc / cnt) * (cnt / n)
                out[metric] = ece
            else:
                if host is None:
                    host = self._host_misclass_inputs()
                out[metric] = self._misclass_metric(metric, *host)
        if output_performance:
            if len(want) != 1:
                raise RuntimeError("Multiple metrics in metric list not suitable for output_performance = True")
            return float(out[want[0]])
        return out

    # AUROC / AUPR are sort-based, N-sized and stay on host sklearn like the reference (:103-142, :197-267)
    def _host_misclass_inputs(self):
        proba, entropy, S = self._reduced()
        pbar = util.central_smoothing((proba / S).cpu()).numpy()
        edu = (entropy / S).cpu().numpy()
        return pbar, self.targets.numpy(), edu

    @staticmethod
    def _misclass_metric(metric, preds, targets, edu):
        from sklearn.metrics import average_precision_score, roc_auc_score
        top1 = torch.from_numpy(preds).topk(1, 1, True, True)[1].view(-1).numpy()
        mis = (top1 != targets)
        if "model_uncertainty" in metric:
            score = np.sum(-preds * np.log(preds), axis=1) - edu
        elif "total_uncertainty" in metric:
            score = np.sum(-preds * np.log(preds), axis=1)
        else:
            score = -preds.max(axis=1)
        fn = roc_auc_score if metric.endswith("auroc") else average_precision_score
        return fn(mis, score)
