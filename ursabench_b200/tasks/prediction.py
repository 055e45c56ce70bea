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
import numpy as np
import torch

from .. import _C, dist as udist, util
from ._engine import BMAAccumulator, _arch_of
from .task_base import _Task

__all__ = ["Prediction"]

class Prediction(_Task, BMAAccumulator):
    supported_metric_list = ["error_rate", "nll", "ll", "brier_score", "ece", "misclass_model_uncertainty_auroc",
                             "misclass_model_uncertainty_aucpr", "misclass_total_uncertainty_auroc",
                             "misclass_total_uncertainty_aucpr", "misclass_confidence_auroc",
                             "misclass_confidence_aucpr"]

    def __init__(self, dataloader, num_classes, device, metric_list, distributed=False, engine="auto",
                 replicated_samples=False):
        """``distributed``: the accumulators of all ranks are summed with one all-reduce before the metrics.
        ``replicated_samples=False``: every rank passes ITS OWN posterior samples (they live where the chain ran) --
        pure sample sharding.  ``replicated_samples=True``: every rank passes the SAME list and evaluates only its
        balanced share of the (sample, image) grid (``dist.shard_pairs``), which keeps all ranks busy when S < world or
        S mod world != 0."""
        super().__init__(dataloader, num_classes, device)
        self.data_loader = dataloader["in_distribution_test"]
        self._setup(self.data_loader, num_classes, device, engine)   # engine='generic' forces the per-sample PyTorch forward
        self._y = self._y_dev.long().contiguous()                    # uploaded by _setup, ahead of the images
        self.distributed = distributed
        self.replicated_samples = bool(replicated_samples)
        self.num_samples_collected = 0
        self.required_metric_list = self.supported_metric_list if metric_list == "ALL" else metric_list
        assert all(metric in self.supported_metric_list for metric in self.required_metric_list)

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
        model_list = self.as_model_list(models)
        if self.distributed and self.replicated_samples and udist.is_distributed():
            rank, world = udist.rank_world()
            # the sample count is all-reduced with the accumulators: every sample is counted once, on rank 0
            self.num_samples_collected += len(model_list) if rank == 0 else 0
            self.accumulate(model_list, udist.shard_pairs(len(model_list), self._n, rank, world))
        else:
            self.num_samples_collected += len(model_list)
            self.accumulate(model_list)
        if output_performance:
            return self.get_performance_metrics(output_performance, smoothing)

    def update_from_bank(self, bank, rows=None):
        """Fast path without module handles: evaluate bank rows directly (used by bench.py / multi-GPU drivers).  With
        ``distributed`` + ``replicated_samples`` the rows are the same on every rank and each rank evaluates its share of
        the (row, image) grid."""
        rows = list(range(bank.count)) if rows is None else list(rows)
        arch = _arch_of(bank.skeleton) if bank.skeleton is not None else None
        if self.distributed and self.replicated_samples and udist.is_distributed():
            rank, world = udist.rank_world()
            self.num_samples_collected += len(rows) if rank == 0 else 0
            pairs = udist.shard_pairs(len(rows), self._n, rank, world)
        else:
            self.num_samples_collected += len(rows)
            pairs = [(i, 0, self._n) for i in range(len(rows))]
        with torch.no_grad():
            i = 0
            while i < len(pairs):
                j = i
                while j + 1 < len(pairs) and pairs[j + 1][1:] == pairs[i][1:]:
                    j += 1
                _, lo, hi = pairs[i]
                if hi > lo:
                    w, b = bank.rows([rows[k] for k, _, _ in pairs[i:j + 1]])
                    self._accumulate_rows(w, b, arch, bank.skeleton, lo, hi)
                i = j + 1
            self._commit_scratch()

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
