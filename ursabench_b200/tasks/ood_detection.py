"""``OODDetection``: out-of-distribution detection from BMA uncertainties (reference tasks/ood_detection.py:11-130).

Reference loop (:52-100): the Prediction loop run twice (in- and out-of-distribution loaders), accumulating the
SMOOTHED probabilities  p~_s = (1-g) softmax_s + g/C  and  entropy(p~_s)  on the CPU.  Here both loaders get a
``BMAAccumulator`` (inputs resident on the device, sample-batched fused forward); the smoothed sum is the affine image
of the engine's unsmoothed accumulator,  sum_s p~_s = (1-g) sum_s softmax_s + S g/C  (util.py:126-134 is affine), and the
entropy accumulator is already  sum_s entropy(p~_s).  Metrics (:103-130): total uncertainty = entropy of the averaged
smoothed probability (no second smoothing, util.py:137-144), model uncertainty = total - mean data uncertainty, AUROC
of "is out-of-distribution" on the host with sklearn like the reference.
"""
import numpy as np
import torch

from .. import util
from ._engine import BMAAccumulator
from .task_base import _Task

__all__ = ["OODDetection"]

_GAMMA = 1e-4                     # util.central_smoothing default (util.py:126)


class OODDetection(_Task):
    def __init__(self, data_loader=None, num_classes=None, device=torch.device("cpu"), engine="auto"):
        super().__init__(data_loader, num_classes, device)
        self.in_distribution_loader = data_loader["in_distribution_test"]
        self.out_distribution_loader = data_loader["out_distribution_test"]
        self.num_classes = num_classes
        self.device = torch.device(device)
        self._in = BMAAccumulator(self.in_distribution_loader, num_classes, device, engine)
        self._out = BMAAccumulator(self.out_distribution_loader, num_classes, device, engine)
        self._clear_metrics()
        self.num_samples_collected = 0

    def _clear_metrics(self):
        self.in_distribution_total_uncertainty = None
        self.out_distribution_total_uncertainty = None
        self.in_distribution_model_uncertainty = None
        self.out_distribution_model_uncertainty = None

    def reset(self):
        """reference :28-37 (unlike Prediction.reset this one also clears the data uncertainty)."""
        for acc in (self._in, self._out):
            acc._proba.zero_()
            acc._entropy.zero_()
        self._clear_metrics()
        self.num_samples_collected = 0

    # -- reference-compatible attribute views (CPU tensors in the reference, :18-21) ---------------------------------
    def _smoothed_sum(self, acc):
        S, C = self.num_samples_collected, self.num_classes
        return ((1.0 - _GAMMA) * acc._proba + S * _GAMMA / C).cpu()

    @property
    def in_distribution_ensemble_proba(self):
        return self._smoothed_sum(self._in)

    @property
    def out_distribution_ensemble_proba(self):
        return self._smoothed_sum(self._out)

    @property
    def in_distribution_data_uncertainty(self):
        return self._in._entropy.cpu()

    @property
    def out_distribution_data_uncertainty(self):
        return self._out._entropy.cpu()

    @property
    def last_engine(self):
        return self._in.last_engine

    def update_statistics(self, models, output_performance=True):
        model_list = BMAAccumulator.as_model_list(models)
        self.num_samples_collected += len(model_list)
        self._in.accumulate(model_list)
        self._out.accumulate(model_list)
        if output_performance:
            return self.get_performance_metrics()

    def get_performance_metrics(self):
        from sklearn.metrics import roc_auc_score
        S = self.num_samples_collected
        self.in_distribution_total_uncertainty = util.compute_predictive_entropy(self.in_distribution_ensemble_proba / S)
        self.out_distribution_total_uncertainty = util.compute_predictive_entropy(self.out_distribution_ensemble_proba / S)
        self.in_distribution_model_uncertainty = self.in_distribution_total_uncertainty - \
            self.in_distribution_data_uncertainty / S
        self.out_distribution_model_uncertainty = self.out_distribution_total_uncertainty - \
            self.out_distribution_data_uncertainty / S
        if S == 1:
            # one sample: total and data uncertainty are the SAME quantity, the reference subtracts two identical tensors
            # and gets exact zeros (AUROC 0.5); here they come from two kernels and differ by rounding noise -- keep the
            # exact answer instead of ranking that noise
            self.in_distribution_model_uncertainty = torch.zeros_like(self.in_distribution_total_uncertainty)
            self.out_distribution_model_uncertainty = torch.zeros_like(self.out_distribution_total_uncertainty)
        label_array = np.concatenate([np.ones(self._out._n), np.zeros(self._in._n)])
        total = np.concatenate([self.out_distribution_total_uncertainty.numpy(), self.in_distribution_total_uncertainty.numpy()])
        model = np.concatenate([self.out_distribution_model_uncertainty.numpy(), self.in_distribution_model_uncertainty.numpy()])
        return {"total_uncertainty_auroc": roc_auc_score(label_array, total),
                "model_uncertainty_auroc": roc_auc_score(label_array, model)}
