"""``Decision``: cost-sensitive decisions from BMA probabilities (reference tasks/decision_making.py:83-152).

Reference loop (:118-142): the Prediction loop accumulating the smoothed probabilities p~_s and the risk
``p~_s @ cost_mat`` per sample on the CPU.  The risk is linear in p~, so  sum_s p~_s @ cost = (sum_s p~_s) @ cost: the
engine's accumulator (``BMAAccumulator``: resident inputs, sample-batched fused forward) is mapped through the affine
smoothing and ONE [N, C] x [C, C] product on the device.  Outputs as in the reference (:144-152): ``Decision`` = argmin of
the averaged risk, ``True_Cost`` = sum of cost_mat[target, decision], ``Pred_cost`` = the accumulated risk.

The cost matrix is chosen by the dataset class exactly like the reference (:95-102, MNIST / CIFAR10 / CIFAR100, anything
else raises NotImplementedError); ``cost_mat=`` is an extension for other datasets.
"""
import torch

from ._engine import BMAAccumulator
from .task_base import _Task

__all__ = ["Decision", "MNIST_cost", "CIFAR10_cost", "CIFAR100_cost", "decision_cost"]

_GAMMA = 1e-4


def _cost(num_classes, rows, high):
    """0 on the diagonal, 0.1 elsewhere, `high` in the rows of the classes whose misclassification is expensive."""
    eye = torch.eye(num_classes)
    c = torch.full((num_classes, num_classes), 0.1)
    c[rows] = high
    c[eye == 1] = 0
    return c


def MNIST_cost(num_classes):                   # decision_making.py:12-19: digits 3 and 7, cost 100
    return _cost(num_classes, [3, 7], 100.0)


def CIFAR10_cost(num_classes):                 # :21-28: plane, automobile, ship, truck
    return _cost(num_classes, [0, 1, 8, 9], 1.0)


def CIFAR100_cost(num_classes):                # :39-51: 'tank', 'rocket', 'pickup_truck' = fine labels 85, 69, 58 (:30-37)
    return _cost(num_classes, [58, 69, 85], 1.0)


def decision_cost(D, y_true, cost_mat=None):   # :69-77
    return cost_mat[y_true, D].sum()


class Decision(_Task):
    def __init__(self, dataloader, num_classes, device, cost_mat=None, engine="auto"):
        super().__init__(dataloader, num_classes, device)
        self.data_loader = dataloader["decision_data_test"]
        self.num_classes = num_classes
        self.device = torch.device(device)
        if cost_mat is None:
            name = type(self.data_loader.dataset).__name__
            module = type(self.data_loader.dataset).__module__
            if not module.startswith("torchvision.datasets") or name not in ("MNIST", "CIFAR10", "CIFAR100"):
                raise NotImplementedError
            cost_mat = {"MNIST": MNIST_cost, "CIFAR10": CIFAR10_cost, "CIFAR100": CIFAR100_cost}[name](num_classes)
        self.cost_mat = cost_mat.float()
        self._acc = BMAAccumulator(self.data_loader, num_classes, device, engine)
        self.targets = self._acc.targets
        self._cost_dev = self.cost_mat.to(self.device)
        self.num_samples_collected = 0

    def reset(self):
        self.num_samples_collected = 0
        self._acc._proba.zero_()
        self._acc._entropy.zero_()

    def _smoothed_sum_dev(self):
        return (1.0 - _GAMMA) * self._acc._proba + self.num_samples_collected * _GAMMA / self.num_classes

    @property
    def ensemble_proba(self):
        return self._smoothed_sum_dev().cpu()

    @property
    def risk(self):
        return (self._smoothed_sum_dev() @ self._cost_dev).cpu()

    @property
    def last_engine(self):
        return self._acc.last_engine

    def update_statistics(self, models, output_performance=True, smoothing=True):
        model_list = BMAAccumulator.as_model_list(models)
        self.num_samples_collected += len(model_list)
        self._acc.accumulate(model_list)
        if output_performance:
            return self.get_performance_metrics(output_performance, smoothing)

    def get_performance_metrics(self, output_performance=False, smoothing=True):
        risk = self.risk
        D = (risk / self.num_samples_collected).argmin(1)
        return {"True_Cost": decision_cost(D, self.targets, self.cost_mat), "Decision": D, "Pred_cost": risk}
