"""``PCASubspaceSampler``: elliptical slice sampling in the PCA subspace of the SWA iterates
(reference inference/pca_subspace.py:13-160; ``util.elliptical_slice`` util.py:287-354, ``util.log_pdf`` :260-274).

What the reference does per ESS proposal: project ``t`` to weights on the HOST (dense ``[D, r] @ [r]``), copy them tensor by
tensor into the model, then re-iterate the training DataLoader for a full train-mode forward.  Here the subspace lives on
the device (``PCASpace`` ring -> Gram -> eigh -> K2b components, ``SubspaceModel`` = one K2b pass), the projection is
written straight into the model's flat parameter buffer (the parameters are views of it), and the training set is
uploaded once and stays resident -- a proposal costs one K2b launch plus the forward passes, with no host<->device
traffic but the final scalar.  The slice-sampling logic itself (bracket, shrinkage, numpy RNG call order) follows
``util.elliptical_slice`` statement by statement so that, given the same ``np.random`` state and log-density values, the
chain is the reference's chain (tests/golden/ess.npz).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .. import _C
from ..bank import SampleBank
from ..flat import FlatParams
from ..ess import elliptical_slice
from ..util import bn_update, check_bn, get_loss_criterion, reset_model
from .inference_base import _Inference, require_cuda
from .projection_model import SubspaceModel
from .swa import SWA

__all__ = ["PCASubspaceSampler", "elliptical_slice"]


class PCASubspaceSampler(_Inference):
    """Hyperparameters (reference :20-23): ``swag_lr, swag_wd, lr_init, num_samples, swag_momentum, swag_burn_in_epochs,
    num_swag_iterates, rank, max_rank, temperature, prior_std``."""

    _defaults = {"swag_lr": 0.001, "swag_wd": 0.001, "lr_init": 0.001, "num_samples": 20, "swag_momentum": 0.1,
                 "swag_burn_in_epochs": 100, "num_swag_iterates": 50, "rank": 20, "max_rank": 20, "temperature": 5000,
                 "prior_std": 2.}

    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu")):
        super().__init__(hyperparameters=hyperparameters, model=model, train_loader=train_loader, device=device)
        self.hyperparameters = dict(self._defaults) if hyperparameters is None else hyperparameters
        _C.lib()
        self.device = require_cuda(device, type(self).__name__)
        self.model = model
        self.train_loader = train_loader
        self.model_loss_type = model_loss
        self.loss_criterion = get_loss_criterion(loss=model_loss)
        self._read_hyp(self.hyperparameters)
        self.swag_model = SWA(hyperparameters=self._swa_hyp(), model=self.model, train_loader=self.train_loader,
                              model_loss=self.model_loss_type, device=self.device, max_rank=self.max_rank, pca_rank=self.rank)
        self.flat = self.swag_model.flat                       # the model's parameters are views of this flat buffer
        self._clear()
        self.bank = SampleBank(self.flat.D, self.flat.nb, self.device, capacity=8, skeleton=self.swag_model._skeleton)
        self._train_x = self._train_y = None
        self.lnpdf_evaluations = 0

    def _read_hyp(self, h):
        self.rank = h["rank"]
        self.max_rank = h["max_rank"]
        self.lr_init = h["lr_init"]
        self.swag_lr = h["swag_lr"]
        self.swag_wd = h["swag_wd"]
        self.swag_burn_in_epochs = h["swag_burn_in_epochs"]
        self.num_samples = h["num_samples"]
        self.num_swag_iterates = h["num_swag_iterates"]
        self.swag_momentum = h["swag_momentum"]
        self.prior_std = h["prior_std"]
        self.temperature = h["temperature"]

    def _swa_hyp(self):
        return {"burn_in_epochs": self.swag_burn_in_epochs, "momentum": self.swag_momentum, "lr_init": self.lr_init,
                "swag_lr": self.swag_lr, "swag_wd": self.swag_wd, "num_iterates": self.num_swag_iterates,
                "subspace_type": "pca"}

    def _clear(self):
        self.subspace_constructed = False
        self.current_theta = None
        self.weight_mean = None
        self.weight_covariance = None
        self.subspace = None

    def update_hyp(self, hyperparameters):
        keep_temperature = self.hyperparameters["temperature"]          # reference :82 re-reads the OLD dict
        self._read_hyp(hyperparameters)
        self.temperature = keep_temperature
        self.model = reset_model(self.model)
        self.swag_model.update_hyp(self._swa_hyp(), max_rank=self.max_rank, pca_rank=self.rank)
        self._clear()
        self.bank = self.bank.fresh()

    # -- log density of a subspace point: full-dataset train-mode forward (reference util.py:260-274) ----------------
    def _resident_train_set(self):
        if self._train_x is None:
            xs, ys = [], []
            for xb, yb in self.train_loader:
                xs.append(xb)
                ys.append(yb)
            self._train_sizes = [len(xb) for xb in xs]
            self._train_x = torch.cat(xs).to(self.device, non_blocking=True)
            self._train_y = torch.cat(ys).to(self.device, non_blocking=True)
        return self._train_x, self._train_y

    def _project_into_model(self, theta):
        """theta [rank] (numpy / tensor) -> the model's flat parameter buffer, in one K2b pass."""
        t = torch.as_tensor(np.asarray(theta, dtype=np.float32)) if not torch.is_tensor(theta) else theta.float()
        w = self.subspace(t.to(self.device))
        self.flat.p[:self.flat.D].copy_(w)
        return w

    def _oracle(self, theta, subspace=None):
        self._project_into_model(theta)
        x, y = self._resident_train_set()
        self.model.train()                                     # the reference evaluates in train mode (:266)
        loss = torch.zeros((), device=self.device)
        with torch.no_grad():
            off = 0
            for b in self._train_sizes:                        # the loader's own batches: BatchNorm batch statistics
                out = self.model(x[off:off + b])
                loss += F.cross_entropy(out, y[off:off + b]) * b
                off += b
        self.lnpdf_evaluations += 1
        return -loss.item() / self.temperature

    # -- sampling --------------------------------------------------------------------------------------------------------
    def sample_iterative(self, update_bn=True, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if self.subspace_constructed is False:
            self.swag_model.sample(val_loader=val_loader, debug_val_loss=debug_val_loss, wandb_debug=wandb_debug)
            self.subspace_constructed = True
        if self.weight_mean is None or self.weight_covariance is None:
            self.weight_mean, _, self.weight_covariance = self.swag_model.get_space()
        if self.subspace is None:
            self.subspace = SubspaceModel(self.weight_mean, self.weight_covariance)
            self.rank = self.subspace.rank                     # fewer iterates than `rank` shrink the space (subspaces.py:128)
        if self.current_theta is None:
            self.current_theta = torch.zeros(self.rank)
        prior_sample = np.random.normal(loc=0.0, scale=self.prior_std, size=self.rank)
        theta, log_prob = elliptical_slice(initial_theta=self.current_theta.numpy().copy(), prior=prior_sample,
                                           lnpdf=self._oracle, subspace=self.subspace)
        self.current_theta = torch.FloatTensor(theta)
        self.last_log_prob = log_prob
        self._project_into_model(self.current_theta)
        if debug_val_loss:
            metrics = {"val_loss": self.compute_val_loss(val_loader)}
            print(metrics)
            if wandb_debug:
                import wandb
                wandb.log(metrics)
        if update_bn and check_bn(self.model):
            bn_update(self.train_loader, self.model, device=self.device)
        row = self.bank.append(self.flat.p, self.flat.b if self.flat.nb else None)
        return self.bank.handle(row)

    def sample(self, num_samples=None, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if num_samples is None:
            num_samples = self.num_samples
        return [self.sample_iterative(update_bn=(i == num_samples - 1), val_loader=val_loader,
                                      debug_val_loss=debug_val_loss, wandb_debug=wandb_debug) for i in range(num_samples)]
