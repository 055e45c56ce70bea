"""SWA: SGD iterates + running first / second moments + deviation ring (reference inference/swa.py:13-178).

The moments live on the device as flat fp32 vectors and are updated by ONE K2a launch per collect (24 B/param),
instead of T device->host copies plus seven CPU passes and a K x D ``torch.cat`` (reference :79-90).
"""
from copy import deepcopy

import torch
from torch.optim import SGD

from .. import _C
from ..bank import SampleBank
from ..flat import FlatParams
from ..util import adjust_learning_rate, bn_update, get_loss_criterion, reset_model
from .inference_base import _Inference, require_cuda
from .subspaces import Subspace


class SWA(_Inference):
    _defaults = {"swag_lr": 0.001, "swag_wd": 0.001, "lr_init": 0.001, "num_samples": 20, "momentum": 0.1,
                 "burn_in_epochs": 100, "num_iterates": 50}

    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu"), **subspace_kwargs):
        super().__init__(hyperparameters, model=None, train_loader=None, device=torch.device("cpu"))
        if hyperparameters is None:
            hyperparameters = dict(self._defaults)
        if not isinstance(model, torch.nn.Module):
            raise NotImplementedError
        _C.lib()
        self.device = require_cuda(device, type(self).__name__)
        self.hyperparameters = hyperparameters
        if next(model.parameters()).device != self.device:
            model.to(self.device)
        self._skeleton = deepcopy(model).cpu()
        self.swag_model = deepcopy(model)
        self.model = model
        self.flat = FlatParams.from_model(self.model, self.device)
        self.swag_flat = FlatParams.from_model(self.swag_model, self.device)
        self.num_parameters = self.flat.D
        self.var_clamp = 1e-30
        self.train_loader = train_loader
        self.loss_criterion = get_loss_criterion(loss=model_loss)
        self.dataset_size = len(train_loader.dataset)
        self.cov_factor = None
        self._subspace_kwargs = dict(subspace_kwargs or {})
        self.bank = SampleBank(self.flat.D, self.flat.nb, self.device, capacity=8, skeleton=self._skeleton)
        self._reset_state(hyperparameters)

    # -- state ------------------------------------------------------------------------------------------------
    def _reset_state(self, hyperparameters, **subspace_kwargs):
        ld = self.flat.ld
        self._mean = torch.zeros(ld, dtype=torch.float32, device=self.device)
        self._sq = torch.zeros(ld, dtype=torch.float32, device=self.device)
        self.num_models_collected = torch.zeros(1, dtype=torch.long)
        self.burnt_in = False
        self.epochs_run = 0
        self.hyperparameters = hyperparameters
        self.burn_in_epochs = hyperparameters["burn_in_epochs"]
        self.num_iterates = hyperparameters["num_iterates"]
        self.momentum = hyperparameters["momentum"]
        self.lr_init = hyperparameters["lr_init"]
        self.swag_lr = hyperparameters["swag_lr"]
        self.swag_wd = hyperparameters["swag_wd"]
        self.optimizer = SGD(params=self.model.parameters(), lr=self.lr_init, momentum=self.momentum,
                             weight_decay=self.swag_wd)
        self.subspace_type = hyperparameters.get("subspace_type", "pca")           # reference :43-46
        kw = dict(self._subspace_kwargs)
        kw.update(subspace_kwargs)
        self.subspace = Subspace.create(self.subspace_type, num_parameters=self.num_parameters, device=self.device,
                                        **kw)
        self.bank = self.bank.fresh()            # earlier sample() handles keep the old rows
        if hasattr(self, "disable_cuda_graph"):
            self.disable_cuda_graph()             # a captured step reads the OLD optimizer's device scalars

    @property
    def weight_mean(self):
        return self._mean[:self.num_parameters]

    @property
    def sq_mean(self):
        return self._sq[:self.num_parameters]

    def update_hyp(self, hyperparameters, **subspace_kwargs):
        self.model = reset_model(self.model)
        self.swag_model = reset_model(self.swag_model)
        self._reset_state(hyperparameters, **subspace_kwargs)

    # -- hot path ---------------------------------------------------------------------------------------------
    def _collect_model(self):
        """K2a: moments + deviation row in one pass (reference :79-90).  ``n`` is whatever
        ``num_models_collected`` holds -- the caller decides when it increments."""
        n = int(self.num_models_collected.item())
        _C.swag_collect(self.flat.p, self._mean, self._sq, self.subspace.next_slot(), n)
        self.subspace.commit()

    def _schedule(self, epoch):
        t = epoch / self.burn_in_epochs                      # reference :92-101
        lr_ratio = self.swag_lr / self.lr_init
        if t <= 0.5:
            factor = 1.0
        elif t <= 0.9:
            factor = 1.0 - (1.0 - lr_ratio) * (t - 0.5) / 0.4
        else:
            factor = lr_ratio
        return self.lr_init * factor

    def _set_swa(self):
        self.swag_flat.load_vector(self._mean)

    def _get_mean_and_variance(self):
        var = torch.empty_like(self._mean)
        _C.swag_variance(self._mean, self._sq, var, self.var_clamp)
        return self.weight_mean, var[:self.num_parameters]

    def fit(self):
        if self.cov_factor is None:
            self.cov_factor = self.subspace.get_space()

    def get_space(self, export_cov_factor=True):
        mean, variance = self._get_mean_and_variance()
        if not export_cov_factor:
            return mean.clone(), variance.clone()
        self.fit()
        return mean.clone(), variance.clone(), self.cov_factor.clone()

    def _sgd_epoch(self, track_loss=False):
        self.model.train()
        lr = self._schedule(self.epochs_run)
        adjust_learning_rate(self.optimizer, lr)
        total = torch.zeros((), device=self.device) if track_loss else None
        for batch_data, batch_labels in self.train_loader:
            batch_data = batch_data.to(self.device, non_blocking=True)
            batch_labels = batch_labels.to(self.device, non_blocking=True)
            loss = self.loss_criterion(self.model(batch_data), batch_labels)
            self.optimizer.zero_grad()
            loss.backward()
            if track_loss:
                total += loss.detach() * len(batch_data)
            self.optimizer.step()
        self.epochs_run += 1
        return total

    def _debug(self, total, val_loader, wandb_debug):
        metrics = {"train_loss": float(total.item()) / self.dataset_size, "val_loss": self.compute_val_loss(val_loader)}
        print(metrics)
        if wandb_debug:
            import wandb
            wandb.log(metrics)

    def sample_iterative(self, update_bn_swa=True, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if self.burnt_in is False:
            epochs = self.burn_in_epochs + 1
            self.burnt_in = True
        else:
            epochs = 1
        self.num_models_collected += 1                       # reference :130: BEFORE collecting (biased mean, Q8)
        for _ in range(epochs):
            total = self._sgd_epoch(track_loss=debug_val_loss)
            if debug_val_loss:
                self._debug(total, val_loader, wandb_debug)
        self._collect_model()
        if update_bn_swa:
            self._set_swa()
            bn_update(self.train_loader, self.swag_model, device=self.device)
        return self.swag_model

    def sample(self, num_samples=None, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if num_samples is None:
            num_samples = self.num_iterates
        return [self.sample_iterative(update_bn_swa=(i == num_samples - 1), val_loader=val_loader,
                                      debug_val_loss=debug_val_loss, wandb_debug=wandb_debug)
                for i in range(num_samples)]
