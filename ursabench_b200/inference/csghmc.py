"""Cyclical SGHMC (reference inference/csghmc.py:13-127) on the fused K1 update."""
import numpy as np
import torch

from ..util import get_loss_criterion, reset_model
from ._loop import SGMCMCLoop
from .inference_base import _Inference
from .optim_sghmc import optimSGHMC


class cSGHMC(_Inference, SGMCMCLoop):
    """Hyperparameters: ``lr_0, prior_std, num_samples_per_cycle, cycle_length, burn_in_epochs, num_cycles, alpha``.
    The step size follows a per-ITERATION cosine inside each cycle, computed on the host in float64 exactly like the
    reference (:64-72, including its float ``num_batch`` over-count) and passed to K1 as an fp32 scalar."""

    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu")):
        if hyperparameters is None:
            hyperparameters = {"lr_0": 0.001, "prior_std": 10.1, "num_samples_per_cycle": 5, "cycle_length": 20,
                               "burn_in_epochs": 5, "num_cycles": 10, "alpha": 1.0}
        super().__init__(hyperparameters, model, train_loader, device)
        self._read_hyp(hyperparameters)
        self.batch_size = train_loader.batch_size
        self.dataloader_batch_size = train_loader.batch_size
        self.num_batch = max(1, len(train_loader.dataset) / self.batch_size + 1)      # float (:30-32)
        self.model_loss = model_loss
        self.dataset_size = len(train_loader.dataset)
        self._attach(model, train_loader, device, type(self).__name__)
        self.loss_criterion = get_loss_criterion(loss=model_loss)
        self._build_optimizer()
        self.total_epochs = self.cycle_length * self.num_cycles
        self.total_iterations = self.total_epochs * self.num_batch
        assert (self.cycle_length - self.burn_in_epochs - self.num_samples_per_cycle) > 0

    def _read_hyp(self, h):
        self.lr_0 = h["lr_0"]
        self.prior_std = h["prior_std"]
        self.num_samples_per_cycle = h["num_samples_per_cycle"]
        self.cycle_length = h["cycle_length"]
        self.alpha = h["alpha"]
        self.burn_in_epochs = h["burn_in_epochs"]
        self.num_cycles = h["num_cycles"]
        self.temperature = h.get("temperature", 1.0)     # optional, not a reference key: sqrt(T) on the noise

    def _build_optimizer(self):
        self.optimizer = optimSGHMC(params=self.model.parameters(), lr=self.lr_0, momentum=1 - self.alpha,
                                    num_training_samples=self.dataset_size, weight_decay=1 / (self.prior_std ** 2),
                                    temperature=self.temperature)
        self.burnt_in = False
        self.epochs_run = 0

    def update_hyp(self, hyperparameters):
        self._read_hyp(hyperparameters)
        self.model = reset_model(self.model)
        self._build_optimizer()
        self.bank = self.bank.fresh()            # earlier sample() handles keep the old rows
        if hasattr(self, "disable_cuda_graph"):
            self.disable_cuda_graph()             # a captured step reads the OLD optimizer's device scalars
        assert (self.cycle_length - self.burn_in_epochs - self.num_samples_per_cycle) > 0
        # NB the reference does not recompute total_iterations here either (:48-62)

    def _adjust_learning_rate(self, optimizer, epoch, batch_idx):
        rcounter = epoch * self.num_batch + batch_idx
        per_cycle = self.total_iterations // self.num_cycles             # float floor-division
        cos_inner = np.pi * (rcounter % per_cycle)
        cos_inner /= per_cycle
        lr = 0.5 * (np.cos(cos_inner) + 1) * self.lr_0
        for group in optimizer.param_groups:
            group["lr"] = lr
        return lr

    def _noise_gate(self):
        return (self.epochs_run % self.cycle_length) + 1 > (self.cycle_length - self.burn_in_epochs
                                                            - self.num_samples_per_cycle)      # :89-90

    def _sample_gate_after(self, epochs_run):
        return ((epochs_run - 1) % self.cycle_length) >= (self.cycle_length - self.num_samples_per_cycle)  # :106

    def sample_iterative(self, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if not isinstance(self.model, torch.nn.Module):
            raise NotImplementedError
        while True:
            noisy = self._noise_gate()
            will_sample = self._sample_gate_after(self.epochs_run + 1)

            def set_lr(batch_idx):
                self.lr = self._adjust_learning_rate(self.optimizer, self.epochs_run, batch_idx)

            row = self._run_epoch(lambda b: noisy, lr_for_batch=set_lr, snapshot_last=will_sample,
                                  track_loss=debug_val_loss)
            self.epochs_run += 1
            print("Epoch: ", self.epochs_run, " lr: ", self.lr)
            if debug_val_loss:
                metrics = {"train_loss": float(self._epoch_loss.item()) / self.dataset_size,
                           "val_loss": self.compute_val_loss(val_loader)}
                print(metrics)
                if wandb_debug:
                    import wandb
                    wandb.log(metrics)
            if will_sample:
                return self.bank.handle(row)

    def sample(self, num_samples=None, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if num_samples is None:
            num_samples = self.num_samples_per_cycle * self.num_cycles
        if not isinstance(self.model, torch.nn.Module):
            raise NotImplementedError
        return [self.sample_iterative(val_loader=val_loader, debug_val_loss=debug_val_loss, wandb_debug=wandb_debug)
                for _ in range(num_samples)]
