"""SGHMC sampler (reference inference/sghmc.py:12-115) on the fused K1 update."""
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from ..util import get_loss_criterion, reset_model
from ._loop import SGMCMCLoop
from .inference_base import _Inference
from .optim_sghmc import optimSGHMC


class SGHMC(_Inference, SGMCMCLoop):
    """Hyperparameters: ``lr, prior_std, num_samples, alpha, burn_in_epochs`` (reference :17-25).
    ``sample()`` returns ``num_samples`` module handles (one per epoch after burn-in) backed by the device bank."""

    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu")):
        if hyperparameters is None:
            hyperparameters = {"lr": 0.001, "prior_std": 10, "num_samples": 2, "alpha": 0.1, "burn_in_epochs": 10}
        super().__init__(hyperparameters, model, train_loader, device)
        self._read_hyp(hyperparameters)
        self.model_loss = model_loss
        self.dataset_size = len(train_loader.dataset)
        self._attach(model, train_loader, device, type(self).__name__)
        self.loss_criterion = get_loss_criterion(loss=model_loss)
        self._build_optimizer(eta_min=0)                     # reference :44-45 (eta_min defaults to 0 in the ctor)

    def _read_hyp(self, h):
        self.lr = h["lr"]
        self.prior_std = h["prior_std"]
        self.num_samples = h["num_samples"]
        self.alpha = h["alpha"]
        self.burn_in_epochs = h["burn_in_epochs"]
        self.temperature = h.get("temperature", 1.0)     # optional, not a reference key: sqrt(T) on the noise

    def _build_optimizer(self, eta_min):
        self.optimizer = optimSGHMC(params=self.model.parameters(), lr=self.lr, momentum=1 - self.alpha,
                                    num_training_samples=self.dataset_size, weight_decay=1 / (self.prior_std ** 2),
                                    temperature=self.temperature)
        self.burnt_in = False
        self.epochs_run = 0
        self.lr_final = self.lr / 2
        self.optimizer_scheduler = CosineAnnealingLR(optimizer=self.optimizer,
                                                     T_max=self.burn_in_epochs + self.num_samples, eta_min=eta_min)

    def update_hyp(self, hyperparameters):
        self._read_hyp(hyperparameters)
        self.model = reset_model(self.model)
        self._build_optimizer(eta_min=self.lr / 2)           # reference :62-63
        self.bank = self.bank.fresh()            # earlier sample() handles keep the old rows
        if hasattr(self, "disable_cuda_graph"):
            self.disable_cuda_graph()             # a captured step reads the OLD optimizer's device scalars

    def sample_iterative(self, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if not isinstance(self.model, torch.nn.Module):
            raise NotImplementedError
        if self.burnt_in is False:
            epochs = self.burn_in_epochs + 1
            self.burnt_in = True                             # set before the loop: noise is on from step 0 (Q2)
        else:
            epochs = 1
        row = None
        for epoch in range(epochs):
            noisy = epoch > 0.8 * epochs or self.burnt_in    # reference :83
            row = self._run_epoch(lambda b: noisy, snapshot_last=(epoch == epochs - 1), track_loss=debug_val_loss)
            self.optimizer_scheduler.step()
            if debug_val_loss:
                metrics = {"train_loss": float(self._epoch_loss.item()) / self.dataset_size,
                           "val_loss": self.compute_val_loss(val_loader),
                           "lr": self.optimizer_scheduler.get_last_lr()}
                print(metrics)
                if wandb_debug:
                    import wandb
                    wandb.log(metrics)
        return self.bank.handle(row)

    def sample(self, num_samples=None, val_loader=None, debug_val_loss=False, wandb_debug=False):
        if num_samples is None:
            num_samples = self.num_samples
        if not isinstance(self.model, torch.nn.Module):
            raise NotImplementedError
        return [self.sample_iterative(val_loader=val_loader, debug_val_loss=debug_val_loss, wandb_debug=wandb_debug)
                for _ in range(num_samples)]
