"""cSGLD = cSGHMC with alpha forced to 1 (reference inference/csgld.py:8-36)."""
import torch

from .csghmc import cSGHMC


class cSGLD(cSGHMC):
    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu")):
        if hyperparameters is None:
            hyperparameters = {"lr_0": 0.001, "prior_std": 10.1, "num_samples_per_cycle": 5, "cycle_length": 20,
                               "burn_in_epochs": 5, "num_cycles": 10, "alpha": 1.0}
        hyperparameters["alpha"] = 1.0                       # mutates the caller's dict like the reference (:21)
        super().__init__(hyperparameters, model, train_loader, model_loss, device)

    def update_hyp(self, hyperparameters):
        hyperparameters = dict(hyperparameters)
        hyperparameters["alpha"] = 1.0                       # reference :29 sets self.alpha = 1. without touching the dict
        super().update_hyp(hyperparameters)
