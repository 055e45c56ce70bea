"""SGLD = SGHMC with alpha forced to 1, i.e. momentum 0 (reference inference/sgld.py:8-35)."""
import torch

from ..util import reset_model
from .optim_sghmc import optimSGHMC
from .sghmc import SGHMC


class SGLD(SGHMC):
    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu")):
        if hyperparameters is None:
            hyperparameters = {"lr": 0.001, "prior_std": 10, "num_samples": 2, "alpha": 0.1, "burn_in_epochs": 10}
        hyperparameters["alpha"] = 1.0                       # mutates the caller's dict like the reference (:22)
        super().__init__(hyperparameters, model, train_loader, model_loss, device)

    def update_hyp(self, hyperparameters):
        """reference :25-35: rebuilds the optimizer but NOT the lr scheduler (it keeps driving the old optimizer's
        param_groups, so after ``update_hyp`` SGLD runs at a constant lr)."""
        self.lr = hyperparameters["lr"]
        self.prior_std = hyperparameters["prior_std"]
        self.num_samples = hyperparameters["num_samples"]
        self.alpha = 1.0
        self.burn_in_epochs = hyperparameters["burn_in_epochs"]
        self.temperature = hyperparameters.get("temperature", 1.0)
        self.model = reset_model(self.model)
        self.optimizer = optimSGHMC(params=self.model.parameters(), lr=self.lr, momentum=1 - self.alpha,
                                    num_training_samples=self.dataset_size, weight_decay=1 / (self.prior_std ** 2),
                                    temperature=self.temperature)
        self.burnt_in = False
        self.epochs_run = 0
        self.bank = self.bank.fresh()            # earlier sample() handles keep the old rows
        if hasattr(self, "disable_cuda_graph"):
            self.disable_cuda_graph()             # a captured step reads the OLD optimizer's device scalars
