"""Base class of the inference wrappers -- the drop-in boundary (reference inference/inference_base.py:12-56)."""
import torch

from ..util import get_loss_criterion


class _Inference:
    """Same constructor / method surface as the reference ``_Inference``."""

    def __init__(self, hyperparameters, model=None, train_loader=None, device=torch.device("cpu"),
                 model_loss="multi_class_linear_output"):
        self.model = model
        self.hyperparameters = hyperparameters
        self.train_loader = train_loader
        self.device = device
        self.loss_criterion = get_loss_criterion(loss=model_loss)

    def update_hyp(self, hyperparameters):
        raise NotImplementedError

    def sample_iterative(self):
        raise NotImplementedError

    def sample(self):
        raise NotImplementedError

    def compute_val_loss(self, val_loader=None):
        """Mean loss over ``val_loader`` (reference :46-56); the loss is accumulated on the device and read once."""
        with torch.no_grad():
            n_seen = 0
            total = None
            self.model.eval()
            for batch_data, batch_labels in val_loader:
                logits = self.model(batch_data.to(self.device))
                loss = self.loss_criterion(logits, batch_labels.to(self.device)) * len(batch_data)
                total = loss if total is None else total + loss
                n_seen += len(batch_data)
            return float(total.item()) / n_seen


def require_cuda(device, who):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("%s: device must be a CUDA device -- ursabench_b200 has no CPU path (got %r). "
                           "The reference's CPU implementation lives in oracle/ for testing only." % (who, device))
    return device
