"""Test infrastructure (see oracle/__init__.py): deterministic weights for WideResNet fixtures.

A WRN bank row is too large to commit (WRN-10-2: 0.3 M floats per sample), so the golden fixture
``tests/golden/prediction_wrn.npz`` stores only inputs and the reference's outputs; both the generator (on the live
reference model, models/wideresnet.py:78-120) and the tests (on ``ursabench_b200.models.WideResNet``) fill the module
from the same ``numpy.random.RandomState`` stream in ``named_parameters()`` / ``named_buffers()`` order -- which is the
flat layout itself (util.flatten, util.py:163-169), so a layout mismatch shows up as a parity failure.
"""
import numpy as np
import torch


def wrn_fill(model, seed, logit_gain=4.0):
    rng = np.random.RandomState(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            shape = tuple(p.shape)
            if p.dim() == 4:                                   # conv filters: He-style scale keeps activations O(1)
                v = rng.randn(*shape) * np.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
            elif p.dim() == 2:                                 # classifier: sharper logits
                v = rng.randn(*shape) * (logit_gain / np.sqrt(shape[1]))
            elif name.endswith("weight"):                      # BatchNorm gamma
                v = rng.uniform(0.5, 1.5, size=shape)
            else:                                              # conv / BatchNorm / linear bias
                v = rng.randn(*shape) * 0.1
            p.copy_(torch.from_numpy(v.astype(np.float32)))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.from_numpy((rng.randn(*b.shape) * 0.3).astype(np.float32)))
            elif name.endswith("running_var"):
                b.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, size=tuple(b.shape)).astype(np.float32)))
    return model
