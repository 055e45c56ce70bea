"""Host-side helpers the inference / task classes need (the hot-path subset of reference util.py).

Same names and argument meaning as the reference; no hamiltorch import (reference util.py:11 imports it
unconditionally although only the HMC wrapper uses it).
"""
import itertools
import json
import os.path
import random
import sys
import time
from io import StringIO

import numpy as np
import torch
from torch.nn import CrossEntropyLoss


def set_random_seed(seed=None):
    """reference util.py:20-29 (torch.manual_seed also seeds the Philox stream of this engine's optimizers)."""
    if seed is None:
        seed = int((time.time() * 1e6) % 1e8)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    return seed


class _NullIO(StringIO):
    def write(self, txt):
        pass


def silent(fn):
    """Decorator that swallows stdout of ``fn`` (reference util.py:40-50)."""

    def silent_fn(*args, **kwargs):
        saved = sys.stdout
        sys.stdout = _NullIO()
        try:
            return fn(*args, **kwargs)
        finally:
            sys.stdout = saved

    return silent_fn


def list_to_dic(names, hyp_list):
    return {name: hyp_list[i] for i, name in enumerate(names)}


def get_loss_criterion(loss="multi_class_linear_output", **kwargs):
    if loss == "multi_class_linear_output":          # reference util.py:80-89
        return CrossEntropyLoss(**kwargs)
    raise NotImplementedError


def reset_model(model):
    """Re-initialise direct children that have ``reset_parameters`` (reference util.py:92-107; nested containers
    are skipped there as well -- SURVEY Q9).  Works in place, so flat-buffer views stay attached."""
    if not isinstance(model, torch.nn.Module):
        raise NotImplementedError
    for _, child in model.named_children():
        fn = getattr(child, "reset_parameters", None)
        if fn is not None:
            fn()
    first = next(model.parameters(), None)
    flat = getattr(first, "_ursa_flat", None) if first is not None else None
    if flat is not None:
        flat.reattach()
    return model


def central_smoothing(proba, gamma=1e-4):
    return (1 - gamma) * proba + gamma * 1 / (proba.shape[1])     # reference util.py:126-134


def compute_predictive_entropy(proba):
    return -(proba * torch.log(proba)).sum(dim=-1)                 # reference util.py:137-144


def json_open_from_file(parser, arg):
    if not os.path.exists(arg):
        parser.error("The file %s does not exist!" % arg)
    with open(arg, encoding="utf-8") as f:
        return json.loads(f.read())


def make_dic_json_format(dic):
    for key in dic.keys():
        if type(dic[key]) is torch.Tensor:
            dic[key] = float(dic[key])
    return dic


def flatten(lst):
    """List of tensors -> one contiguous vector (reference util.py:163-169)."""
    return torch.cat([t.contiguous().view(-1) for t in lst])


def set_weights(model, vector, device=None):
    offset = 0                                                     # reference util.py:172-176
    for param in model.parameters():
        n = param.numel()
        param.data.copy_(vector[offset:offset + n].view(param.size()).to(param.device if device is None else device))
        offset += n


def adjust_learning_rate(optimizer, lr):
    for group in optimizer.param_groups:
        group["lr"] = lr
    return lr


def unflatten_like(vector, like_tensor_list):
    out, i = [], 0
    for t in like_tensor_list:
        n = t.numel()
        out.append(vector[:, i:i + n].view(t.shape))
        i += n
    return out


def _is_bn(module):
    return isinstance(module, torch.nn.modules.batchnorm._BatchNorm)


def check_bn(model):
    return any(_is_bn(m) for m in model.modules())


def bn_update(loader, model, subset=None, device=None, **kwargs):
    """Re-estimate BatchNorm running statistics with one pass over ``loader`` (reference util.py:212-247, which
    hard-codes ``input.cuda()``; here the model's own device is used)."""
    if not check_bn(model):
        return
    device = device if device is not None else next(model.parameters()).device
    model.train()
    momenta = {}
    for m in model.modules():
        if _is_bn(m):
            m.running_mean.zero_()
            m.running_var.fill_(1.0)
            momenta[m] = m.momentum
    n = 0
    num_batches = len(loader)
    with torch.no_grad():
        if subset is not None:
            num_batches = int(num_batches * subset)
            loader = itertools.islice(loader, num_batches)
        for inp, _ in loader:
            inp = inp.to(device, non_blocking=True)
            b = inp.size(0)
            momentum = b / (n + b)
            for m in momenta:
                m.momentum = momentum
            model(inp, **kwargs)
            n += b
    for m, mom in momenta.items():
        m.momentum = mom


# ---- PCA-subspace / ESS helpers: the reference keeps them in util (util.py:260-354); here they live in ``ess.py`` -----------
from .ess import cross_entropy, elliptical_slice, log_pdf  # noqa: E402,F401


def reset_bn(module):
    """Zero mean / unit variance running statistics of one BatchNorm module (reference util.py:196-199); for ``model.apply``."""
    if _is_bn(module):
        module.running_mean = torch.zeros_like(module.running_mean)
        module.running_var = torch.ones_like(module.running_var)


def convert_sample_to_net(flat_tensor, model):
    """A deep copy of ``model`` whose parameters are the slices of the flat vector, in ``parameters()`` order (reference
    util.py:110-123, which goes through ``hamiltorch.util.unflatten``; ``HMC.sample`` here returns bank handles instead)."""
    import copy
    net = copy.deepcopy(model)
    flat_tensor = flat_tensor.reshape(-1)
    total = sum(p.numel() for p in net.parameters())
    if total != flat_tensor.numel():
        raise ValueError("flat tensor has %d elements, the model %d parameters" % (flat_tensor.numel(), total))
    offset = 0
    for param in net.parameters():
        param.data = flat_tensor[offset:offset + param.numel()].view(param.size()).to(param.device, param.dtype).clone()
        offset += param.numel()
    return net
