"""torch-CPU port of the reference's hot path, issuing the same op sequence the reference issues
(per-parameter-tensor Python loops, ``randn_like`` per tensor, per-batch per-sample forwards with CPU accumulation).

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): ``bench.py --impl reference`` and the
``cpu_baseline`` leg time this on the box's host cores, and the GPU tests use it for trajectory checks.  It is pinned
to the live reference by tests/test_oracle_golden.py::test_port_* (same inputs -> same outputs as the fixtures).
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F


class PortOptimSGHMC:
    """Op-for-op port of optimSGHMC.step (reference inference/optim_sghmc.py:30-68)."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, num_training_samples=None):
        self.params = list(params)
        self.lr, self.momentum, self.weight_decay, self.n = lr, momentum, weight_decay, num_training_samples
        self.buf = {}

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, add_langevin_noise=True, noise=None):
        """noise: optional list of per-tensor N(0,1) tensors replacing randn_like (identical-noise protocol)."""
        for i, p in enumerate(self.params):
            if p.grad is None:                                                # :44-45
                continue
            d = p.grad
            if self.weight_decay != 0:
                d = d.add(p, alpha=self.weight_decay / self.n)                # :47-48
            if self.momentum != 0:
                if i not in self.buf:
                    self.buf[i] = torch.clone(d).detach()                     # :51-52
                b = self.buf[i]
                b.mul_(self.momentum).add_(d, alpha=-self.lr)                 # :53 / :56
                d = b                                                         # :60
            else:
                d = d.mul(-self.lr)                                           # :62
            if add_langevin_noise:
                z = torch.randn_like(d) if noise is None else noise[i].view_as(d)
                d = d.add(z * math.sqrt(2 * (1 - self.momentum) * self.lr) / self.n)   # :63-64
            p.add_(d)                                                         # :65
            if self.momentum != 0:
                self.buf[i] = d                                               # :66-67


def port_csghmc_lr(lr_0, epoch, batch_idx, num_batch, cycle_length, num_cycles):
    """reference inference/csghmc.py:64-72."""
    total_iterations = cycle_length * num_cycles * num_batch
    per_cycle = total_iterations // num_cycles
    inner = np.pi * ((epoch * num_batch + batch_idx) % per_cycle)
    inner /= per_cycle
    return 0.5 * (np.cos(inner) + 1) * lr_0


def port_sgmcmc_epochs(model, batches, opt, n_epochs, lr_fn=None, noise_fn=lambda e, b: True, noise_for=None):
    """The sampler inner loop (reference sghmc.py:72-86 / csghmc.py:80-93) on CPU: fwd, zero_grad, loss, backward,
    ``loss.item()``, step.  ``batches``: list of (x, y) CPU tensors.  Returns the flat weights after the last step."""
    crit = torch.nn.CrossEntropyLoss()
    for e in range(n_epochs):
        model.train()
        for b, (x, y) in enumerate(batches):
            if lr_fn is not None:
                opt.lr = lr_fn(e, b)
            logits = model(x)
            opt.zero_grad()
            loss = crit(logits, y)
            loss.backward()
            loss.item()
            opt.step(add_langevin_noise=noise_fn(e, b), noise=None if noise_for is None else noise_for(e, b))
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def port_swa_collect(w, mean, sq_mean, n):
    """reference inference/swa.py:79-90 on CPU tensors (in place); returns the deviation vector."""
    mean.mul_(n / (n + 1.0))
    mean.add_(w / (n + 1.0))
    sq_mean.mul_(n / (n + 1.0))
    sq_mean.add_(w ** 2 / (n + 1.0))
    return w - mean


def port_swag_draw(mean, var, ring, max_rank, n_draws, full_cov=True):
    """The draw the reference intends (inference/swag.py:85-97), one sample at a time like its sample() loop."""
    out = []
    for _ in range(n_draws):
        if not full_cov:
            out.append(torch.normal(mean, torch.sqrt(var)))
            continue
        var_sample = var.sqrt() * torch.randn_like(var)
        cov_sample = ring.t().matmul(ring.new_empty((ring.size(0),)).normal_())
        cov_sample /= (max_rank - 1) ** 0.5
        out.append(mean + var_sample + cov_sample)
    return torch.stack(out)


def port_prediction_update(models, batches, num_classes, gamma=1e-4):
    """reference tasks/prediction.py:52-75 on CPU: batches outer, samples inner, softmax twice, CPU accumulate."""
    n = sum(len(x) for x, _ in batches)
    proba = torch.zeros(n, num_classes)
    ent = torch.zeros(n)
    with torch.no_grad():
        start = 0
        for x, _ in batches:
            end = start + len(x)
            for m in models:
                m.eval()
                logits = m(x)
                proba[start:end] += F.log_softmax(logits, dim=-1).exp_()
                p = F.log_softmax(logits, dim=-1).exp_()
                q = (1 - gamma) * p + gamma * 1 / p.shape[1]
                ent[start:end] += -(q * torch.log(q)).sum(dim=-1)
            start = end
    return proba, ent


def port_metrics(proba_sum, num_samples, targets, gamma=1e-4, n_bins=15):
    """reference tasks/prediction.py:79-102,152-194 (error_rate, nll, brier_score, ece)."""
    pbar_t = proba_sum / num_samples
    pbar = pbar_t.numpy()
    y = targets.numpy()
    out = {"error_rate": 1 - np.mean(np.argmax(pbar, axis=1) == y)}
    out["nll"] = F.nll_loss(torch.log((1 - gamma) * pbar_t + gamma * 1 / pbar_t.shape[1]), targets).item()
    onehot = np.zeros(pbar.shape)
    onehot[np.arange(len(y)), y] = 1.0
    out["brier_score"] = np.mean(np.sum((pbar - onehot) ** 2, axis=1))
    bounds = np.linspace(0, 1, n_bins + 1)
    conf, pred = np.max(pbar, 1), np.argmax(pbar, 1)
    acc = pred == y
    ece = 0.0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        inb = np.logical_and(conf > lo, conf <= hi)
        prop = np.mean(inb)
        if prop > 0:
            ece += np.abs(np.mean(conf[inb]) - np.mean(acc[inb])) * prop
    out["ece"] = ece
    return out


def clone_module(m):
    return copy.deepcopy(m)
