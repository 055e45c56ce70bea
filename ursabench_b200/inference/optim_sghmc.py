"""``optimSGHMC`` on the fused K1 kernel.

Same constructor, ``param_groups`` keys and ``step(add_langevin_noise=True, closure=None)`` as the reference
optimizer (inference/optim_sghmc.py:7-68), so ``CosineAnnealingLR`` and cSGHMC's per-iteration
``param_group['lr'] = lr`` (inference/csghmc.py:70-71) keep working.  The per-tensor loop of 7-8 ATen launches
(:43-67) is replaced by ONE ``ursa_sgmcmc_step`` launch over the flat buffer of each parameter group.

Differences that are deliberate:
* noise is drawn in-register from counter-based Philox (seeded from ``torch.initial_seed()`` unless ``seed`` is
  given) instead of ``torch.randn_like`` (:64); pass ``noise=`` to ``step`` to inject a tensor (parity mode);
* ``zero_grad`` keeps the gradients as views of the flat buffer (one memset) instead of setting them to None;
* parameters whose gradient is None are treated as having a zero gradient (the reference skips them, :44-45);
* an optional ``temperature`` (default 1 = the reference) scales the injected noise by sqrt(T).
"""
import math

import torch
from torch.optim.optimizer import Optimizer, required

from .. import _C
from ..flat import FlatParams


class optimSGHMC(Optimizer):
    def __init__(self, params, lr=required, momentum=0, dampening=0, weight_decay=0, num_training_samples=None,
                 nesterov=False, seed=None, elem_offset=0, temperature=1.0):
        if lr is not required and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        if nesterov:
            raise NotImplementedError("the reference never enables nesterov (optim_sghmc.py:57-58); not built")
        if temperature < 0.0:
            raise ValueError("Invalid temperature value: {}".format(temperature))
        # temperature: the noise term becomes z * sqrt(2 (1 - momentum) lr T) / N.  The reference has no such factor
        # (optim_sghmc.py:63-64) = T = 1, which -- with its mean-loss / N scalings -- samples exp(-N U);
        # T = N targets the posterior exp(-U) itself.
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                        num_training_samples=num_training_samples, temperature=temperature)
        super().__init__(params, defaults)
        _C.lib()                                   # fail now, loudly, if the CUDA library is missing
        self._flats = []
        for group in self.param_groups:
            ps = group["params"]
            flat = getattr(ps[0], "_ursa_flat", None)
            if flat is None or not (len(flat.params) == len(ps) and all(a is b for a, b in zip(flat.params, ps))
                                    and flat.is_attached()):
                flat = FlatParams(ps)
            self._flats.append(flat)
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.elem_offset = int(elem_offset)
        self._first = [True] * len(self.param_groups)
        self.steps_done = 0
        self.launches = 0
        self._dyn = None               # device scalars for CUDA-graph replay (see use_device_scalars)

    @property
    def flat(self):
        return self._flats[0]

    def zero_grad(self, set_to_none=False):
        for flat in self._flats:
            flat.sync_grads()
            flat.zero_grad()

    # -- CUDA-graph support: the K1 launch reads lr / momentum / wd / noise scale / step counter from device memory
    def use_device_scalars(self, enable=True):
        if len(self._flats) != 1:
            raise ValueError("device scalars need a single parameter group")
        self._dyn = torch.zeros(8, dtype=torch.float32, device=self.flat.device) if enable else None

    def _scalars(self, group, add_langevin_noise):
        momentum, lr, wd, n_train = group["momentum"], group["lr"], group["weight_decay"], group["num_training_samples"]
        if (add_langevin_noise or wd != 0) and not n_train:
            raise ValueError("num_training_samples is required")
        wd_over_n = (wd / n_train) if wd != 0 else 0.0
        noise_mul = math.sqrt(2 * (1 - momentum) * lr * group.get("temperature", 1.0)) if add_langevin_noise else 0.0
        noise_div = float(n_train) if add_langevin_noise else 1.0
        return lr, momentum, wd_over_n, noise_mul, noise_div

    def refresh_device_scalars(self, add_langevin_noise=True):
        """Publish this step's scalars to the device (one 1-thread launch) and advance the step counter; call right
        before replaying a graph that contains ``step_captured``."""
        lr, momentum, wd_over_n, noise_mul, noise_div = self._scalars(self.param_groups[0], add_langevin_noise)
        _C.sgmcmc_set_dyn(self._dyn, lr, momentum, wd_over_n, noise_mul / noise_div, self.steps_done)
        self.steps_done += 1
        self.launches += 1

    @torch.no_grad()
    def step_captured(self, zero_grad=True):
        """The K1 launch for use INSIDE graph capture (scalars come from ``refresh_device_scalars``)."""
        flat = self.flat
        momentum = self.param_groups[0]["momentum"]
        if momentum != 0 and self._first[0]:
            raise RuntimeError("run the first (momentum-initialising) step eagerly before capturing")
        _C.sgmcmc_step_dyn(flat.p, flat.g, flat.momentum() if momentum != 0 else None, None, None, self._dyn,
                           add_noise=True, zero_grad=zero_grad, seed=self.seed, elem_offset=self.elem_offset)

    @torch.no_grad()
    def step(self, add_langevin_noise=True, closure=None, noise=None, snapshot=None, zero_grad=False):
        """One SG-MCMC update.  ``noise``: optional flat N(0,1) tensor (parity mode); ``snapshot``: optional flat
        row that receives the updated weights (thinned sample); ``zero_grad``: fuse ``optimizer.zero_grad()``."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if (noise is not None or snapshot is not None) and len(self._flats) != 1:
            raise ValueError("noise= / snapshot= need a single parameter group")
        for gi, (group, flat) in enumerate(zip(self.param_groups, self._flats)):
            lr, momentum, wd_over_n, noise_mul, noise_div = self._scalars(group, add_langevin_noise)
            flat.sync_grads()
            v = flat.momentum() if momentum != 0 else None
            first = momentum != 0 and self._first[gi]
            _C.sgmcmc_step(flat.p, flat.g, v, snapshot, noise,
                           lr=lr, momentum=momentum, wd_over_n=wd_over_n, noise_mul=noise_mul, noise_div=noise_div,
                           first_step=first, add_noise=bool(add_langevin_noise), zero_grad=zero_grad,
                           seed=self.seed, step=self.steps_done, elem_offset=self.elem_offset + gi * (1 << 40))
            self.launches += 1
            if first:
                self._first[gi] = False
                for p, off, n in zip(flat.params, flat.offsets, flat.sizes):
                    self.state[p]["momentum_buffer"] = v[off:off + n].view(p.shape)
        self.steps_done += 1
        return loss
