"""Shared device-side training loop of the SG-MCMC samplers (SGLD / SGHMC / cSGLD / cSGHMC).

One step = H2D of the batch, forward + backward through PyTorch autograd (the north star keeps these on stock
PyTorch; gradients land directly in the flat gradient buffer), then ONE fused K1 launch that applies the update,
draws the Langevin noise in-register, zeroes the gradients for the next step and -- on the last step of a
sampling epoch -- writes the thinned sample into its bank row.  Nothing syncs with the host inside an epoch:
the reference's per-step ``loss.item()`` (sghmc.py:82) becomes a device-side running sum read once per epoch.
"""
import copy

import torch

from ..bank import SampleBank
from ..flat import FlatParams
from .inference_base import require_cuda


class SGMCMCLoop:
    def _attach(self, model, train_loader, device, who):
        if not isinstance(model, torch.nn.Module):
            raise NotImplementedError
        self.device = require_cuda(device, who)
        if next(model.parameters()).device != self.device:
            model.to(self.device)
        # conv nets: feed NHWC activations -> cuDNN's channels-last BatchNorm / conv kernels (1.6x faster fwd+bwd for
        # PreResNet-20 at batch 128 on B200, profiles/r01_launches_bench_summary.txt); weights keep the flat layout
        self._channels_last = any(isinstance(m, torch.nn.Conv2d) for m in model.modules())
        self._skeleton = copy.deepcopy(model).cpu()
        self.flat = FlatParams.from_model(model, self.device)
        self.bank = SampleBank(self.flat.D, self.flat.nb, self.device, capacity=8, skeleton=self._skeleton)
        self._epoch_loss = None

    # -- one training step (public: bench.py and tests drive it directly) ------------------------------------------
    def train_step(self, batch_data, batch_labels, add_langevin_noise=True, snapshot=None):
        """H2D (non-blocking; pinned host tensors overlap) -> forward -> backward -> fused K1.  Returns the loss as a
        0-d device tensor (no host sync).  With ``enable_cuda_graph()`` the forward/backward/update run as one
        captured CUDA graph replay."""
        graph = getattr(self, "_graph", None)
        if graph is not None and snapshot is None:
            return graph.run(batch_data, batch_labels, add_langevin_noise)
        batch_data = batch_data.to(self.device, non_blocking=True)
        if self._channels_last and batch_data.dim() == 4:
            batch_data = batch_data.contiguous(memory_format=torch.channels_last)
        batch_labels = batch_labels.to(self.device, non_blocking=True)
        logits = self.model(batch_data)
        loss = self.loss_criterion(logits, batch_labels)
        loss.backward()
        self.optimizer.step(add_langevin_noise=bool(add_langevin_noise), snapshot=snapshot, zero_grad=True)
        return loss.detach()

    def enable_cuda_graph(self, example_data, example_labels):
        """Capture forward + backward + K1 into one CUDA graph (fixed batch shape).  The per-step scalars (lr, noise
        gate, Philox step counter) are read by K1 from device memory, so replays follow the schedule."""
        self._graph = _GraphedStep(self, example_data, example_labels)
        return self._graph

    def disable_cuda_graph(self):
        self._graph = None

    def _num_batches(self):
        try:
            return len(self.train_loader)
        except TypeError:
            return None

    def _run_epoch(self, noise_for_batch, lr_for_batch=None, snapshot_last=False, track_loss=False):
        """One pass over ``train_loader``.  ``noise_for_batch(batch_idx) -> bool``; ``lr_for_batch(batch_idx)`` may
        set the learning rate before each step.  Returns the bank row index if a snapshot was taken."""
        self.model.train()
        nb = self._num_batches()
        total = torch.zeros((), device=self.device) if track_loss else None
        row_idx = None
        self.optimizer.zero_grad()
        for batch_idx, (batch_data, batch_labels) in enumerate(self.train_loader):
            if lr_for_batch is not None:
                lr_for_batch(batch_idx)
            snap = None
            if snapshot_last and nb is not None and batch_idx == nb - 1:
                row_idx, snap = self.bank.next_row()
            loss = self.train_step(batch_data, batch_labels, noise_for_batch(batch_idx), snapshot=snap)
            if track_loss:
                total += loss * len(batch_data)
        if snapshot_last:
            if row_idx is None:                               # loader without __len__: copy after the fact
                row_idx = self.bank.append(self.flat.p)
            self.bank.set_buffers(row_idx, self.flat.b)
        self._epoch_loss = total
        return row_idx


class _GraphedStep:
    """forward + backward + K1 captured once; ``run`` = copy the batch into the static inputs, publish the scalars,
    replay.  Falls back to the eager step for a batch of a different shape (e.g. a ragged last batch)."""

    def __init__(self, owner, example_data, example_labels):
        self.owner = owner
        opt = owner.optimizer
        dev = owner.device
        if opt._dyn is None:
            opt.use_device_scalars(True)
        if getattr(owner, "_channels_last", False) and example_data.dim() == 4:
            self.x = torch.empty(example_data.shape, dtype=example_data.dtype, device=dev).contiguous(
                memory_format=torch.channels_last)
        else:
            self.x = torch.empty_like(example_data, device=dev)
        self.y = torch.empty_like(example_labels, device=dev)
        self.x.copy_(example_data)
        self.y.copy_(example_labels)
        flat = owner.flat
        keep = (flat.p.clone(), flat.b.clone(), None if flat.v is None else flat.v.clone(), opt.steps_done,
                list(opt._first), opt.launches)
        owner.model.train()
        opt.zero_grad()
        # Autograd caches one AccumulateGrad node per parameter for as long as any graph that references it is alive,
        # and the node remembers the stream it was created on.  Every loss below is therefore deleted before the next
        # forward, so the capture creates fresh nodes on the capture stream.
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            if opt.param_groups[0]["momentum"] != 0 and opt._first[0]:
                # the momentum-initialising first step (optim_sghmc.py:51-52) is a different kernel variant: run it
                # once so that the captured variant is the steady-state one (state is restored below)
                loss = owner.loss_criterion(owner.model(self.x), self.y)
                loss.backward()
                opt.step(add_langevin_noise=False, zero_grad=True)
                del loss
            for _ in range(3):                                   # warm-up on a side stream (cuDNN / cuBLAS handles)
                opt.refresh_device_scalars(False)
                loss = owner.loss_criterion(owner.model(self.x), self.y)
                loss.backward()
                opt.step_captured(zero_grad=True)
                del loss
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        opt.refresh_device_scalars(False)
        with torch.cuda.graph(self.graph):
            loss = owner.loss_criterion(owner.model(self.x), self.y)
            loss.backward()
            opt.step_captured(zero_grad=True)
            self.loss = loss.detach()
        del loss
        # capture does not execute; undo the warm-up updates so the sampler state is exactly what it was
        flat.p.copy_(keep[0])
        flat.b.copy_(keep[1])
        if keep[2] is not None:
            flat.v.copy_(keep[2])
        flat.g.zero_()
        opt.steps_done, opt.launches = keep[3], keep[5]
        first_was = keep[4]
        self._needs_eager_first = bool(first_was[0]) and opt.param_groups[0]["momentum"] != 0
        opt._first = first_was
        self.replays = 0

    def run(self, batch_data, batch_labels, add_langevin_noise):
        owner, opt = self.owner, self.owner.optimizer
        if self._needs_eager_first or tuple(batch_data.shape) != tuple(self.x.shape):
            self._needs_eager_first = False
            graph, owner._graph = owner._graph, None
            try:
                return owner.train_step(batch_data, batch_labels, add_langevin_noise)
            finally:
                owner._graph = graph
        self.x.copy_(batch_data, non_blocking=True)
        self.y.copy_(batch_labels, non_blocking=True)
        opt.refresh_device_scalars(bool(add_langevin_noise))
        self.graph.replay()
        self.replays += 1
        return self.loss
