"""Flat HBM layout for a model's weights, gradients and momentum.

The reference keeps T separate tensors (T = 6 / 61 / 108) and loops over them
in Python (inference/optim_sghmc.py:43-67); ``util.flatten`` (util.py:163-169)
defines the flat order we keep: ``model.parameters()`` order, each tensor
row-major, no padding between tensors.  Here the flat buffer is the storage:
every ``param.data`` and ``param.grad`` becomes a view into one allocation, so
the whole update is one kernel launch and a posterior sample is one row copy.
Float buffers (BatchNorm running statistics) get the same treatment in a
second buffer so a sample snapshot captures them too.

Never call ``model.to()`` / ``.cpu()`` on a flattened model: that replaces
``param.data`` and silently detaches the views (the reference does exactly
that to snapshot a sample, sghmc.py:99-100 -- we snapshot by row copy).
"""
import torch


def _round_up(n, m):
    return (n + m - 1) // m * m


class FlatParams:
    """Owns ``p`` / ``g`` / ``v`` ([ld] fp32 on the device) and re-points the parameters at them."""

    PAD = 4   # 16-byte rows for 128-bit vector access and bulk-async copies

    def __init__(self, params, buffers=(), device=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatParams: no parameters")
        device = torch.device(device) if device is not None else self.params[0].device
        if device.type != "cuda":
            raise ValueError("FlatParams needs a CUDA device: this engine has no CPU path (got %s)" % device)
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("FlatParams: only float32 parameters are supported")
        self.device = device
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.D = self.offsets[-1]
        self.ld = _round_up(self.D, self.PAD)
        self.p = torch.zeros(self.ld, dtype=torch.float32, device=device)
        self.g = torch.zeros(self.ld, dtype=torch.float32, device=device)
        self.v = None
        for p, off, n in zip(self.params, self.offsets, self.sizes):
            view = self.p[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.g[off:off + n].view(p.shape)
            p._ursa_flat = self
        # float buffers (BatchNorm running_mean / running_var) in named_buffers() order
        self.buffer_owners = []
        for mod, name in buffers:
            b = mod._buffers[name]
            if b is not None and b.dtype == torch.float32:
                self.buffer_owners.append((mod, name, b.numel(), tuple(b.shape)))
        self.nb = sum(o[2] for o in self.buffer_owners)
        self.ldb = _round_up(max(self.nb, 1), self.PAD)
        self.b = torch.zeros(self.ldb, dtype=torch.float32, device=device)
        off = 0
        for mod, name, n, shape in self.buffer_owners:
            view = self.b[off:off + n].view(shape)
            view.copy_(mod._buffers[name])
            mod._buffers[name] = view
            off += n

    @classmethod
    def from_model(cls, model, device=None):
        first = next(model.parameters())
        flat = getattr(first, "_ursa_flat", None)
        if flat is not None and flat.is_attached(model):
            return flat
        bufs = []
        for mod in model.modules():
            for name, b in mod._buffers.items():
                if b is not None and b.dtype == torch.float32:
                    bufs.append((mod, name))
        return cls(list(model.parameters()), bufs, device if device is not None else first.device)

    def is_attached(self, model=None):
        params = self.params if model is None else list(model.parameters())
        if len(params) != len(self.params):
            return False
        base = self.p.data_ptr()
        return all(q is p and p.data.data_ptr() == base + 4 * off
                   for q, p, off in zip(params, self.params, self.offsets))

    def momentum(self):
        if self.v is None:
            self.v = torch.zeros(self.ld, dtype=torch.float32, device=self.device)
        return self.v

    def sync_grads(self):
        """Make sure every ``p.grad`` is its view of ``g`` (a caller may have assigned ``p.grad = t`` or set it to
        None); foreign gradients are copied in.  Returns False for a tensor whose grad is None (zeros are used)."""
        base = self.g.data_ptr()
        for p, off, n in zip(self.params, self.offsets, self.sizes):
            gr = p.grad
            if gr is not None and gr.data_ptr() == base + 4 * off:
                continue
            view = self.g[off:off + n].view(p.shape)
            if gr is None:
                view.zero_()
            else:
                view.copy_(gr)
            p.grad = view

    def zero_grad(self):
        self.g.zero_()

    def load_vector(self, vec):
        """Copy a flat [>=D] vector (any device) into the parameters (reference util.set_weights, util.py:172-176)."""
        self.p[:self.D].copy_(vec[:self.D])

    def load_buffers(self, vec):
        if self.nb:
            self.b[:self.nb].copy_(vec[:self.nb])

    def reattach(self):
        """Re-point parameters whose ``.data`` was replaced (e.g. by ``reset_parameters`` creating new storage)."""
        for p, off, n in zip(self.params, self.offsets, self.sizes):
            view = self.p[off:off + n].view(p.shape)
            if p.data.data_ptr() != view.data_ptr():
                view.copy_(p.data)
                p.data = view
        self.sync_grads()
