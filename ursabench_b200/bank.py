"""Device-resident posterior-sample bank.

The reference hands samples around as a Python list of deep-copied CPU modules
(``deepcopy(self.model.cpu())``, inference/sghmc.py:99) and ``Prediction`` moves
every one of them to the device and back for every test batch
(tasks/prediction.py:57,64).  Here a sample is one row of a ``[S, ld]`` fp32
matrix in HBM (+ one row of BatchNorm running statistics); ``BankedSample`` is
the ``nn.Module`` handle returned to callers: tasks of this package read the
rows in place, anything else sees an ordinary CPU module materialised on first
use.
"""
import copy
from contextlib import nullcontext as _nullcontext

import torch
import torch.nn as nn

from .flat import _round_up


# Pinned host staging for from_modules(): cudaHostAlloc of ~100 MB costs tens of milliseconds, so the buffers are kept and
# handed out round-robin; a buffer is reused only after the H2D copy that read it has completed (event).
_STAGING = {}


def _staging(n, ld, ldb, cuda, slots=3):
    if not cuda:
        return torch.empty(n, ld), torch.zeros(n, ldb), None
    key = (ld, ldb, torch.cuda.current_device() if cuda is True else cuda)
    ring = _STAGING.setdefault(key, {"next": 0, "slots": []})
    if len(ring["slots"]) < slots:
        cap = max(n, 8)
        ring["slots"].append([torch.empty(cap, ld).pin_memory(), torch.zeros(cap, ldb).pin_memory(), torch.cuda.Event()])
        slot = ring["slots"][-1]
    else:
        slot = ring["slots"][ring["next"] % slots]
        ring["next"] += 1
        slot[2].synchronize()
        if slot[0].shape[0] < n:
            slot[0], slot[1] = torch.empty(n, ld).pin_memory(), torch.zeros(n, ldb).pin_memory()
    return slot[0][:n], slot[1][:n], slot[2]


class SampleBank:
    def __init__(self, D, nb, device, capacity=16, skeleton=None):
        self.D, self.nb = D, nb
        self.ld = _round_up(D, 4)
        self.ldb = _round_up(max(nb, 1), 4)
        self.device = torch.device(device)
        self.w = torch.zeros(max(capacity, 1), self.ld, dtype=torch.float32, device=self.device)
        self.b = torch.zeros(max(capacity, 1), self.ldb, dtype=torch.float32, device=self.device)
        self.count = 0
        self.skeleton = skeleton          # CPU module used to materialise samples lazily

    @property
    def capacity(self):
        return self.w.shape[0]

    def fresh(self):
        """A new, empty bank of the same shape.  Samplers switch to a fresh bank when they restart (``update_hyp``, a new
        ``HMC.sample()``), so ``BankedSample`` handles returned earlier keep their own storage alive instead of silently
        aliasing the new run's rows -- the reference hands out independent deep copies (inference/sghmc.py:99)."""
        return SampleBank(self.D, self.nb, self.device, capacity=min(self.capacity, 8), skeleton=self.skeleton)

    def reserve(self, n):
        if n <= self.capacity:
            return
        cap = max(n, 2 * self.capacity)
        w = torch.zeros(cap, self.ld, dtype=torch.float32, device=self.device)
        b = torch.zeros(cap, self.ldb, dtype=torch.float32, device=self.device)
        w[:self.count].copy_(self.w[:self.count])
        b[:self.count].copy_(self.b[:self.count])
        self.w, self.b = w, b

    def next_row(self):
        """Reserve the next row and return (index, weight row view [ld]) -- the K1 launch writes it (snapshot)."""
        self.reserve(self.count + 1)
        i = self.count
        self.count += 1
        return i, self.w[i]

    def append(self, flat_p, flat_b=None):
        i, row = self.next_row()
        row[:self.D].copy_(flat_p[:self.D])
        if flat_b is not None and self.nb:
            self.b[i, :self.nb].copy_(flat_b[:self.nb])
        return i

    def set_buffers(self, i, flat_b):
        if self.nb:
            self.b[i, :self.nb].copy_(flat_b[:self.nb])

    def handle(self, i):
        return BankedSample(self, i)

    def handles(self):
        """``BankedSample`` handles of rows [0, count): the list ``sample()`` returns / ``update_statistics`` takes."""
        return [BankedSample(self, i) for i in range(self.count)]

    def rows(self, idx):
        """(w [n, ld], b [n, ldb]) for a list of row indices; a view when the indices are one contiguous run."""
        if len(idx) and idx == list(range(idx[0], idx[0] + len(idx))):
            return self.w[idx[0]:idx[0] + len(idx)], self.b[idx[0]:idx[0] + len(idx)]
        sel = torch.as_tensor(idx, dtype=torch.long, device=self.device)
        return self.w.index_select(0, sel), self.b.index_select(0, sel)

    @classmethod
    def from_modules(cls, models, device):
        """Pack ordinary modules (e.g. CPU samples from the reference) into a bank: one host pass per sample into a
        pinned staging matrix, then ONE asynchronous H2D copy of the matrix."""
        first = models[0]
        D = sum(p.numel() for p in first.parameters())
        nb = sum(b.numel() for b in first.buffers() if b.dtype == torch.float32)
        bank = cls(D, nb, device, capacity=len(models), skeleton=None)
        cuda = bank.device.type == "cuda"
        dev_index = bank.device.index if bank.device.index is not None else (torch.cuda.current_device() if cuda else 0)
        with torch.cuda.device(dev_index) if cuda else _nullcontext():
            stage_w, stage_b, done = _staging(len(models), bank.ld, bank.ldb, cuda)
        for i, m in enumerate(models):
            ps = [p.detach().reshape(-1) for p in m.parameters()]
            if sum(p.numel() for p in ps) != D:
                raise ValueError("models in one ensemble must share an architecture")
            stage_w[i, :D].copy_(torch.cat(ps))
            if nb:
                stage_b[i, :nb].copy_(torch.cat([b.detach().reshape(-1) for b in m.buffers() if b.dtype == torch.float32]))
        bank.w[:len(models)].copy_(stage_w, non_blocking=True)
        bank.b[:len(models)].copy_(stage_b, non_blocking=True)
        if cuda:
            with torch.cuda.device(dev_index):
                done.record(torch.cuda.current_stream(bank.device))
        bank.count = len(models)
        return bank

    # ---- on-disk wire format ---------------------------------------------------------------------
    # The reference's interchange format is one ``state_dict`` pickle per sample (``sghmc_sample_%d.pt``,
    # experiment.py:77-80; read back one file at a time by trtprof/run_prediction.py:50-57,218-221).  Here an ensemble
    # is ONE file: the unpadded ``[S, D]`` weight matrix, the ``[S, nb]`` float-buffer matrix and a layout descriptor
    # (names / shapes / offsets in ``parameters()`` / ``buffers()`` order) that lets a reader rebuild per-sample
    # ``state_dict``s without this package.
    FORMAT = "ursa_b200.sample_bank"
    VERSION = 1

    def layout(self):
        """Layout descriptor of one row (None without a skeleton): parameter and float-buffer names, shapes, offsets."""
        if self.skeleton is None:
            return None
        params, off = [], 0
        for name, p in self.skeleton.named_parameters():
            params.append({"name": name, "shape": list(p.shape), "offset": off})
            off += p.numel()
        bufs, off = [], 0
        for name, b in self.skeleton.named_buffers():
            if b.dtype == torch.float32:
                bufs.append({"name": name, "shape": list(b.shape), "offset": off})
                off += b.numel()
        return {"params": params, "buffers": bufs}

    def save(self, path):
        """Write rows [0, count) as one file (one D2H copy of the two matrices)."""
        n = self.count
        torch.save({"format": self.FORMAT, "version": self.VERSION, "D": self.D, "nb": self.nb, "count": n,
                    "w": self.w[:n, :self.D].detach().cpu().contiguous(),
                    "b": self.b[:n, :self.nb].detach().cpu().contiguous(),
                    "layout": self.layout()}, path)

    @classmethod
    def load(cls, path, device, skeleton=None):
        """Read a bank file back onto ``device``.  ``skeleton`` (a CPU module of the same architecture) is checked
        against the stored layout and enables lazy materialisation of ``BankedSample`` handles."""
        blob = torch.load(path, map_location="cpu", weights_only=True)
        if not isinstance(blob, dict) or blob.get("format") != cls.FORMAT:
            raise ValueError("%s is not a %s file" % (path, cls.FORMAT))
        if blob["version"] > cls.VERSION:
            raise ValueError("%s: bank format version %d is newer than this reader (%d)" % (path, blob["version"], cls.VERSION))
        D, nb, n = int(blob["D"]), int(blob["nb"]), int(blob["count"])
        w, b = blob["w"], blob["b"]
        if tuple(w.shape) != (n, D) or tuple(b.shape) != (n, nb) or w.dtype != torch.float32:
            raise ValueError("%s: matrix shapes do not match the header" % path)
        bank = cls(D, nb, device, capacity=n, skeleton=skeleton)
        if skeleton is not None and blob["layout"] is not None and bank.layout() != blob["layout"]:
            raise ValueError("%s: stored layout does not match the skeleton module" % path)
        bank.w[:n, :D].copy_(w)
        if nb:
            bank.b[:n, :nb].copy_(b)
        bank.count = n
        return bank

    def state_dict_of(self, i):
        """Per-sample ``state_dict`` on the CPU in the reference's format (what ``torch.save(model.state_dict())`` writes)."""
        if self.skeleton is None:
            raise RuntimeError("this bank has no module skeleton to name its tensors")
        if not 0 <= i < self.count:
            raise IndexError(i)
        sd = copy.deepcopy(self.skeleton.state_dict())
        lay = self.layout()
        w, b = self.w[i].cpu(), self.b[i].cpu()
        for group, row in ((lay["params"], w), (lay["buffers"], b)):
            for e in group:
                if e["name"] in sd:          # non-persistent buffers are not part of a state_dict
                    n = 1
                    for d in e["shape"]:
                        n *= d
                    sd[e["name"]] = row[e["offset"]:e["offset"] + n].view(e["shape"]).clone()
        return sd

    def export_state_dicts(self, pattern):
        """Write one reference-style ``.pt`` per sample; ``pattern`` holds one ``%d`` (e.g. ``'run/sghmc_sample_%d.pt'``)."""
        paths = []
        for i in range(self.count):
            path = pattern % i
            torch.save(self.state_dict_of(i), path)
            paths.append(path)
        return paths

    @classmethod
    def from_state_dict_files(cls, paths, skeleton, device):
        """Pack reference-style per-sample ``state_dict`` files into a bank (the reverse of ``export_state_dicts``)."""
        models = []
        for path in paths:
            m = copy.deepcopy(skeleton)
            m.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
            models.append(m)
        bank = cls.from_modules(models, device)
        bank.skeleton = copy.deepcopy(skeleton).cpu()
        return bank


def _load_row_into(module, w_row, b_row):
    off = 0
    for p in module.parameters():
        n = p.numel()
        p.data.copy_(w_row[off:off + n].view(p.shape))
        off += n
    off = 0
    for b in module.buffers():
        if b.dtype == torch.float32:
            n = b.numel()
            b.copy_(b_row[off:off + n].view(b.shape))
            off += n


class BankedSample(nn.Module):
    """``nn.Module`` handle of one bank row.  Behaves like the deep-copied CPU module the reference returns
    (materialised lazily on the CPU); ``ursabench_b200.tasks`` recognise it and use the device row directly."""

    def __init__(self, bank, row):
        super().__init__()
        object.__setattr__(self, "_ursa_bank", bank)
        object.__setattr__(self, "_ursa_row", row)
        object.__setattr__(self, "_ursa_inner", None)

    def materialize(self):
        inner = self._ursa_inner
        if inner is None:
            bank = self._ursa_bank
            if bank.skeleton is None:
                raise RuntimeError("this bank has no module skeleton to materialise samples from")
            inner = copy.deepcopy(bank.skeleton)
            _load_row_into(inner, bank.w[self._ursa_row].cpu(), bank.b[self._ursa_row].cpu())
            inner.train(self.training)
            object.__setattr__(self, "_ursa_inner", inner)
        return inner

    def is_pristine(self):
        """True while nobody has touched the materialised copy (so the bank row is still the sample)."""
        return self._ursa_inner is None

    def forward(self, *args, **kwargs):
        return self.materialize()(*args, **kwargs)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name.startswith("_ursa"):
                raise
            return getattr(self.materialize(), name)

    # -- nn.Module surface that must act on the materialised module, not on this empty shell
    def parameters(self, recurse=True):
        return self.materialize().parameters(recurse)

    def named_parameters(self, *a, **k):
        return self.materialize().named_parameters(*a, **k)

    def buffers(self, recurse=True):
        return self.materialize().buffers(recurse)

    def named_buffers(self, *a, **k):
        return self.materialize().named_buffers(*a, **k)

    def children(self):
        return self.materialize().children()

    def named_children(self):
        return self.materialize().named_children()

    def modules(self):
        return self.materialize().modules()

    def named_modules(self, *a, **k):
        return self.materialize().named_modules(*a, **k)

    def state_dict(self, *a, **k):
        return self.materialize().state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self.materialize().load_state_dict(*a, **k)

    def apply(self, fn):
        self.materialize().apply(fn)
        return self

    def to(self, *a, **k):
        self.materialize().to(*a, **k)
        return self

    def cpu(self):
        self.materialize().cpu()
        return self

    def cuda(self, device=None):
        self.materialize().cuda(device)
        return self

    def train(self, mode=True):
        if self._ursa_inner is not None:
            self._ursa_inner.train(mode)
        return super().train(mode)

    def eval(self):
        return self.train(False)

    def __deepcopy__(self, memo):
        return copy.deepcopy(self.materialize(), memo)
