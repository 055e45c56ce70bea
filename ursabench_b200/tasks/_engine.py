"""Sample-batched BMA accumulation engine shared by the tasks (Prediction, OODDetection, Decision).

The three reference tasks run the same loop -- for every test batch, for every posterior sample: move the model to the
device, forward, softmax, copy to the host, accumulate (tasks/prediction.py:52-75, tasks/ood_detection.py:52-100,
tasks/decision_making.py:118-142) -- and differ only in what they accumulate.  Here one engine keeps the test inputs
resident on the device, runs the sample-batched forward (fused K3 kernels for MLP / PreResNet, the model's own PyTorch
forward otherwise) and accumulates  sum_s softmax_s  and  sum_s entropy(smooth(softmax_s))  on the device with
``ursa_bma_accumulate``; the tasks derive their own statistics from those two accumulators.
"""
import copy
import os
import weakref

import torch

from .. import _C
from ..bank import BankedSample, SampleBank
from ..flat import FlatParams

_LOGIT_CHUNK_BYTES = 256 << 20
_PACK_BATCH = 8            # = the conv forwards' sample chunk
_PACK_SPLIT_MIN = int(os.environ.get("URSA_PACK_SPLIT_MIN", "4"))    # lists longer than this are packed in overlapped sub-batches


def pack_plan(n_modules, family):
    """Sizes of the sub-batches a list of ``n_modules`` host modules is packed and uploaded in: 2 first (nothing runs on the
    device while the first sub-batch is being packed), then ``_PACK_BATCH`` at a time, never a one-module tail (its launch chain
    costs more than its packing hides).  An MLP module packs in microseconds and a whole MLP evaluation takes a few ms, so
    only long MLP lists are split."""
    split_min = _PACK_SPLIT_MIN if family != "mlp" else 2 * _PACK_BATCH
    if n_modules <= max(split_min, 3):
        return [n_modules] if n_modules else []
    plan, left = [2], n_modules - 2
    while left:
        n = min(_PACK_BATCH, left)
        if left - n == 1 and n > 2:
            n -= 1
        plan.append(n)
        left -= n
    return plan


_ARCH_CACHE = weakref.WeakKeyDictionary()


def _arch_of(module):
    """('mlp', in_dim, hidden, C) / ('preresnet', depth, C) / ('wrn', depth, widen, C) / None -- structural match against
    models.py (cached per module object: the walk over ~100 tensors is not free when an ensemble has hundreds of members)."""
    try:
        return _ARCH_CACHE[module]
    except (KeyError, TypeError):
        pass
    arch = _arch_of_uncached(module)
    try:
        _ARCH_CACHE[module] = arch
    except TypeError:
        pass
    return arch


def _arch_of_uncached(module):
    name = type(module).__name__
    if name == "MLP" and all(hasattr(module, a) for a in ("fc1", "fc2", "fc3")):
        f1, f2, f3 = module.fc1, module.fc2, module.fc3
        if all(isinstance(f, torch.nn.Linear) and f.bias is not None for f in (f1, f2, f3)) \
                and f2.in_features == f1.out_features == f2.out_features == f3.in_features \
                and len(list(module.parameters())) == 6:
            return ("mlp", f1.in_features, f1.out_features, f3.out_features)
    if name == "PreResNet" and all(hasattr(module, a) for a in ("conv1", "layer1", "layer2", "layer3", "bn", "fc")):
        blocks = list(module.layer1)
        n = len(blocks)
        if n and type(blocks[0]).__name__ == "BasicBlock" and module.fc.in_features == 64 \
                and len(list(module.layer2)) == n and len(list(module.layer3)) == n \
                and tuple(module.conv1.weight.shape) == (16, 3, 3, 3) and module.conv1.bias is None \
                and len(list(module.parameters())) == 6 * n * 3 + 3 + 2 + 2:
            # 61 tensors at depth 20: conv1, 3n blocks x (bn1 w/b, conv1, bn2 w/b, conv2), 2 downsamples, bn w/b, fc w/b
            C = module.fc.out_features
            if sum(q.numel() for q in module.parameters()) == _preresnet_numel(n, C):
                return ("preresnet", 6 * n + 2, C)
    if name == "WideResNet" and all(hasattr(module, a) for a in ("conv1", "layer1", "layer2", "layer3", "bn1", "linear")):
        blocks = list(module.layer1)
        n = len(blocks)
        if n and type(blocks[0]).__name__ == "WideBasic" and module.linear.in_features % 64 == 0 \
                and blocks[0].conv1.bias is not None and len(list(module.layer2)) == n and len(list(module.layer3)) == n \
                and tuple(module.conv1.weight.shape[1:]) == (3, 3, 3):   # dropout is the identity in eval mode (prediction.py:58)
            return ("wrn", 6 * n + 4, module.linear.in_features // 64, module.linear.out_features)
    return None


def _preresnet_numel(n, C):
    """Parameter count of the reference's BasicBlock PreResNet (models/preresnet.py:98-151) with n blocks per stage."""
    total = 16 * 27
    cin = 16
    for stage, ch in enumerate((16, 32, 64)):
        for b in range(n):
            c_in = cin if b == 0 else ch
            total += 2 * c_in + ch * c_in * 9 + 2 * ch + ch * ch * 9
            if b == 0 and stage > 0:
                total += ch * c_in
        cin = ch
    return total + 2 * 64 + 64 * C + C


def wrn_dropout_is_identity(module):
    """True when every Dropout of a (Wide)ResNet has p == 0, i.e. the train-mode engine pass equals the module's."""
    return all(m.p == 0 for m in module.modules() if isinstance(m, torch.nn.Dropout))



class BMAAccumulator:
    """Device-resident inputs + accumulators for ONE data loader.  ``engine``: 'auto' | 'ffma' (pin the fp32 CUDA-core
    kernels) | 'generic' (per-sample PyTorch forward)."""

    def __init__(self, loader=None, num_classes=None, device=None, engine="auto"):
        if loader is not None:
            self._setup(loader, num_classes, device, engine)

    def _setup(self, loader, num_classes, device, engine="auto"):
        _C.lib()
        self.num_classes = num_classes
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("%s: device must be a CUDA device -- ursabench_b200 has no CPU path" % type(self).__name__)
        self.engine = engine
        self._ws = None           # K3 workspace, kept across calls
        self.last_algo = None
        # one pass over the loader (the reference tasks do the same to cache the targets); the inputs are uploaded once
        # and stay resident -- the loader must not shuffle
        whole = self._whole_tensors(loader)
        if whole is not None:
            # a sequential DataLoader over a TensorDataset yields exactly dataset.tensors in order: upload them in one
            # copy instead of re-collating N / batch_size batches on the host
            x_host, self.targets = whole
            bs = loader.batch_size
            self._batch_sizes = [min(bs, len(x_host) - i) for i in range(0, len(x_host), bs)]
        else:
            xs, ys = [], []
            for batch_data, batch_labels in loader:
                xs.append(batch_data)
                ys.append(batch_labels)
            self.targets = torch.cat(ys)
            self._batch_sizes = [len(x) for x in xs]
            x_host = torch.cat(xs)
        self._n = len(loader.dataset)
        self.h2d_bytes = x_host.numel() * x_host.element_size()       # test set uploaded by this constructor
        self.h2d_sample_bytes = 0                                      # posterior samples uploaded by accumulate()
        # labels first: a blocking copy issued AFTER the image upload would hold the host until those ~120 MB have landed
        # (2.3 ms during which no posterior sample gets packed)
        self._y_dev = self.targets.to(self.device, non_blocking=self.targets.is_pinned())
        self._x = x_host.to(self.device, non_blocking=True).float().contiguous()
        self._proba = torch.zeros(self._n, num_classes, device=self.device)
        self._entropy = torch.zeros(self._n, device=self.device)
        self._workers = {}
        self._pending = []        # FP16-split forwards of the current accumulate() awaiting the range check
        self._ws_key = None       # launch shape + address of the workspace whose plane-image pads are known to be zero
        self.last_engine = None
        self.kernel_launches = 0

    @staticmethod
    def _whole_tensors(loader):
        """(x, y) when ``loader`` is a plain sequential, non-dropping DataLoader over a 2-tensor TensorDataset, else None."""
        data = torch.utils.data
        ds = getattr(loader, "dataset", None)
        if type(loader) is not data.DataLoader or type(ds) is not data.TensorDataset or len(ds.tensors) != 2:
            return None
        if not isinstance(loader.sampler, data.SequentialSampler) or loader.drop_last or loader.batch_size is None \
                or loader.collate_fn is not data.default_collate:
            return None
        return ds.tensors[0], ds.tensors[1]

    @staticmethod
    def as_model_list(models):
        """The reference's argument check (prediction.py:38-50): a list of modules or a single module."""
        if isinstance(models, list):
            if not all(isinstance(m, torch.nn.Module) for m in models):
                raise NotImplementedError
            return models
        if isinstance(models, torch.nn.Module):
            return [models]
        raise NotImplementedError

    def accumulate(self, model_list, pairs=None):
        """``pairs`` = None: every model over every image.  Otherwise a list of ``(index into model_list, img_lo, img_hi)``
        -- this rank's share of the (sample, image) grid (``dist.shard_pairs``); the other ranks' shares meet in the
        all-reduce of the accumulators."""
        if not model_list:
            return
        with torch.no_grad():
            if pairs is None:
                self._accumulate(model_list, 0, self._n)
            else:
                i = 0
                while i < len(pairs):                   # runs of samples over the same image range go down in one call
                    j = i
                    while j + 1 < len(pairs) and pairs[j + 1][1:] == pairs[i][1:]:
                        j += 1
                    _, lo, hi = pairs[i]
                    if hi > lo:
                        self._accumulate([model_list[k] for k, _, _ in pairs[i:j + 1]], lo, hi)
                    i = j + 1
            self._commit_scratch()

    def _accumulate(self, model_list, lo, hi):
        banked = all(isinstance(m, BankedSample) and m.is_pristine() for m in model_list)
        if banked and len({id(m._ursa_bank) for m in model_list}) == 1:
            bank = model_list[0]._ursa_bank
            w, b = bank.rows([m._ursa_row for m in model_list])
            self._accumulate_rows(w, b, _arch_of(bank.skeleton), bank.skeleton, lo, hi)
            return
        plain = [m.materialize() if isinstance(m, BankedSample) else m for m in model_list]
        arch = _arch_of(plain[0])
        if self.engine in ("auto", "ffma") and arch is not None and self._fused_available(arch) \
                and all(_arch_of(m) == arch for m in plain):
            # one H2D per sample instead of 2 per batch -- in sub-batches, so that packing sub-batch i + 1 on the host
            # (a Python walk over ~100 tensors per module) overlaps the forward of sub-batch i on the device
            # The first sub-batch is small: nothing runs on the device while it is being packed.
            # (≈ 1 ms of host time per PreResNet-20 module: a rank's 12-sample share of an 8-rank evaluation packed in one go
            # left the device idle for a fifth of the call.)
            s0 = 0
            for n in pack_plan(len(plain), arch[0]):
                bank = SampleBank.from_modules(plain[s0:s0 + n], self.device)
                self.h2d_sample_bytes += bank.count * (bank.ld + bank.ldb) * 4
                self._accumulate_rows(bank.w[:bank.count], bank.b[:bank.count], arch, None, lo, hi)
                s0 += n
            return
        self._accumulate_generic_modules(plain, lo, hi)

    def _fused_available(self, arch):
        return self._pick_algo(arch) is not None

    def _pick_algo(self, arch):
        """Fastest engine whose workspace query accepts the shape: the tcgen05 paths (3xTF32, fp32-level accuracy,
        parity-tested at the same 1e-5 bar) first, the fp32 CUDA-core kernels for shapes they do not cover
        (e.g. an MLP width that is not a multiple of 4).  ``engine='ffma'`` pins the CUDA-core kernels."""
        lib = _C.lib()
        if arch[0] == "mlp":
            order = (_C.ALGO_FFMA,) if self.engine == "ffma" else (_C.ALGO_TCGEN05_F16, _C.ALGO_TCGEN05, _C.ALGO_FFMA)
            for algo in order:
                if lib.ursa_bma_mlp_workspace(1, 1, arch[1], arch[2], arch[3], algo) > 0:
                    return algo
        elif arch[0] == "preresnet":
            order = (_C.ALGO_FFMA,) if self.engine == "ffma" else (_C.ALGO_TCGEN05_FUSED_F16, _C.ALGO_TCGEN05_FUSED,
                                                                  _C.ALGO_TCGEN05, _C.ALGO_FFMA)
            for algo in order:
                if lib.ursa_bma_preresnet_workspace(1, 1, arch[1], arch[2], algo) > 0:
                    return algo
        elif arch[0] == "wrn" and self.engine != "ffma":
            for algo in (_C.ALGO_TCGEN05_F16, _C.ALGO_TCGEN05):
                if lib.ursa_bma_wrn_workspace(1, 1, arch[1], arch[2], arch[3], algo) > 0:
                    return algo
        return None

    def _accumulate_rows(self, w, b, arch, skeleton, lo=0, hi=None):
        S = w.shape[0]
        hi = self._n if hi is None else hi
        x, proba, entropy = self._x[lo:hi], self._proba[lo:hi], self._entropy[lo:hi]     # contiguous row ranges
        if self.engine in ("auto", "ffma") and arch is not None and self._fused_available(arch) \
                and self._inputs_match(arch):
            algo = self._pick_algo(arch)
            if arch[0] == "mlp":
                _, in_dim, hidden, C = arch
                x2 = self._x.view(self._n, -1)[lo:hi]
                if x2.shape[1] != in_dim or C != self.num_classes:
                    raise ValueError("MLP input / class dimensions do not match the task")
                if w.shape[1] < (in_dim + 1) * hidden + (hidden + 1) * hidden + (hidden + 1) * C:
                    raise ValueError("bank rows are shorter than the MLP's parameter vector")
                if algo == _C.ALGO_TCGEN05_F16:
                    # FP16-split operands: inputs / activations beyond 65 504 overflow to inf -> NaN logits (loud by
                    # construction); same scratch-and-commit protocol as the PreResNet FP16-split engine below
                    sp, se = self._scratch()
                    self._ws = _C.bma_mlp_forward(w, S, x2, in_dim, hidden, C, sp[lo:hi], se[lo:hi], algo=algo, workspace=self._ws)
                    self._pending.append(("mlp", w, S, in_dim, hidden, C, lo, hi))
                else:
                    self._ws = _C.bma_mlp_forward(w, S, x2, in_dim, hidden, C, proba, entropy, algo=algo, workspace=self._ws)
                self._ws_key = None
                self.last_engine = "fused_mlp"
            elif arch[0] == "wrn":
                _, depth, widen, C = arch
                if C != self.num_classes:
                    raise ValueError("WideResNet class dimension does not match the task")
                if algo == _C.ALGO_TCGEN05_F16:       # fp16's range: scratch-and-commit like the other FP16-split engines
                    sp, se = self._scratch()
                    self._ws = _C.bma_wrn_forward(w, b, S, x, depth, widen, C, sp[lo:hi], se[lo:hi], algo=algo, workspace=self._ws)
                    self._pending.append(("wrn", w, b, S, depth, widen, C, lo, hi))
                else:
                    self._ws = _C.bma_wrn_forward(w, b, S, x, depth, widen, C, proba, entropy, algo=algo, workspace=self._ws)
                self._ws_key = None
                self.last_engine = "fused_wrn"
            else:
                _, depth, C = arch
                if C != self.num_classes:
                    raise ValueError("PreResNet class dimension does not match the task")
                if algo == _C.ALGO_TCGEN05_FUSED_F16:
                    # FP16-split operands: activations beyond ~1e6 overflow to inf -> NaN logits (loud by construction).
                    # The calls of one accumulate() go into a zeroed scratch pair; _commit_scratch() reads ONE flag at the
                    # end and either adds the scratch or redoes the calls on the TF32 engine: the accumulators never see
                    # a NaN, nothing is cloned and the host does not synchronise between the calls.
                    sp, se = self._scratch()
                    # the engine's plane images need zeroed pad positions (2.7 GB of memset per call at the default chunking):
                    # skipped when this object's previous call left them so -- same private workspace, same launch shape
                    key = (min(S, 8), hi - lo, depth, C, None if self._ws is None else self._ws.data_ptr())
                    kept = _C.ALGO_FLAG_WS_KEPT if key == self._ws_key else 0
                    self._ws = _C.bma_preresnet_forward(w, b, S, x, depth, C, sp[lo:hi], se[lo:hi], algo=algo | kept,
                                                        workspace=self._ws)
                    self._ws_key = key[:4] + (self._ws.data_ptr(),)
                    self._pending.append(("preresnet", w, b, S, depth, C, lo, hi))
                else:
                    self._ws = _C.bma_preresnet_forward(w, b, S, x, depth, C, proba, entropy, algo=algo, workspace=self._ws)
                    self._ws_key = None
                self.last_engine = "fused_preresnet"
            self.last_algo = algo
            self.kernel_launches += 1
            return
        if skeleton is None:
            raise RuntimeError("no module skeleton available for the generic forward")
        worker = self._worker_for(skeleton)
        flat = worker._ursa_worker_flat

        def load(i):
            flat.load_vector(w[i])
            flat.load_buffers(b[i])
            return worker

        self._accumulate_generic(S, load, None, lo, hi)

    def _scratch(self):
        if getattr(self, "_scratch_p", None) is None:
            self._scratch_p = torch.empty(self._n, self.num_classes, device=self.device)
            self._scratch_e = torch.empty(self._n, device=self.device)
        if not self._pending:
            self._scratch_p.zero_()
            self._scratch_e.zero_()
        return self._scratch_p, self._scratch_e

    def _commit_scratch(self):
        """End of an accumulate(): fold the FP16-split engine's scratch sums into the accumulators, or -- if any logit
        overflowed -- redo those calls on the TF32 engine, which has fp32's range."""
        if not self._pending:
            return
        pending, self._pending = self._pending, []
        if bool(torch.isfinite(self._scratch_p.sum())):
            self._proba.add_(self._scratch_p)
            self._entropy.add_(self._scratch_e)
            return
        ws = None
        for job in pending:
            if job[0] == "wrn":
                _, w, b, S, depth, widen, C, lo, hi = job
                self.last_algo = _C.ALGO_TCGEN05
                ws = _C.bma_wrn_forward(w, b, S, self._x[lo:hi], depth, widen, C, self._proba[lo:hi], self._entropy[lo:hi],
                                        algo=_C.ALGO_TCGEN05, workspace=ws)
            elif job[0] == "mlp":
                _, w, S, in_dim, hidden, C, lo, hi = job
                self.last_algo = _C.ALGO_TCGEN05
                ws = _C.bma_mlp_forward(w, S, self._x.view(self._n, -1)[lo:hi], in_dim, hidden, C, self._proba[lo:hi],
                                        self._entropy[lo:hi], algo=_C.ALGO_TCGEN05, workspace=ws)
            else:
                _, w, b, S, depth, C, lo, hi = job
                self.last_algo = _C.ALGO_TCGEN05_FUSED
                ws = _C.bma_preresnet_forward(w, b, S, self._x[lo:hi], depth, C, self._proba[lo:hi], self._entropy[lo:hi],
                                              algo=_C.ALGO_TCGEN05_FUSED, workspace=ws)

    def _inputs_match(self, arch):
        """The fused conv forwards take only (pointer, N): the resident test tensor must be the [N, 3, 32, 32] the
        CIFAR-shaped networks expect -- anything else goes through the module's own forward (which raises the
        reference's shape error)."""
        if arch[0] == "mlp":
            return True
        return tuple(self._x.shape[1:]) == (3, 32, 32)

    def _worker_for(self, skeleton):
        key = id(skeleton)
        if key not in self._workers:
            worker = copy.deepcopy(skeleton).to(self.device)
            worker._ursa_worker_flat = FlatParams.from_model(worker, self.device)
            worker.eval()
            self._workers[key] = worker
        return self._workers[key]

    def _accumulate_generic_modules(self, plain, lo=0, hi=None):
        homes = [next(m.parameters()).device for m in plain]

        def load(i):
            plain[i].to(self.device)
            plain[i].eval()
            return plain[i]

        def unload(i):
            plain[i].to(homes[i])

        self._accumulate_generic(len(plain), load, unload, lo, hi)

    def _accumulate_generic(self, S, load, unload=None, lo=0, hi=None):
        """Per-sample PyTorch forward on the device over the resident test set; logits are gathered per chunk of
        samples and reduced by ONE ``ursa_bma_accumulate`` launch per chunk (sample order preserved)."""
        hi = self._n if hi is None else hi
        N, C = hi - lo, self.num_classes
        chunk = max(1, min(S, _LOGIT_CHUNK_BYTES // max(1, N * C * 4)))
        logits = torch.empty(chunk, N, C, device=self.device)
        tf32_matmul = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False          # fp32 parity with the reference's CPU forward
        try:
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
                self._generic_chunks(S, chunk, logits, load, unload, lo, hi)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_matmul
        self.last_engine = "generic"

    def _generic_chunks(self, S, chunk, logits, load, unload, lo, hi):
        bs = max(self._batch_sizes) if self._batch_sizes else 128
        for s0 in range(0, S, chunk):
            ns = min(chunk, S - s0)
            for j in range(ns):
                model = load(s0 + j)
                for off in range(lo, hi, bs):                    # eval mode: batch boundaries do not change the result
                    end = min(hi, off + bs)
                    logits[j, off - lo:end - lo] = model(self._x[off:end]).float()
                if unload is not None:
                    unload(s0 + j)
            _C.bma_accumulate(logits[:ns], self._proba[lo:hi], self._entropy[lo:hi])
            self.kernel_launches += 1

