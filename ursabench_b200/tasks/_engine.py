"""Sample-batched BMA accumulation engine shared by the tasks (Prediction, OODDetection, Decision).

The three reference tasks run the same loop -- for every test batch, for every posterior sample: move the model to the
device, forward, softmax, copy to the host, accumulate (tasks/prediction.py:52-75, tasks/ood_detection.py:52-100,
tasks/decision_making.py:118-142) -- and differ only in what they accumulate.  Here one engine keeps the test inputs
resident on the device, runs the sample-batched forward (fused K3 kernels for MLP / PreResNet, the model's own PyTorch
forward otherwise) and accumulates  sum_s softmax_s  and  sum_s entropy(smooth(softmax_s))  on the device with
``ursa_bma_accumulate``; the tasks derive their own statistics from those two accumulators.
"""
import copy

import torch

from .. import _C
from ..bank import BankedSample, SampleBank
from ..flat import FlatParams

_LOGIT_CHUNK_BYTES = 256 << 20


def _arch_of(module):
    """('mlp', in_dim, hidden, C) / ('preresnet', depth, C) / ('wrn', depth, widen, C) / None -- structural match against models.py."""
    name = type(module).__name__
    if name == "MLP" and all(hasattr(module, a) for a in ("fc1", "fc2", "fc3")):
        f1, f2, f3 = module.fc1, module.fc2, module.fc3
        if all(isinstance(f, torch.nn.Linear) and f.bias is not None for f in (f1, f2, f3)) \
                and f2.in_features == f1.out_features == f2.out_features == f3.in_features \
                and len(list(module.parameters())) == 6:
            return ("mlp", f1.in_features, f1.out_features, f3.out_features)
    if name == "PreResNet" and hasattr(module, "layer1") and hasattr(module, "fc"):
        blocks = list(module.layer1)
        if blocks and type(blocks[0]).__name__ == "BasicBlock" and module.fc.in_features == 64:
            return ("preresnet", 6 * len(blocks) + 2, module.fc.out_features)
    if name == "WideResNet" and hasattr(module, "layer1") and hasattr(module, "linear"):
        blocks = list(module.layer1)
        if blocks and type(blocks[0]).__name__ == "WideBasic" and module.linear.in_features % 64 == 0 \
                and blocks[0].conv1.bias is not None:            # dropout is the identity in eval mode (prediction.py:58)
            return ("wrn", 6 * len(blocks) + 4, module.linear.in_features // 64, module.linear.out_features)
    return None



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
        xs, ys = [], []
        for batch_data, batch_labels in loader:
            xs.append(batch_data)
            ys.append(batch_labels)
        self.targets = torch.cat(ys)
        self._n = len(loader.dataset)
        self._batch_sizes = [len(x) for x in xs]
        self._x = torch.cat(xs).to(self.device, non_blocking=True).float().contiguous()
        self._proba = torch.zeros(self._n, num_classes, device=self.device)
        self._entropy = torch.zeros(self._n, device=self.device)
        self._workers = {}
        self.last_engine = None
        self.kernel_launches = 0

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

    def accumulate(self, model_list):
        if model_list:
            with torch.no_grad():
                self._accumulate(model_list)

    def _accumulate(self, model_list):
        banked = all(isinstance(m, BankedSample) and m.is_pristine() for m in model_list)
        if banked and len({id(m._ursa_bank) for m in model_list}) == 1:
            bank = model_list[0]._ursa_bank
            w, b = bank.rows([m._ursa_row for m in model_list])
            self._accumulate_rows(w, b, _arch_of(bank.skeleton), bank.skeleton)
            return
        plain = [m.materialize() if isinstance(m, BankedSample) else m for m in model_list]
        arch = _arch_of(plain[0])
        if self.engine in ("auto", "ffma") and arch is not None and self._fused_available(arch) \
                and all(_arch_of(m) == arch for m in plain):
            bank = SampleBank.from_modules(plain, self.device)       # one H2D per sample instead of 2 per batch
            self._accumulate_rows(bank.w[:bank.count], bank.b[:bank.count], arch, None)
            return
        self._accumulate_generic_modules(plain)

    def _fused_available(self, arch):
        return self._pick_algo(arch) is not None

    def _pick_algo(self, arch):
        """Fastest engine whose workspace query accepts the shape: the tcgen05 paths (3xTF32, fp32-level accuracy,
        parity-tested at the same 1e-5 bar) first, the fp32 CUDA-core kernels for shapes they do not cover
        (e.g. an MLP width that is not a multiple of 4).  ``engine='ffma'`` pins the CUDA-core kernels."""
        lib = _C.lib()
        if arch[0] == "mlp":
            order = (_C.ALGO_FFMA,) if self.engine == "ffma" else (_C.ALGO_TCGEN05, _C.ALGO_FFMA)
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
            if lib.ursa_bma_wrn_workspace(1, 1, arch[1], arch[2], arch[3], _C.ALGO_TCGEN05) > 0:
                return _C.ALGO_TCGEN05
        return None

    def _accumulate_rows(self, w, b, arch, skeleton):
        S = w.shape[0]
        if self.engine in ("auto", "ffma") and arch is not None and self._fused_available(arch):
            algo = self._pick_algo(arch)
            if arch[0] == "mlp":
                _, in_dim, hidden, C = arch
                x2 = self._x.view(self._n, -1)
                if x2.shape[1] != in_dim or C != self.num_classes:
                    raise ValueError("MLP input / class dimensions do not match the task")
                self._ws = _C.bma_mlp_forward(w, S, x2, in_dim, hidden, C, self._proba, self._entropy, algo=algo,
                                              workspace=self._ws)
                self.last_engine = "fused_mlp"
            elif arch[0] == "wrn":
                _, depth, widen, C = arch
                if C != self.num_classes:
                    raise ValueError("WideResNet class dimension does not match the task")
                self._ws = _C.bma_wrn_forward(w, b, S, self._x, depth, widen, C, self._proba, self._entropy, algo=algo,
                                              workspace=self._ws)
                self.last_engine = "fused_wrn"
            else:
                _, depth, C = arch
                guard = algo == _C.ALGO_TCGEN05_FUSED_F16
                if guard:                                   # FP16-split operands: activations beyond ~1e6 overflow to NaN
                    keep = (self._proba.clone(), self._entropy.clone())
                self._ws = _C.bma_preresnet_forward(w, b, S, self._x, depth, C, self._proba, self._entropy, algo=algo,
                                                    workspace=self._ws)
                if guard and not bool(torch.isfinite(self._proba).all()):
                    # loud by construction (inf -> NaN logits): redo this call on the TF32 engine, which has fp32's range
                    self._proba.copy_(keep[0])
                    self._entropy.copy_(keep[1])
                    algo = _C.ALGO_TCGEN05_FUSED
                    self._ws = _C.bma_preresnet_forward(w, b, S, self._x, depth, C, self._proba, self._entropy, algo=algo,
                                                        workspace=None)
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

        self._accumulate_generic(S, load)

    def _worker_for(self, skeleton):
        key = id(skeleton)
        if key not in self._workers:
            worker = copy.deepcopy(skeleton).to(self.device)
            worker._ursa_worker_flat = FlatParams.from_model(worker, self.device)
            worker.eval()
            self._workers[key] = worker
        return self._workers[key]

    def _accumulate_generic_modules(self, plain):
        homes = [next(m.parameters()).device for m in plain]

        def load(i):
            plain[i].to(self.device)
            plain[i].eval()
            return plain[i]

        def unload(i):
            plain[i].to(homes[i])

        self._accumulate_generic(len(plain), load, unload)

    def _accumulate_generic(self, S, load, unload=None):
        """Per-sample PyTorch forward on the device over the resident test set; logits are gathered per chunk of
        samples and reduced by ONE ``ursa_bma_accumulate`` launch per chunk (sample order preserved)."""
        N, C = self._n, self.num_classes
        chunk = max(1, min(S, _LOGIT_CHUNK_BYTES // max(1, N * C * 4)))
        logits = torch.empty(chunk, N, C, device=self.device)
        tf32_matmul = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False          # fp32 parity with the reference's CPU forward
        try:
            with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
                self._generic_chunks(S, chunk, logits, load, unload)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_matmul
        self.last_engine = "generic"

    def _generic_chunks(self, S, chunk, logits, load, unload):
        for s0 in range(0, S, chunk):
            ns = min(chunk, S - s0)
            for j in range(ns):
                model = load(s0 + j)
                off = 0
                for bs in self._batch_sizes:
                    out = model(self._x[off:off + bs])
                    logits[j, off:off + bs] = out.float()
                    off += bs
                if unload is not None:
                    unload(s0 + j)
            _C.bma_accumulate(logits[:ns], self._proba, self._entropy)
            self.kernel_launches += 1

