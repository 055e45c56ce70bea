"""SWAG: SWA moments + batched diagonal / low-rank posterior draws (reference inference/swag.py:12-147).

Reference quirks (verified on the live reference, tests/golden/swag_compat_facts.json):
  Q5  every draw is overwritten by the mean (``weight_sample = self.weight_mean``, :98/:118);
  Q6  ``num_models_collected`` is never incremented, so the "mean" is the last iterate and the variance 1e-30;
  Q7  ``full_cov=True`` raises AttributeError (``self.swag_model.subspace``, :90).
``hyperparameters['reference_compat'] = True`` reproduces Q5/Q6 bit for bit (Q7 raises the same AttributeError).
The default is the algorithm the code was written to be (Maddox et al. 2019): n increments after each collect and
the draw of swag.py:85-97 is returned -- all S draws produced by ONE pass over the deviation ring (K2b).

Multi-GPU: ``hyperparameters['shard_draws'] = True`` under ``torch.distributed`` makes ONE fit serve
all ranks: rank 0 trains and collects, (mean, second moment, deviation ring, counters) are broadcast once (2.9 GB for
WRN-28-10 at K = 20: milliseconds over NVLink), and ``sample(num_samples)`` returns on rank r the draws
s = r, r + G, r + 2G, ... -- each rank draws from its own Philox / generator substream, finishes (BatchNorm
re-estimation) and keeps its draws in its own bank, where a sample-sharded ``Prediction`` evaluates them.
"""
import torch

from .. import _C, dist as udist
from ..util import bn_update, check_bn
from .swa import SWA


class SWAG(SWA):
    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu"), **subspace_kwargs):
        if hyperparameters is None:
            hyperparameters = dict(self._defaults)
        super().__init__(hyperparameters, model=model, train_loader=train_loader, model_loss=model_loss,
                         device=device, **subspace_kwargs)
        self.num_samples = hyperparameters["num_samples"]
        self.reference_compat = bool(hyperparameters.get("reference_compat", False))
        self.shard_draws = bool(hyperparameters.get("shard_draws", False))
        self.weight_variance = None
        self._draw_calls = 0

    def update_hyp(self, hyperparameters, **subspace_kwargs):
        super().update_hyp(hyperparameters, **subspace_kwargs)
        self.num_samples = hyperparameters["num_samples"]
        self.reference_compat = bool(hyperparameters.get("reference_compat", False))
        self.shard_draws = bool(hyperparameters.get("shard_draws", False))
        self.weight_variance = None

    # -- one fit for all ranks -------------------------------------------------------------------------------------
    def _sharded(self):
        return self.shard_draws and udist.is_distributed()

    def _fit_or_receive(self, val_loader, debug_val_loss, wandb_debug):
        """Single process: fit.  Sharded draws: rank 0 fits, everybody receives its state by ONE round of broadcasts."""
        if not self._sharded():
            self._fit_moments(val_loader, debug_val_loss, wandb_debug)
            return
        import torch.distributed as dist
        rank, _ = udist.rank_world()
        if rank == 0:
            self._fit_moments(val_loader, debug_val_loss, wandb_debug)
        meta = torch.zeros(4, dtype=torch.int64, device=self.device)
        if rank == 0:
            meta[0] = int(self.num_models_collected.item())
            meta[1] = self.subspace.collected
            meta[2] = int(self.subspace.rank.item())
            meta[3] = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        dist.broadcast(meta, 0)
        dist.broadcast(self._mean, 0)
        dist.broadcast(self._sq, 0)
        dist.broadcast(self.subspace.ring, 0)
        n, collected, srank, seed = (int(v) for v in meta.tolist())
        self.num_models_collected = torch.full((1,), n, dtype=torch.long)
        self.subspace.collected = collected
        self.subspace.rank = torch.full((1,), srank, dtype=torch.long)
        self._draw_seed = seed
        self.burnt_in = True
        _, self.weight_variance = self._get_mean_and_variance()

    # -- training + collection (first call only), reference :54-83 ---------------------------------------------
    def _fit_moments(self, val_loader, debug_val_loss, wandb_debug):
        epochs = self.burn_in_epochs + self.num_iterates
        for epoch in range(epochs):
            total = self._sgd_epoch(track_loss=debug_val_loss)
            if debug_val_loss:
                self._debug(total, val_loader, wandb_debug)
            if epoch >= self.burn_in_epochs:
                self._collect_model()
                if not self.reference_compat:
                    self.num_models_collected += 1           # the increment the reference forgot (Q6)
        self.burnt_in = True
        _, self.weight_variance = self._get_mean_and_variance()

    # -- K2b -----------------------------------------------------------------------------------------------------
    def _draw_into_bank(self, num, full_cov):
        """Draw ``num`` weight vectors into fresh bank rows with ONE ``ursa_swag_draw`` call (one pass over the ring per 30 draws)."""
        if full_cov and self.reference_compat:
            # reference :90 dereferences self.swag_model.subspace, which does not exist (Q7)
            raise AttributeError("'%s' object has no attribute 'subspace'" % type(self.swag_model).__name__)
        first = self.bank.count
        self.bank.reserve(first + num)
        D = self.num_parameters
        var_full = torch.empty_like(self._mean)
        _C.swag_variance(self._mean, self._sq, var_full, self.var_clamp)
        rows = self.subspace.rows() if full_cov else None
        K = 0 if rows is None else rows.shape[0]
        # sharded draws: one seed for the whole job (broadcast with the fit), one substream per rank
        rank, world = udist.rank_world() if self._sharded() else (0, 1)
        seed = getattr(self, "_draw_seed", None) if self._sharded() else None
        seed = (int(torch.initial_seed()) if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        gen = None
        if world > 1:
            gen = torch.Generator(device=self.device)
            gen.manual_seed((seed + 7919 * (rank + 1)) & 0x7FFFFFFFFFFFFFFF)
            gen.set_offset(4 * 4096 * _C.DRAW_MAX_K * self._draw_calls)
        out = self.bank.w[first:first + num]
        if self.reference_compat:
            out[:, :D] = self._mean[:D]                        # Q5: the draw is discarded, the mean is returned
        else:
            z2 = torch.randn(num, K, device=self.device, generator=gen) if K else None
            _C.swag_draw(out, self._mean, var_full, D, ring=rows if K else None, z2=z2,
                         rank_div=float((self.subspace.max_rank - 1) ** 0.5),              # reference :95
                         seed=seed, step=self._draw_calls * world + rank)
            self._draw_calls += 1
        self.bank.count = first + num
        return list(range(first, first + num))

    def _engine_bn_update(self, row):
        """BatchNorm re-estimation on the tcgen05 conv engine (``ursa_wrn_bn_update``, WideResNets): the train-mode pass runs
        straight from the bank row and writes the row's running statistics in place.  The training images are uploaded
        once (one pass over ``train_loader``, its batch size kept) and reused for every sample -- the reference re-iterates
        the loader per sample (util.py:233), i.e. sees a fresh shuffle / augmentation draw each time; the estimator is the
        same.  Returns False when the model / loader shape is not covered (the caller then runs the PyTorch pass)."""
        from ..tasks._engine import _arch_of
        arch = _arch_of(self.swag_model)
        batch = getattr(self.train_loader, "batch_size", None)
        if arch is None or arch[0] != "wrn" or not batch or getattr(self.train_loader, "drop_last", False):
            return False
        _, depth, widen, C = arch
        if _C.lib().ursa_wrn_bn_update_workspace(max(len(self.train_loader.dataset), 1), batch, depth, widen, C) == 0:
            return False
        if getattr(self, "_bn_x", None) is None:
            xs = [xb for xb, _ in self.train_loader]
            if any(len(xb) != batch for xb in xs[:-1]) or xs[0].dim() != 4 or tuple(xs[0].shape[1:]) != (3, 32, 32):
                return False
            self._bn_x = torch.cat(xs).to(self.device, non_blocking=True).float().contiguous()
            self._bn_ws = None
        # FP16-split conv engine first (1.3x the 3xTF32 one, closer to an fp64 pass); statistics that are not finite mean an
        # activation left fp16's range: the pass is repeated on the 3xTF32 engine (one flag read per 50 000-image pass)
        self._bn_ws = _C.wrn_bn_update(self.bank.w[row], self.bank.b[row], self._bn_x, batch, depth, widen, C,
                                       workspace=self._bn_ws, algo=_C.ALGO_TCGEN05_F16)
        if self._bn_ws is not None and not bool(torch.isfinite(self.bank.b[row]).all()):
            self._bn_ws = _C.wrn_bn_update(self.bank.w[row], self.bank.b[row], self._bn_x, batch, depth, widen, C,
                                           workspace=self._bn_ws, algo=_C.ALGO_TCGEN05)
        return self._bn_ws is not None

    def _engine_bn_update_rows(self, rows):
        """PreResNets: ONE sample-batched train-mode pass for all ``rows`` (``ursa_preresnet_bn_update``, 8 samples per launch
        on the tcgen05 layer kernel with a statistics epilogue) instead of one PyTorch pass over the train set per sample
        (util.py:212-247, swag.py:123-124).  Returns False when the model / loader shape is not covered."""
        from ..tasks._engine import _arch_of
        arch = _arch_of(self.swag_model)
        batch = getattr(self.train_loader, "batch_size", None)
        if arch is None or arch[0] != "preresnet" or not batch or getattr(self.train_loader, "drop_last", False) or not rows:
            return False
        _, depth, C = arch
        n_train = max(len(self.train_loader.dataset), 1)
        if _C.lib().ursa_preresnet_bn_update_workspace(len(rows), n_train, batch, depth, C) == 0:
            return False
        if getattr(self, "_bn_x", None) is None:
            xs = [xb for xb, _ in self.train_loader]
            if any(len(xb) != batch for xb in xs[:-1]) or xs[0].dim() != 4 or tuple(xs[0].shape[1:]) != (3, 32, 32):
                return False
            self._bn_x = torch.cat(xs).to(self.device, non_blocking=True).float().contiguous()
            self._bn_ws = None
        w, b = self.bank.rows(list(rows))
        contiguous = b.data_ptr() == self.bank.b[rows[0]].data_ptr()        # a view of the bank: written in place
        self._bn_ws = _C.preresnet_bn_update(w, b, self._bn_x, batch, depth, C, workspace=self._bn_ws)
        if self._bn_ws is None:
            return False
        if not contiguous:
            for i, r in enumerate(rows):
                self.bank.b[r].copy_(b[i])
        return True

    def _finish_sample(self, row, update_bn):
        """Load the draw into ``swag_model``, re-estimate BatchNorm statistics (reference :99-102,:123-124 -- one full
        pass over the train set per sample) and store them with the row."""
        if update_bn and check_bn(self.swag_model):
            if self._engine_bn_update(row) or self._engine_bn_update_rows([row]):
                return self.bank.handle(row)
            self.swag_flat.load_vector(self.bank.w[row])
            bn_update(self.train_loader, self.swag_model, device=self.device)
        self.bank.set_buffers(row, self.swag_flat.b)
        return self.bank.handle(row)

    def sample_iterative(self, update_bn=True, val_loader=None, debug_val_loss=False, wandb_debug=False,
                         full_cov=False):
        if self.burnt_in is False:
            self._fit_or_receive(val_loader, debug_val_loss, wandb_debug)
        row = self._draw_into_bank(1, full_cov)[0]
        return self._finish_sample(row, update_bn)

    def sample(self, num_samples=None, val_loader=None, debug_val_loss=False, wandb_debug=False, full_cov=False):
        if num_samples is None:
            num_samples = self.num_samples
        if self.burnt_in is False:
            self._fit_or_receive(val_loader, debug_val_loss, wandb_debug)
        if self._sharded():
            rank, world = udist.rank_world()
            num_samples = len(range(rank, num_samples, world))      # this rank's draws s = rank (mod world)
        rows = self._draw_into_bank(num_samples, full_cov)     # all draws: one pass over the ring
        if check_bn(self.swag_model) and self._engine_bn_update_rows(rows):
            return [self.bank.handle(r) for r in rows]         # BatchNorm statistics of all draws: one sample-batched pass
        return [self._finish_sample(r, True) for r in rows]
