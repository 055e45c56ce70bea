"""``HMC``: full-batch Hamiltonian Monte Carlo over flattened weights (reference inference/hmc.py:21-85).

The reference wrapper concatenates the whole training set on the device (:44-50) and hands everything to
``hamiltorch.sample_model`` (:71-75) -- a third-party package that is not vendored in the reference and is not
installed here, so the arithmetic follows its published algorithm (restated in ``oracle/restate.py::hmc_*``;
parity unpinned).  This class keeps the wrapper's constructor, hyper-parameter keys, ``update_hyp`` and
``sample(debug)`` (including the ``samples[burn*L::L]`` thinning of :80) and runs C independent chains at once:

  state      theta, r, saved  [C, ld] fp32 in HBM (chain c = row c)
  gradient   d/dtheta sum_i loss_i  for all chains through ``torch.func.vmap(grad)`` of the module's own forward
             (PyTorch is the autograd plumbing, as on the SG-MCMC path)
  K5 kernels momentum draw (Philox in-register), fused leapfrog kick + drift (+ prior gradient), per-chain energy
             sums in fp64, Metropolis accept / restore with the thinned sample written straight into its bank row

``num_chains`` (default 1 = the reference) and ``chain_chunk`` are optional hyper-parameter keys of this engine; with
``torch.distributed`` initialised every rank runs its own ``num_chains`` chains on disjoint Philox streams and there
is no collective.  Nothing in the loop synchronises with the host.
"""
import copy
import math

import torch

from .. import _C, dist as udist
from ..bank import SampleBank
from ..flat import _round_up
from ..util import reset_model
from .inference_base import _Inference, require_cuda

__all__ = ["HMC", "kept_iterations"]


def _loss_sum(model_loss):
    """Summed negative log-likelihood of hamiltorch's ``model_loss`` options (tau_out multiplies it outside)."""
    F = torch.nn.functional
    if model_loss == "multi_class_linear_output":
        return lambda out, y: F.cross_entropy(out, y.long().view(-1), reduction="sum")
    if model_loss == "multi_class_log_softmax_output":
        return lambda out, y: F.nll_loss(out, y.long().view(-1), reduction="sum")
    if model_loss == "binary_class_linear_output":
        return lambda out, y: F.binary_cross_entropy_with_logits(out, y.to(out.dtype).view_as(out), reduction="sum")
    if model_loss == "regression":
        return lambda out, y: 0.5 * ((out - y.to(out.dtype).view_as(out)) ** 2).sum()
    raise NotImplementedError(model_loss)


def kept_iterations(num_samples, L, burn):
    """Which entries of hamiltorch's returned list survive ``samples[burn*L::L]`` (reference :80).  The list holds the
    initial point followed by L positions per iteration.  Returns (first_iteration, use_first_leapfrog_state):
    burn >= 0 keeps the chain state after iterations burn..num_samples (0 = the initial point); burn < 0 keeps
    list[len + burn*L :: L], i.e. the FIRST leapfrog position of the last -burn trajectories (the default burn = -1
    therefore returns one model -- reproduced as is)."""
    if burn >= 0:
        return burn, False
    if -burn * L > num_samples * L + 1:                         # the slice start clamps to 0
        return 0, False
    return num_samples + burn + 1, True


class HMC(_Inference):
    """Hyperparameters: ``step_size, num_samples, L, tau, burn, mass`` (reference :31-41) [+ ``num_chains``]."""

    def __init__(self, hyperparameters, model=None, train_loader=None, model_loss="multi_class_linear_output",
                 device=torch.device("cpu")):
        super().__init__(hyperparameters, model, train_loader, device)
        if hyperparameters is None:
            hyperparameters = {"step_size": 0.001, "num_samples": 10, "L": 1, "tau": 0.1, "burn": -1, "mass": 1.0}
        if not isinstance(model, torch.nn.Module):
            raise NotImplementedError
        self.device = require_cuda(device, "HMC")
        _C.lib()
        self._read_hyp(hyperparameters)
        self.model_loss = model_loss
        self._nll = _loss_sum(model_loss)
        self.tau_out = 1.0                                    # reference :67 ("For Regression make this a hyperparameter")
        x_train, y_train = [], []
        for data, target in train_loader:                     # reference :44-50: the whole train set lives on the device
            x_train.append(data.clone().to(self.device))
            y_train.append(target.clone().to(self.device))
        self.x = torch.cat(x_train)
        self.y = torch.cat(y_train)
        if next(model.parameters()).device != self.device:
            model.to(self.device)
        self._names = [n for n, _ in model.named_parameters()]
        self._shapes = [tuple(p.shape) for p in model.parameters()]
        self._sizes = [p.numel() for p in model.parameters()]
        self.D = sum(self._sizes)
        self.ld = _round_up(self.D, 4)
        self._skeleton = copy.deepcopy(model).cpu()
        nb = sum(b.numel() for b in model.buffers() if b.dtype == torch.float32)
        self.bank = SampleBank(self.D, nb, self.device, capacity=1, skeleton=self._skeleton)
        self.seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.acceptance_rate = None
        self.kernel_launches = 0
        self._inject = None                                   # tests: (z[n][C, ld], logu[n][C]) replayed instead of Philox

    def _read_hyp(self, h):
        self.step_size = h["step_size"]
        self.num_samples = h["num_samples"]
        self.L = h["L"]
        self.tau = h["tau"]
        self.burn = h["burn"]
        self.mass = h["mass"]
        self.num_chains = int(h.get("num_chains", 1))
        self.chain_chunk = int(h.get("chain_chunk", 0))
        if self.num_chains < 1 or self.L < 1:
            raise ValueError("HMC: num_chains and L must be >= 1")

    def update_hyp(self, hyperparameters):
        self._read_hyp(hyperparameters)
        self.model = reset_model(self.model)
        self.bank = self.bank.fresh()            # earlier sample() handles keep the old rows

    # -- gradient of the data term for all chains ---------------------------------------------------------------------
    def _unflatten(self, row):
        params, off = {}, 0
        for name, shape, n in zip(self._names, self._shapes, self._sizes):
            params[name] = row[off:off + n].reshape(shape)
            off += n
        return params

    def _build_grad_fn(self):
        buffers = {n: b for n, b in self.model.named_buffers()}
        model, x, y, nll = self.model, self.x, self.y, self._nll

        def nll_of(row):
            out = torch.func.functional_call(model, (self._unflatten(row), buffers), (x,))
            return nll(out, y)

        return torch.func.vmap(torch.func.grad_and_value(nll_of))

    def _build_grad_fn_mlp(self, arch):
        """Chain-batched gradient of the summed cross entropy for the 3-layer MLP (models/mlp.py:8-23) written out as GEMMs:
        the shared input makes layers 1 of ALL chains one GEMM ([C*h, in] x [in, N]) forward and one backward, layers 2 / 3
        are batched GEMMs in a feature-major [C, h, N] layout -- no per-chain graph, no transposed copies.  Same arithmetic
        as ``vmap(grad)`` of the module (fp32), 2-3x its throughput on 128 chains x 1000 points."""
        _, in_dim, hid, ncls = arch
        xT = self.x.reshape(self.x.shape[0], -1).t().contiguous()          # [in, N]
        xN = xT.t().contiguous()                                            # [N, in]
        y = self.y.long()
        o_w1, o_b1 = 0, hid * in_dim
        o_w2, o_b2 = o_b1 + hid, o_b1 + hid + hid * hid
        o_w3, o_b3 = o_b2 + hid, o_b2 + hid + ncls * hid

        def fn(rows):
            c = rows.shape[0]
            W1 = rows[:, o_w1:o_b1].reshape(c * hid, in_dim)
            b1 = rows[:, o_b1:o_w2].reshape(c, hid, 1)
            W2 = rows[:, o_w2:o_b2].reshape(c, hid, hid)
            b2 = rows[:, o_b2:o_w3].reshape(c, hid, 1)
            W3 = rows[:, o_w3:o_b3].reshape(c, ncls, hid)
            b3 = rows[:, o_b3:o_b3 + ncls].reshape(c, ncls, 1)
            a1 = torch.relu_(torch.mm(W1, xT).view(c, hid, -1).add_(b1))
            a2 = torch.relu_(torch.baddbmm(b2, W2, a1))
            lo = torch.baddbmm(b3, W3, a2)                                   # [c, C, N]
            logp = torch.log_softmax(lo, dim=1)
            val = -logp.gather(1, y.view(1, 1, -1).expand(c, 1, -1)).sum(dim=(1, 2))
            dlo = logp.exp_()
            dlo.scatter_add_(1, y.view(1, 1, -1).expand(c, 1, -1), torch.full((1, 1, 1), -1.0, device=rows.device).expand(c, 1, y.numel()))
            gr = torch.empty(c, self.D, dtype=torch.float32, device=rows.device)
            gr[:, o_w3:o_b3] = torch.bmm(dlo, a2.transpose(1, 2)).reshape(c, -1)
            gr[:, o_b3:o_b3 + ncls] = dlo.sum(2)
            da2 = torch.bmm(W3.transpose(1, 2), dlo).mul_(a2 > 0)
            gr[:, o_w2:o_b2] = torch.bmm(da2, a1.transpose(1, 2)).reshape(c, -1)
            gr[:, o_b2:o_w3] = da2.sum(2)
            da1 = torch.bmm(W2.transpose(1, 2), da2).mul_(a1 > 0)
            gr[:, o_w1:o_b1] = torch.mm(da1.view(c * hid, -1), xN).view(c, -1)
            gr[:, o_b1:o_w2] = da1.sum(2)
            return gr, val

        return fn

    def _build_grad_fn_mlp_tc(self, arch):
        """The same nine GEMMs on the tcgen05 engine (``ursa_gemm_nt_3xtf32``: 3xTF32 with two-level accumulation, fp32-level
        accuracy) instead of cuBLAS' fp32 SGEMM, in a point-major [C, N, h] layout; the weight-gradient GEMMs contract over the
        data points and read feature-major copies of the activations."""
        _, in_dim, hid, ncls = arch
        X2 = self.x.reshape(self.x.shape[0], -1).contiguous()              # [N, in]
        XT = X2.t().contiguous()                                            # [in, N]
        npts = X2.shape[0]
        y = self.y.long()
        o_w1, o_b1 = 0, hid * in_dim
        o_w2, o_b2 = o_b1 + hid, o_b1 + hid + hid * hid
        o_w3, o_b3 = o_b2 + hid, o_b2 + hid + ncls * hid
        ws = [None]

        def mm(A, B, out_shape, bias=None, relu=False):
            out = torch.empty(out_shape, dtype=torch.float32, device=B.device)
            ws[0] = _C.gemm_nt(A, B, out, bias=bias, relu=relu, workspace=ws[0])
            return out

        def fn(rows):
            c = rows.shape[0]
            W1 = rows[:, o_w1:o_b1].unflatten(1, (hid, in_dim))
            W2 = rows[:, o_w2:o_b2].unflatten(1, (hid, hid))
            W3 = rows[:, o_w3:o_b3].unflatten(1, (ncls, hid))
            b1, b2, b3 = rows[:, o_b1:o_w2], rows[:, o_b2:o_w3], rows[:, o_b3:o_b3 + ncls]
            a1 = mm(X2, W1, (c, npts, hid), b1, True)
            a2 = mm(a1, W2, (c, npts, hid), b2, True)
            lo = mm(a2, W3, (c, npts, ncls), b3)
            logp = torch.log_softmax(lo, dim=2)
            idx = y.view(1, -1, 1).expand(c, -1, 1)
            val = -logp.gather(2, idx).sum(dim=(1, 2))
            dlo = logp.exp_()
            dlo.scatter_add_(2, idx, torch.full((1, 1, 1), -1.0, device=rows.device).expand(c, npts, 1))
            gr = torch.empty(c, self.D, dtype=torch.float32, device=rows.device)
            a2T = a2.transpose(1, 2).contiguous()
            gr[:, o_w3:o_b3] = mm(dlo.transpose(1, 2).contiguous(), a2T, (c, ncls, hid)).flatten(1)
            gr[:, o_b3:o_b3 + ncls] = dlo.sum(1)
            da2 = mm(dlo, W3.transpose(1, 2).contiguous(), (c, npts, hid)).mul_(a2 > 0)
            del a2T
            a1T = a1.transpose(1, 2).contiguous()
            gr[:, o_w2:o_b2] = mm(da2.transpose(1, 2).contiguous(), a1T, (c, hid, hid)).flatten(1)
            gr[:, o_b2:o_w3] = da2.sum(1)
            da1 = mm(da2, W2.transpose(1, 2).contiguous(), (c, npts, hid)).mul_(a1 > 0)
            del a1T
            gr[:, o_w1:o_b1] = mm(XT, da1.transpose(1, 2).contiguous(), (c, in_dim, hid)).transpose(1, 2).flatten(1)
            gr[:, o_b1:o_w2] = da1.sum(1)
            return gr, val

        return fn

    def _build_grad_fn_mlp_fused(self, arch, engine="tf32"):
        """``ursa_hmc_mlp_grad`` / ``ursa_hmc_mlp_grad_f16``: forward, loss and backward of all chains as eight hand-written
        tcgen05 GEMMs whose operands are produced in split form by the previous GEMM's epilogue (csrc/bma_mlp_tc.cu: 3xTF32,
        one tile per CTA; csrc/bma_mlp_f16.cu: 2xFP16-split, persistent).  Writes straight into the [C, ld] gradient / energy
        buffers: ``fn(theta, g, ce)``."""
        _, in_dim, hid, ncls = arch
        x2 = self.x.reshape(self.x.shape[0], -1).contiguous().float()
        y = self.y.long().contiguous()
        ws = [None]

        def fn(theta, g, ce):
            ws[0] = _C.hmc_mlp_grad(theta, x2, y, in_dim, hid, ncls, g, ce, workspace=ws[0], engine=engine)
            if ws[0] is None:
                raise RuntimeError("ursa_hmc_mlp_grad does not cover this MLP shape")

        fn.in_place = True
        return fn

    def _grad(self, theta, g, ce):
        """g[c, :D] = d/dtheta sum_i loss_i(theta_c) ; ce[c] = sum_i loss_i(theta_c)   (fp32 forward/backward)."""
        if getattr(self._grad_fn, "in_place", False):
            self._grad_fn(theta, g, ce)
            return
        C = theta.shape[0]
        chunk = self.chain_chunk if self.chain_chunk > 0 else C
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False           # the energy difference needs an fp32 likelihood
        try:
            for c0 in range(0, C, chunk):
                c1 = min(C, c0 + chunk)
                gr, val = self._grad_fn(theta[c0:c1, :self.D])
                g[c0:c1, :self.D].copy_(gr)
                ce[c0:c1].copy_(val)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32

    def _hamiltonian(self, theta, r, ce, out):
        """H = tau_out * sum loss - log prior + 0.5 * inv_mass * r.r  (float64 [C])."""
        self._sums, self._energy_ws = _C.hmc_energy(theta, r, self.D, out=self._sums, workspace=self._energy_ws)
        self.kernel_launches += 2
        torch.mul(self._sums[0], 0.5 * self.tau, out=out)
        out.add_(0.5 * self.D * math.log(2.0 * math.pi / self.tau))
        out.add_(ce.double(), alpha=self.tau_out)
        out.add_(self._sums[1], alpha=0.5 / self.mass)
        return out

    # -- chain initialisation --------------------------------------------------------------------------------------------
    def _initial_state(self):
        """Chain 0 starts from the model's current weights (hamiltorch.util.flatten(model), reference :69); further
        chains start from independent re-initialisations of the same architecture."""
        C = self.num_chains
        theta = torch.zeros(C, self.ld, dtype=torch.float32, device=self.device)
        theta[0, :self.D].copy_(torch.cat([p.detach().reshape(-1) for p in self.model.parameters()]))
        if C > 1:
            rank, _ = udist.rank_world()
            stage = torch.empty(C - 1, self.D)
            scratch = copy.deepcopy(self._skeleton)
            # reset_parameters() draws from the global CPU generator: fork it (CPU only, devices=[]) so that neither the
            # caller's CPU stream nor any CUDA generator is reseeded
            with torch.random.fork_rng(devices=[]):
                for c in range(1, C):
                    torch.manual_seed((self.seed + 7919 * (rank * C + c)) & 0x7FFFFFFF)
                    for m in scratch.modules():
                        fn = getattr(m, "reset_parameters", None)
                        if fn is not None:
                            fn()
                    stage[c - 1].copy_(torch.cat([p.detach().reshape(-1) for p in scratch.parameters()]))
            theta[1:, :self.D].copy_(stage)
        return theta

    # -- sampling ---------------------------------------------------------------------------------------------------------
    def sample(self, debug=False):
        if not isinstance(self.model, torch.nn.Module):
            raise NotImplementedError
        C, L, ld, dev = self.num_chains, self.L, self.ld, self.device
        eps, inv_mass = float(self.step_size), 1.0 / float(self.mass)
        rank, _ = udist.rank_world()
        chain0 = rank * C
        first_it, use_first = kept_iterations(self.num_samples, L, self.burn)
        n_keep = max(0, self.num_samples - first_it + 1)
        self.bank = self.bank.fresh()            # earlier sample() handles keep the old rows
        self.bank.reserve(max(1, n_keep * C))
        self.model.eval()
        self._grad_fn = self._build_grad_fn()
        from ..tasks._engine import _arch_of
        arch = _arch_of(self.model)
        self.grad_engine = "vmap"
        if arch is not None and arch[0] == "mlp" and self.model_loss == "multi_class_linear_output" \
                and getattr(self, "force_vmap_grad", False) is False and self.D == arch[2] * arch[1] + arch[2] \
                + arch[2] * arch[2] + arch[2] + arch[3] * arch[2] + arch[3]:
            # round 1, ms per HMC iteration at 128 chains x 1000 points, MLP 784-200-200-10, L = 10 (eager loop): vmap(grad)
            # 102.9, fp32 GEMMs on cuBLAS 92.3, the same GEMMs one by one on ursa_gemm_nt_3xtf32 102.8 (split passes and
            # feature-major copies per call).  Round 2: ursa_hmc_mlp_grad keeps every operand in split form between GEMMs
            engine = getattr(self, "grad_engine_request", None)          # None = the default below; tests / benches pin one
            fused_ok = self.chain_chunk <= 0 and _C.lib().ursa_hmc_mlp_grad_workspace(self.num_chains, self.x.shape[0], arch[1],
                                                                                      arch[2], arch[3]) > 0
            if engine is None:
                engine = "mlp_tcgen05_fused_f16" if fused_ok else "mlp_gemm"
            if getattr(self, "force_tc_grad", False):
                engine = "mlp_tcgen05"
            if engine == "mlp_tcgen05_fused_f16":
                # the product path: persistent FP16-split GEMMs.  Its range is fp16's: if the gradient at the initial state is
                # not finite where the 3xTF32 engine's is, this run stays on the 3xTF32 engine
                self._grad_fn = self._build_grad_fn_mlp_fused(arch, "f16")
                self._f16_probe = arch
            elif engine == "mlp_tcgen05_fused":
                self._grad_fn = self._build_grad_fn_mlp_fused(arch)      # the same dataflow on the 3xTF32 kernel
            elif engine == "mlp_tcgen05":
                self._grad_fn = self._build_grad_fn_mlp_tc(arch)         # the same GEMMs one by one through ursa_gemm_nt_3xtf32
            else:
                self._grad_fn = self._build_grad_fn_mlp(arch)            # fp32 GEMM formulation on cuBLAS (comparison engine)
            self.grad_engine = engine
        self._sums, self._energy_ws = None, None
        with torch.no_grad():
            theta = self._initial_state()
            saved = theta.clone()
            r = torch.zeros_like(theta)
            g = torch.zeros_like(theta)
            first_cand = torch.zeros_like(theta) if use_first else None
            first_kept = theta.clone() if use_first else None
            ce = torch.zeros(C, dtype=torch.float32, device=dev)
            h_old = torch.zeros(C, dtype=torch.float64, device=dev)
            h_new = torch.zeros(C, dtype=torch.float64, device=dev)
            accept = torch.zeros(C, dtype=torch.int32, device=dev)
            n_accept = torch.zeros(C, dtype=torch.int64, device=dev)
            rows = []
            if self.grad_engine == "mlp_tcgen05_fused_f16":
                self._grad(theta, g, ce)
                if not (bool(torch.isfinite(ce).all()) and bool(torch.isfinite(g).all())):
                    tf32_fn = self._build_grad_fn_mlp_fused(self._f16_probe)
                    tf32_fn(theta, g, ce)
                    if bool(torch.isfinite(ce).all()):                # fp16 range, not a diverged chain: use fp32's range
                        self._grad_fn, self.grad_engine = tf32_fn, "mlp_tcgen05_fused"

            def keep_rows():
                lo = self.bank.count
                self.bank.count += C
                rows.extend(range(lo, lo + C))
                return self.bank.w[lo:lo + C]

            if first_it == 0 and not use_first:
                keep_rows().copy_(theta)                        # samples[0] = the initial point
            def trajectory():
                """Everything between the momentum draw and the accept step: H_old, L leapfrog steps with their gradients,
                H_new.  All arguments are constants of the run, so the whole trajectory (~40 launches per gradient) is ONE
                CUDA graph replay per iteration -- the Python loop was host-bound once eight ranks share the host cores
                (round 1: 100 -> 116 ms per iteration from 1 to 8 GPUs with no collective anywhere)."""
                self._grad(theta, g, ce)
                self._hamiltonian(theta, r, ce, h_old)
                for step in range(L):
                    _C.hmc_leapfrog(theta, r, g, kick=0.5 * eps if step == 0 else eps, drift=eps * inv_mass,
                                    tau=self.tau, tau_out=self.tau_out,
                                    snapshot=first_cand if (use_first and step == 0) else None)
                    self._grad(theta, g, ce)
                _C.hmc_leapfrog(theta, r, g, kick=0.5 * eps, drift=0.0, tau=self.tau, tau_out=self.tau_out)
                self._hamiltonian(theta, r, ce, h_new)

            graph = None
            want_graph = bool(getattr(self, "use_cuda_graph", True)) and self.grad_engine != "vmap" and self.num_samples >= 3
            self.graph_replays = 0
            ev_mid, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for n in range(1, self.num_samples + 1):
                if n == 3:
                    ev_mid.record()                             # iterations 1-2 warm up / capture: steady state from here
                inj = self._inject
                _C.hmc_momentum(r, math.sqrt(self.mass), noise=None if inj is None else inj[0][n - 1], seed=self.seed,
                                step=n, elem_offset=chain0 * ld)
                if graph is not None:
                    graph.replay()
                    self.graph_replays += 1
                elif want_graph and n == 2:
                    # iteration 1 ran eagerly (workspaces, cuBLAS handles and heuristics are warm); capture this one
                    torch.cuda.synchronize(dev)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        trajectory()
                    graph.replay()                              # capture does not execute: run iteration 2
                    self.graph_replays += 1
                else:
                    trajectory()
                out = keep_rows() if n >= max(first_it, 1) else None
                _C.hmc_accept(theta, saved, h_old, h_new, accept, logu=None if inj is None else inj[1][n - 1],
                              keep_dst=first_kept, keep_src=first_cand, out=out, seed=self.seed ^ 0x9E3779B97F4A7C15,
                              step=n, chain_offset=chain0)
                self.kernel_launches += L + 5
                n_accept += accept
                if debug:
                    print("HMC iteration %d: accepted %d / %d chains, H_old[0] = %.4f, H_new[0] = %.4f"
                          % (n, int(accept.sum().item()), C, float(h_old[0].item()), float(h_new[0].item())))
            self.ms_per_iteration = None                        # device time per iteration, iterations 3 .. num_samples
            if self.num_samples >= 3:
                ev_end.record()
                ev_end.synchronize()
                self.ms_per_iteration = ev_mid.elapsed_time(ev_end) / (self.num_samples - 2)
            self.acceptance_rate = (n_accept.double() / max(1, self.num_samples)).cpu()
            self.theta = theta
        # leave the live model on chain 0's final state, like hamiltorch leaves `params`
        off = 0
        for p, nel in zip(self.model.parameters(), self._sizes):
            p.data.copy_(theta[0, off:off + nel].view(p.shape))
            off += nel
        buf = [b.detach().reshape(-1) for b in self.model.buffers() if b.dtype == torch.float32]
        if buf:
            flat_b = torch.cat(buf)
            for i in rows:
                self.bank.set_buffers(i, flat_b)
        return [self.bank.handle(i) for i in rows]
