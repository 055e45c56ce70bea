"""Subspace projection (reference inference/projection_model.py:6-14): ``theta = mean + cov_factor^T t``.

The reference evaluates this with a dense ``[D, r] @ [r]`` product on the host for every proposal of the subspace
samplers; here it is one launch of the K2b kernel (``ursa_swag_draw`` with z2 = t, var = 0): the ``[r, D]`` factor is
staged tile by tile by the TMA engine and contracted on tcgen05 (3xTF32, fp32 accumulate in TMEM).
"""
import torch

from .. import _C
from ..flat import _round_up


class SubspaceModel(torch.nn.Module):
    """Buffers ``mean`` [D] and ``cov_factor`` [rank, D] as in the reference (same ``state_dict`` keys and shapes).  The kernel
    wants rows padded to a multiple of 4 floats: padded copies are built lazily and rebuilt when the buffers change (in-place
    writes, ``load_state_dict``, ``.to()``).  With ``t.requires_grad`` the forward is the reference's differentiable expression."""

    def __init__(self, mean, cov_factor):
        super().__init__()
        if not cov_factor.is_cuda:
            raise ValueError("SubspaceModel needs CUDA tensors: this engine has no CPU path")
        self.rank = cov_factor.size(0)
        if self.rank > _C.DRAW_MAX_K:
            raise NotImplementedError("SubspaceModel covers rank <= %d" % _C.DRAW_MAX_K)
        self.num_parameters = cov_factor.size(1)
        dev = cov_factor.device
        self.register_buffer("mean", mean.detach().reshape(-1).to(device=dev, dtype=torch.float32).clone())
        self.register_buffer("cov_factor", cov_factor.detach().to(dtype=torch.float32).clone())
        self._pad = None          # (key, mean_pad [ld], factor_pad [rank, ld], zero_var [ld])

    def _padded(self):
        key = (self.mean.data_ptr(), self.mean._version, self.cov_factor.data_ptr(), self.cov_factor._version)
        if self._pad is None or self._pad[0] != key:
            D, dev = self.num_parameters, self.cov_factor.device
            ld = _round_up(D, 4)
            m = torch.zeros(ld, dtype=torch.float32, device=dev)
            m[:D].copy_(self.mean)
            f = torch.zeros(self.rank, ld, dtype=torch.float32, device=dev)
            f[:, :D].copy_(self.cov_factor)
            self._pad = (key, m, f, torch.zeros(ld, dtype=torch.float32, device=dev))
        return self._pad[1:]

    def forward(self, t):
        """t: [rank] -> theta [D]; t: [S, rank] -> [S, D] in the same single pass over the factor (30 proposals per launch)."""
        if torch.is_grad_enabled() and t.requires_grad:
            return self.mean + t.to(self.cov_factor.dtype) @ self.cov_factor      # reference projection_model.py:13-14, differentiable
        mean_pad, factor_pad, zero_var = self._padded()
        single = t.dim() == 1
        z2 = t.detach().reshape(-1, self.rank).to(device=mean_pad.device, dtype=torch.float32).contiguous()
        out = torch.empty(z2.shape[0], mean_pad.numel(), dtype=torch.float32, device=mean_pad.device)
        _C.swag_draw(out, mean_pad, zero_var, self.num_parameters, ring=factor_pad, z2=z2, rank_div=1.0)
        out = out[:, :self.num_parameters]
        return out[0] if single else out
