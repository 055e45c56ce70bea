"""Subspace projection (reference inference/projection_model.py:6-14): ``theta = mean + cov_factor^T t``.

The reference evaluates this with a dense ``[D, r] @ [r]`` product on the host for every proposal of the subspace
samplers; here it is one launch of the K2b kernel (``ursa_swag_draw`` with z2 = t, var = 0): the ``[r, D]`` factor is
staged tile by tile by the TMA engine and contracted on the warp-level tensor-core MMA.
"""
import torch

from .. import _C
from ..flat import _round_up


class SubspaceModel(torch.nn.Module):
    def __init__(self, mean, cov_factor):
        super().__init__()
        if not cov_factor.is_cuda:
            raise ValueError("SubspaceModel needs CUDA tensors: this engine has no CPU path")
        self.rank = cov_factor.size(0)
        if self.rank > _C.DRAW_MAX_K:
            raise NotImplementedError("SubspaceModel covers rank <= %d" % _C.DRAW_MAX_K)
        D = cov_factor.size(1)
        ld = _round_up(D, 4)
        self.num_parameters = D
        dev = cov_factor.device
        m = torch.zeros(ld, dtype=torch.float32, device=dev)
        m[:D].copy_(mean.view(-1))
        f = torch.zeros(self.rank, ld, dtype=torch.float32, device=dev)
        f[:, :D].copy_(cov_factor)
        self.register_buffer("_mean_pad", m)
        self.register_buffer("_factor_pad", f)
        self.register_buffer("_zero_var", torch.zeros(ld, dtype=torch.float32, device=dev))

    @property
    def mean(self):
        return self._mean_pad[:self.num_parameters]

    @property
    def cov_factor(self):
        return self._factor_pad[:, :self.num_parameters]

    def forward(self, t):
        """t: [rank] -> theta [D]; t: [S, rank] (S <= 32 proposals) -> [S, D] in the same single pass over the factor."""
        single = t.dim() == 1
        z2 = t.reshape(-1, self.rank).to(device=self._mean_pad.device, dtype=torch.float32).contiguous()
        out = torch.empty(z2.shape[0], self._mean_pad.numel(), dtype=torch.float32, device=self._mean_pad.device)
        _C.swag_draw(out, self._mean_pad, self._zero_var, self.num_parameters, ring=self._factor_pad, z2=z2, rank_div=1.0)
        out = out[:, :self.num_parameters]
        return out[0] if single else out
