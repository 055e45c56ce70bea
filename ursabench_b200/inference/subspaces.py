"""Deviation-ring subspace for SWAG (reference inference/subspaces.py:17-43, 71-100).

``CovarianceSpace`` keeps the most recent ``max_rank`` deviation vectors.  The reference re-allocates a K x D CPU
matrix with ``torch.cat`` on every collect (:85-89); here the ring is one preallocated ``[max_rank, ld]`` device
buffer and the K2a kernel writes the new row in place (slot = collects mod max_rank).  ``cov_mat_sqrt`` presents the
rows oldest-first like the reference.  ``PCASpace.get_space`` runs on the device: Gram matrix of the ring
in one streaming pass (``ursa_swag_gram``), eigen-decomposition of the K x K matrix in fp64, components s V^T = U^T A by
one K2b-shaped pass (``ursa_swag_draw`` with z2 = U^T, var = 0).  FreqDir / Random spaces and the ``'mle'`` rank selection
(Minka's criterion through a private sklearn function) are out of scope.
"""
import torch

from .. import _C

from ..flat import _round_up


class Subspace(torch.nn.Module):
    subclasses = {}

    @classmethod
    def register_subclass(cls, subspace_type):
        def decorator(subclass):
            cls.subclasses[subspace_type] = subclass
            return subclass
        return decorator

    @classmethod
    def create(cls, subspace_type, **kwargs):
        if subspace_type not in cls.subclasses:
            raise ValueError("Bad subspaces type {}".format(subspace_type))
        return cls.subclasses[subspace_type](**kwargs)

    def collect_vector(self, vector):
        raise NotImplementedError

    def get_space(self):
        raise NotImplementedError


@Subspace.register_subclass("covariance")
class CovarianceSpace(Subspace):
    def __init__(self, num_parameters, max_rank=20, device=None):
        super().__init__()
        self.num_parameters = num_parameters
        self.max_rank = max_rank
        self.ld = _round_up(num_parameters, 4)
        self.device = torch.device("cuda" if device is None else device)
        self.ring = torch.zeros(max_rank, self.ld, dtype=torch.float32, device=self.device)
        self.collected = 0                                   # total rows ever written
        self.register_buffer("rank", torch.zeros(1, dtype=torch.long))

    def next_slot(self):
        """Row the next deviation goes to (written by ``ursa_swag_collect``); call ``commit()`` after the launch."""
        return self.ring[self.collected % self.max_rank]

    def commit(self):
        self.collected += 1
        self.rank = torch.clamp(self.rank + 1, max=self.max_rank).view(-1)      # reference :89

    def collect_vector(self, vector):
        """API-compatible path: copy a deviation vector into the ring (reference :85-89)."""
        self.next_slot()[:self.num_parameters].copy_(vector.view(-1)[:self.num_parameters])
        self.commit()

    def rows(self):
        """[r, ld] device view/copy of the valid rows in ring order (order is irrelevant to the draw)."""
        return self.ring[:min(self.collected, self.max_rank)]

    @property
    def cov_mat_sqrt(self):
        r = min(self.collected, self.max_rank)
        if self.collected <= self.max_rank:
            return self.ring[:r, :self.num_parameters]
        start = self.collected % self.max_rank                # oldest row
        return torch.roll(self.ring, -start, dims=0)[:, :self.num_parameters]

    def get_space(self):
        m = self.cov_mat_sqrt
        return m.clone() / (m.size(0) - 1) ** 0.5             # reference :91-92


@Subspace.register_subclass("pca")
class PCASpace(CovarianceSpace):
    """The reference's default ``subspace_type`` (swa.py:43-46).  Collection is the inherited ring; ``get_space`` returns
    ``s[:k, None] * Vt[:k]`` of A = ring / sqrt(max(1, rank - 1)) like the reference (subspaces.py:116-131,154-156), as a
    device tensor.  The reference's randomized SVD (n_iter = 5 on <= 24 rows) converges to the exact SVD and fixes signs
    with sklearn's ``svd_flip`` (largest |u| entry of every left singular vector positive); both are reproduced."""

    def __init__(self, num_parameters, pca_rank=20, max_rank=20, device=None):
        super().__init__(num_parameters, max_rank=max_rank, device=device)
        assert pca_rank == "mle" or isinstance(pca_rank, int)
        if pca_rank != "mle":
            assert 1 <= pca_rank <= max_rank
        self.pca_rank = pca_rank

    def decompose(self):
        """(s [r] float64, U [r, r] float64 with svd_flip signs, rows [r, ld]) of A = rows / sqrt(max(1, rank - 1))."""
        rows = self.rows()
        r = rows.shape[0]
        if r == 0:
            raise RuntimeError("PCASpace: no vectors collected")
        if r > _C.DRAW_MAX_K:
            raise NotImplementedError("PCASpace.get_space on the device covers max_rank <= %d" % _C.DRAW_MAX_K)
        rank = int(self.rank.item())
        gram = _C.swag_gram(rows, self.num_parameters) / float(max(1, rank - 1))          # A A^T, fp64
        lam, U = torch.linalg.eigh(gram.cpu())                                            # r x r: host LAPACK is the right tool
        order = torch.argsort(lam, descending=True)
        lam, U = lam[order].clamp_min(0.0), U[:, order]
        idx = U.abs().argmax(dim=0)                                                       # sklearn svd_flip, u-based
        sgn = torch.sign(U[idx, torch.arange(r)])
        sgn[sgn == 0] = 1.0
        return lam.sqrt(), U * sgn[None, :], rows

    def get_space(self):
        if self.pca_rank == "mle":
            raise NotImplementedError("PCASpace(pca_rank='mle'): Minka's rank selection (sklearn's private "
                                      "_assess_dimension_) is outside this engine's hot path")
        s, U, rows = self.decompose()
        rank = int(self.rank.item())
        k = max(1, min(int(self.pca_rank), rank))                                         # reference :128
        D = self.num_parameters
        zeros = torch.zeros(self.ld, dtype=torch.float32, device=self.device)
        out = torch.empty(k, self.ld, dtype=torch.float32, device=self.device)
        z2 = U[:, :k].t().contiguous().to(device=self.device, dtype=torch.float32)        # s V^T = U^T A
        _C.swag_draw(out, zeros, zeros, D, ring=rows, z2=z2, rank_div=float(max(1, rank - 1)) ** 0.5)
        self.singular_values = s[:k]
        return out[:, :D]
