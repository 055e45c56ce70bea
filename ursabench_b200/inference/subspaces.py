"""Deviation-ring subspace for SWAG (reference inference/subspaces.py:17-43, 71-100).

``CovarianceSpace`` keeps the most recent ``max_rank`` deviation vectors.  The reference re-allocates a K x D CPU
matrix with ``torch.cat`` on every collect (:85-89); here the ring is one preallocated ``[max_rank, ld]`` device
buffer and the K2a kernel writes the new row in place (slot = collects mod max_rank).  ``cov_mat_sqrt`` presents the
rows oldest-first like the reference.  PCA / FreqDir / Random ``get_space`` (CPU SVDs used only by PCA-ESS) are out
of scope for this engine.
"""
import torch

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
    """The reference's default ``subspace_type`` (swa.py:43-46).  Collection is the inherited ring; ``get_space`` is a
    CPU sklearn SVD used only by the PCA-ESS sampler (subspaces.py:116-156) and is out of scope here."""

    def __init__(self, num_parameters, pca_rank=20, max_rank=20, device=None):
        super().__init__(num_parameters, max_rank=max_rank, device=device)
        assert pca_rank == "mle" or isinstance(pca_rank, int)
        if pca_rank != "mle":
            assert 1 <= pca_rank <= max_rank
        self.pca_rank = pca_rank

    def get_space(self):
        raise NotImplementedError("PCASpace.get_space (randomized SVD for PCA-ESS) is outside this engine's hot path")
