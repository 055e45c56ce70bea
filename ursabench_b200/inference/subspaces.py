"""Deviation-ring subspace for SWAG (reference inference/subspaces.py:17-43, 71-100).

``CovarianceSpace`` keeps the most recent ``max_rank`` deviation vectors.  The reference re-allocates a K x D CPU
matrix with ``torch.cat`` on every collect (:85-89); here the ring is one preallocated ``[max_rank, ld]`` device
buffer and the K2a kernel writes the new row in place (slot = collects mod max_rank).  ``cov_mat_sqrt`` presents the
rows oldest-first like the reference.  ``PCASpace.get_space`` runs on the device: Gram matrix of the ring
in one streaming pass (``ursa_swag_gram``), eigen-decomposition of the K x K matrix in fp64, components s V^T = U^T A by
one K2b-shaped pass (``ursa_swag_draw`` with z2 = U^T, var = 0).  ``pca_rank='mle'`` (Minka's criterion, which the reference
reaches through a private scikit-learn <= 0.22 function) is evaluated on the <= 24 eigenvalues on the host in fp64.  FreqDir /
Random spaces are out of scope.
"""
import math

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

    @staticmethod
    def minka_log_evidence(eigs, n_samples, n_features):
        """Laplace-approximated log evidence of the PCA model for every candidate rank 0 .. len(eigs) - 1 (Minka, "Automatic
        choice of dimensionality for PCA", 2000, eq. 30) as the reference evaluates it through scikit-learn <= 0.22's
        ``_assess_dimension_(spectrum, rank, n_samples, n_features)`` (subspaces.py:13,143-146).  ``eigs``: descending
        covariance eigenvalues, fp64 on the host -- at most ``DRAW_MAX_K`` numbers, the ring itself never leaves the device.
        A non-positive argument of a logarithm gives NaN / -inf like the reference's; the caller's nanargmax skips it."""
        lam = torch.as_tensor(eigs, dtype=torch.float64)
        r = lam.numel()
        logn = math.log(n_samples)
        out = torch.full((r,), float("nan"), dtype=torch.float64)
        half_dims = (n_features - torch.arange(r, dtype=torch.float64)) / 2.0
        stiefel = torch.lgamma(half_dims) - math.log(math.pi) * half_dims                 # one term per retained direction
        for k in range(r):
            kept, rest = lam[:k], lam[k:]
            if k == n_features:
                noise, log_noise_term = 1.0, 0.0
            else:
                noise = rest.sum() / (n_features - k)                                     # ML noise variance
                log_noise_term = float(-torch.log(noise) * n_samples * (n_features - k) / 2.0)
            dof = n_features * k - k * (k + 1.0) / 2.0
            smoothed = lam.clone()
            smoothed[k:n_features] = noise
            # log |A_Z|: pairs (i < j), i among the retained directions
            gap = kept[:, None] - lam[None, :]
            inv_gap = 1.0 / smoothed[None, :] - 1.0 / smoothed[:k, None]
            pair = torch.triu(torch.ones(k, r, dtype=torch.bool), diagonal=1)
            hess = (torch.log((gap * inv_gap)[pair]) + logn).sum() if k else torch.zeros((), dtype=torch.float64)
            out[k] = (-k * math.log(2.0) + float(stiefel[:k].sum())
                      - float(torch.log(kept).sum()) * n_samples / 2.0
                      + log_noise_term
                      + math.log(2.0 * math.pi) * (dof + k + 1.0) / 2.0
                      - float(hess) / 2.0
                      - k * logn / 2.0)
        return out

    def select_mle_rank(self, s):
        """Post-selection of the reference's ``'mle'`` branch (subspaces.py:133-151): evidence of every rank 0 .. r - 1 minus
        the 0.5 m log(rows) correction, rank = nanargmax.  Sets ``ll``, ``corrected_ll`` and -- like the reference -- overwrites
        ``pca_rank`` with the chosen integer, so later calls take the integer branch."""
        r, D = int(s.numel()), self.num_parameters
        ll = self.minka_log_evidence(s.double() ** 2, n_samples=max(r, D), n_features=min(r, D))
        ranks = torch.arange(r, dtype=torch.float64)
        correction = 0.5 * (D * ranks - ranks * (ranks + 1) / 2.0) * math.log(r)          # reference :139-140
        corrected = ll - correction
        if bool(torch.isnan(corrected).all()):
            raise ValueError("All-NaN slice encountered")                                # numpy.nanargmax's error
        k = int(torch.where(torch.isnan(corrected), torch.full_like(corrected, -float("inf")), corrected).argmax())
        self.ll, self.corrected_ll = ll.numpy(), corrected.numpy()
        self.pca_rank = k
        print("PCA Rank is: ", k)                                                         # reference :152
        return k

    def get_space(self):
        s, U, rows = self.decompose()
        rank = int(self.rank.item())
        D = self.num_parameters
        if self.pca_rank == "mle":
            k = self.select_mle_rank(s[:max(1, rank)])                                    # reference :123-126: all components
            if k == 0:                                                                    # s[:0, None] * Vt[:0]
                self.singular_values = s[:0]
                return torch.empty(0, D, dtype=torch.float32, device=self.device)
        else:
            k = max(1, min(int(self.pca_rank), rank))                                     # reference :128
        zeros = torch.zeros(self.ld, dtype=torch.float32, device=self.device)
        out = torch.empty(k, self.ld, dtype=torch.float32, device=self.device)
        z2 = U[:, :k].t().contiguous().to(device=self.device, dtype=torch.float32)        # s V^T = U^T A
        _C.swag_draw(out, zeros, zeros, D, ring=rows, z2=z2, rank_div=float(max(1, rank - 1)) ** 0.5)
        self.singular_values = s[:k]
        return out[:, :D]
