"""PCA subspace: PCASpace.get_space and SubspaceModel.forward.

CPU: the numpy restatement (oracle/restate.py::pca_space, subspace_project) against goldens produced by the live reference
(sklearn randomized_svd + svd_flip, inference/subspaces.py:116-156, inference/projection_model.py:13-14).
GPU: ursa_swag_gram + eigh + one K2b pass against the same goldens and the oracle, ragged D, ring wrap-around."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pca_space.npz")


@pytest.mark.parametrize("tag", ["wrap", "short"])
def test_oracle_pca_space_matches_reference_golden(tag):
    g = np.load(GOLD)
    D, max_rank, pca_rank, ncollect = (int(v) for v in g[tag + "/cfg"])
    space = R.pca_space(g[tag + "/ring"], int(g[tag + "/rank"][0]), pca_rank)
    ref = g[tag + "/space"]
    assert space.shape == ref.shape == (max(1, min(pca_rank, min(ncollect, max_rank))), D)
    np.testing.assert_allclose(space, ref, atol=2e-5 * np.abs(ref).max(), rtol=1e-4)
    proj = R.subspace_project(g[tag + "/mean"], ref, g[tag + "/t"])
    np.testing.assert_allclose(proj, g[tag + "/projected"], atol=1e-5, rtol=1e-5)


MLE_TAGS = ["mle3", "mle5", "mle0"]


def test_restated_assess_dimension_matches_installed_sklearn():
    """Pin of oracle/restate.py::assess_dimension_sklearn022 (the <= 0.22 four-argument function the reference imports): the
    installed scikit-learn's three-argument ``_assess_dimension`` is the same formula with n_features = len(spectrum), up to
    one rank-independent constant."""
    _pca = pytest.importorskip("sklearn.decomposition._pca")
    rng = np.random.RandomState(0)
    for n in (4, 9, 20):
        spectrum = np.sort(rng.gamma(2.0, 1.0, n))[::-1].copy()
        for n_samples in (n, 50, 36546980):
            for rank in range(1, n):
                want = _pca._assess_dimension(spectrum, rank, n_samples)
                got = R.assess_dimension_sklearn022(spectrum, rank, n_samples, n)
                # 0.23 changed the prior-volume exponent from (m + rank + 1) / 2 to (m + rank) / 2: a rank-independent
                # constant, log(2 pi) / 2, which no argmax over ranks sees
                assert got - 0.5 * np.log(2.0 * np.pi) == pytest.approx(want, rel=1e-12, abs=1e-9), (n, n_samples, rank)


@pytest.mark.parametrize("tag", MLE_TAGS)
def test_oracle_and_host_mle_rank_match_reference_golden(tag):
    """pca_rank='mle' (subspaces.py:123-153): golden = the live reference's branch; the oracle restatement and the product's
    host-side criterion (PCASpace.minka_log_evidence, no device work) both reproduce its evidence curve and chosen rank."""
    from ursabench_b200.inference.subspaces import PCASpace
    g = np.load(GOLD)
    D, max_rank, ncollect = (int(v) for v in g[tag + "/cfg"])
    rank = int(g[tag + "/rank"][0])
    space, k, ll, corrected = R.pca_space_mle(g[tag + "/ring"], rank)
    assert k == int(g[tag + "/chosen"][0]) and space.shape == g[tag + "/space"].shape == (k, D)
    np.testing.assert_allclose(ll, g[tag + "/ll"], rtol=1e-4)                    # the reference's spectrum is float32
    np.testing.assert_allclose(corrected, g[tag + "/corrected_ll"], rtol=1e-4)
    if k:
        np.testing.assert_allclose(space, g[tag + "/space"], atol=2e-5 * np.abs(g[tag + "/space"]).max(), rtol=1e-4)
    A = g[tag + "/ring"].astype(np.float64) / max(1, rank - 1) ** 0.5
    eigs = np.linalg.svd(A, compute_uv=False) ** 2
    host = PCASpace.minka_log_evidence(eigs, n_samples=max(A.shape), n_features=min(A.shape)).numpy()
    np.testing.assert_allclose(host, ll, rtol=1e-7)            # eigenvalues from two SVD routes
    np.testing.assert_allclose(host, g[tag + "/ll"], rtol=1e-4)


def test_host_mle_evidence_degenerate_spectrum_is_nan_not_an_error():
    """Equal or zero eigenvalues put a non-positive number under a logarithm: the reference gets NaN / -inf there and
    nanargmax skips it (subspaces.py:151)."""
    from ursabench_b200.inference.subspaces import PCASpace
    eigs = np.array([4.0, 4.0, 1.0, 0.0])
    host = PCASpace.minka_log_evidence(eigs, n_samples=100, n_features=4).numpy()
    with np.errstate(all="ignore"):
        want = np.array([R.assess_dimension_sklearn022(eigs, k, 100, 4) for k in range(4)])
    assert np.isfinite(host[0]) and host[0] == pytest.approx(want[0], rel=1e-12)
    assert np.array_equal(np.isnan(host), np.isnan(want))
    assert np.array_equal(np.isposinf(host), np.isposinf(want)) and np.array_equal(np.isneginf(host), np.isneginf(want))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", MLE_TAGS)
def test_pca_space_mle_on_device_matches_reference_golden(tag, capsys):
    from ursabench_b200.inference import PCASpace
    g = np.load(GOLD)
    D, max_rank, ncollect = (int(v) for v in g[tag + "/cfg"])
    sp = PCASpace(num_parameters=D, pca_rank="mle", max_rank=max_rank, device="cuda")
    for v in g[tag + "/vecs"]:
        sp.collect_vector(torch.from_numpy(v).cuda())
    space = sp.get_space()
    ref = g[tag + "/space"]
    assert "PCA Rank is" in capsys.readouterr().out                                          # reference :152
    assert sp.pca_rank == int(g[tag + "/chosen"][0]) and tuple(space.shape) == ref.shape and space.is_cuda
    np.testing.assert_allclose(sp.ll, g[tag + "/ll"], rtol=1e-4)
    np.testing.assert_allclose(sp.corrected_ll, g[tag + "/corrected_ll"], rtol=1e-4)
    if ref.shape[0]:
        np.testing.assert_allclose(space.cpu().numpy(), ref, atol=5e-5 * np.abs(ref).max(), rtol=1e-3)
    again = sp.get_space()                                # pca_rank is now the chosen integer (reference :151): integer branch
    assert again.shape[0] == max(1, ref.shape[0])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["wrap", "short"])
def test_pca_space_on_device_matches_reference_golden(tag):
    from ursabench_b200.inference import PCASpace, SubspaceModel
    g = np.load(GOLD)
    D, max_rank, pca_rank, ncollect = (int(v) for v in g[tag + "/cfg"])
    sp = PCASpace(num_parameters=D, pca_rank=pca_rank, max_rank=max_rank, device="cuda")
    for v in g[tag + "/vecs"]:
        sp.collect_vector(torch.from_numpy(v).cuda())
    assert int(sp.rank.item()) == int(g[tag + "/rank"][0])
    np.testing.assert_array_equal(sp.cov_mat_sqrt.cpu().numpy(), g[tag + "/ring"])          # ring order = reference order
    space = sp.get_space()
    ref = g[tag + "/space"]
    assert tuple(space.shape) == ref.shape
    np.testing.assert_allclose(space.cpu().numpy(), ref, atol=5e-5 * np.abs(ref).max(), rtol=1e-3)
    model = SubspaceModel(torch.from_numpy(g[tag + "/mean"]).cuda(), torch.from_numpy(ref).cuda())
    out = model(torch.from_numpy(g[tag + "/t"]).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), g[tag + "/projected"], atol=2e-5, rtol=2e-5)
    many = model(torch.from_numpy(np.stack([g[tag + "/t"], 2 * g[tag + "/t"]])).cuda())     # batched proposals, one pass
    np.testing.assert_allclose(many[0].cpu().numpy(), g[tag + "/projected"], atol=2e-5, rtol=2e-5)
    np.testing.assert_allclose((many[1] - model.mean).cpu().numpy(), 2 * (g[tag + "/projected"] - g[tag + "/mean"]), atol=4e-5, rtol=4e-5)
    # the reference's state_dict keys / shapes, and its differentiable expression when t requires grad (projection_model.py:13-14)
    sd = model.state_dict()
    assert set(sd) == {"mean", "cov_factor"} and tuple(sd["cov_factor"].shape) == tuple(ref.shape)
    t = torch.from_numpy(g[tag + "/t"]).cuda().requires_grad_(True)
    out_g = model(t)
    np.testing.assert_allclose(out_g.detach().cpu().numpy(), g[tag + "/projected"], atol=2e-5, rtol=2e-5)
    out_g.sum().backward()
    np.testing.assert_allclose(t.grad.cpu().numpy(), ref.sum(1), atol=1e-4, rtol=1e-4)
    # buffers changed in place (load_state_dict): the padded copies the kernel reads follow
    model.load_state_dict({"mean": sd["mean"] * 0 + 1.0, "cov_factor": sd["cov_factor"]})
    out2 = model(torch.from_numpy(g[tag + "/t"]).cuda())
    np.testing.assert_allclose(out2.cpu().numpy(), g[tag + "/projected"] - g[tag + "/mean"] + 1.0, atol=4e-5, rtol=4e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("K,D", [(1, 5), (7, 1001), (20, 272282), (24, 4099)])
def test_swag_gram_matches_fp64_numpy(K, D):
    from ursabench_b200 import _C
    rng = np.random.RandomState(K + D)
    ld = (D + 3) // 4 * 4
    ring = np.zeros((K, ld), np.float32)
    ring[:, :D] = rng.randn(K, D).astype(np.float32) * np.linspace(1.0, 0.01, K)[:, None].astype(np.float32)
    ring[:, D:] = 7.0                                                       # padding must not leak into the Gram matrix
    gram = _C.swag_gram(torch.from_numpy(ring).cuda(), D).cpu().numpy()
    ref = ring[:, :D].astype(np.float64) @ ring[:, :D].astype(np.float64).T
    np.testing.assert_allclose(gram, ref, rtol=2e-5, atol=2e-6 * np.abs(ref).max())
    np.testing.assert_array_equal(gram, gram.T)
    # fixed reduction order: a second call returns the same bits
    assert np.array_equal(gram, _C.swag_gram(torch.from_numpy(ring).cuda(), D).cpu().numpy())
    # rows that are not 16-byte aligned take the CUDA-core path (the tensor-core kernel streams rows with bulk copies)
    if D > 1:
        buf = torch.empty(K * ld + 1, device="cuda")
        odd = buf[1:].view(K, ld)                                          # contiguous, base 4 bytes off a 16-byte boundary
        odd.copy_(torch.from_numpy(ring))
        g2 = _C.swag_gram(odd, D).cpu().numpy()
        np.testing.assert_allclose(g2, ref, rtol=2e-5, atol=2e-6 * np.abs(ref).max())
        np.testing.assert_array_equal(g2, g2.T)


@pytest.mark.gpu
def test_pca_space_properties_at_preresnet_size():
    """D = 272 282, K = 20: components are orthogonal with squared norms = eigenvalues of A A^T, and P^T P reproduces A^T A on
    random probes (the definition of the PCA factor), against an fp64 SVD on the host."""
    from ursabench_b200.inference import PCASpace
    D, K = 272_282, 20
    torch.manual_seed(0)
    sp = PCASpace(num_parameters=D, pca_rank=K, max_rank=K, device="cuda")
    scales = torch.linspace(1.0, 0.05, K)
    for i in range(K):
        sp.collect_vector(torch.randn(D, device="cuda") * scales[i])
    P = sp.get_space().double()
    A = sp.cov_mat_sqrt.double() / (K - 1) ** 0.5
    sv = torch.linalg.svdvals(A.cpu())
    G = (P @ P.t()).cpu()
    assert (G - torch.diag(sv ** 2)).abs().max().item() < 2e-4 * (sv[0] ** 2).item()
    probe = torch.randn(D, 3, device="cuda", dtype=torch.float64)
    lhs, rhs = P.t() @ (P @ probe), A.t() @ (A @ probe)
    assert (lhs - rhs).abs().max().item() < 2e-4 * rhs.abs().max().item()
