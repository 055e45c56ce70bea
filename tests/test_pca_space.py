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
