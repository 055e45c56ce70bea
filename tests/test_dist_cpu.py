"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding of independent units and the single BMA
all-reduce.  The device kernels are not involved; this covers ursabench_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ursabench_b200 import dist as udist


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 100, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [udist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_pairs_covers_the_grid_once_and_balances():
    """Hybrid (sample, image-block) sharding: every pair exactly once, shares within a few 64-image blocks of each other,
    whole samples when S divides over the ranks (so the bank rows stay where the chains ran)."""
    for S, N, world in ((10, 10_000, 8), (100, 10_000, 8), (3, 1000, 4), (1, 10, 4), (7, 130, 2), (100, 10_000, 4), (5, 63, 3)):
        seen = np.zeros((S, N), dtype=np.int32)
        sizes = []
        for r in range(world):
            sh = udist.shard_pairs(S, N, r, world)
            sizes.append(sum(hi - lo for _, lo, hi in sh))
            for s, lo, hi in sh:
                assert 0 <= lo < hi <= N
                seen[s, lo:hi] += 1
            if S % world == 0:
                assert all(lo == 0 and hi == N for _, lo, hi in sh) and len(sh) == S // world
        assert (seen == 1).all()
        if S * ((N + 63) // 64) >= world:
            assert max(sizes) - min(sizes) <= 6 * 64, (S, N, world, sizes)          # 1 block + 2 snapped blocks at either end
        if (S, N, world) == (100, 10_000, 8):
            # no 64- / 128-image stubs: every rank runs whole samples plus at most two half-sample pieces
            for r in range(world):
                sh = udist.shard_pairs(S, N, r, world)
                assert all(hi - lo == N or hi - lo > 1000 for _, lo, hi in sh), (r, sh)
                assert sum(1 for _, lo, hi in sh if hi - lo != N) <= 2
    assert udist.shard_pairs(0, 10, 0, 2) == [] and udist.shard_pairs(4, 0, 1, 2) == []


def test_shard_pairs_properties_random_shapes():
    """Property check over random (S, N, world, quantum): the shares tile the grid exactly once, in sample-major order
    across ranks, and stay within one block + two snapped blocks per end of the ideal share."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=150, deadline=None)
    @hyp.given(S=st.integers(1, 40), N=st.integers(1, 20_000), world=st.integers(1, 16), quantum=st.sampled_from([16, 64, 100]))
    def check(S, N, world, quantum):
        flat = []
        sizes = []
        for r in range(world):
            sh = udist.shard_pairs(S, N, r, world, quantum=quantum)
            sizes.append(sum(hi - lo for _, lo, hi in sh))
            for s, lo, hi in sh:
                assert 0 <= lo < hi <= N
                flat.append((s, lo, hi))
        # sample-major, contiguous, nothing twice, nothing missing
        pos = 0
        for s, lo, hi in flat:
            assert s * N + lo == pos, (S, N, world, quantum, flat)
            pos = s * N + hi
        assert pos == S * N
        if S % world:
            assert max(sizes) - min(sizes) <= 6 * quantum, (S, N, world, quantum, sizes)
        else:
            assert max(sizes) == min(sizes) == (S // world) * N
    check()


def test_pack_plan_covers_the_list_with_a_small_first_sub_batch(monkeypatch):
    """Host modules handed to a task are packed / uploaded in sub-batches that overlap the device (tasks/_engine.py):
    every module exactly once and in order, 2 first when the list is split, at most 8 per sub-batch, no one-module tail."""
    from ursabench_b200.tasks import _engine
    for family in ("preresnet", "wrn", "mlp"):
        for n in range(0, 70):
            plan = _engine.pack_plan(n, family)
            assert sum(plan) == n and all(k >= 1 for k in plan)
            if len(plan) > 1:
                assert plan[0] == 2 and max(plan) <= 8 and plan[-1] >= 2
            else:
                assert n <= (4 if family != "mlp" else 16)
    assert _engine.pack_plan(12, "preresnet") == [2, 8, 2]          # a rank's share of S = 100 on 8 ranks
    assert _engine.pack_plan(11, "preresnet") == [2, 7, 2]
    assert _engine.pack_plan(12, "mlp") == [12]
    monkeypatch.setattr(_engine, "_PACK_SPLIT_MIN", 0)              # a silly knob value must not loop
    assert _engine.pack_plan(1, "wrn") == [1] and _engine.pack_plan(2, "wrn") == [2] and _engine.pack_plan(3, "wrn") == [3]
    assert _engine.pack_plan(4, "wrn") == [2, 2]


def test_chain_elem_offsets_disjoint_and_aligned():
    D = 272_282
    offs = [udist.chain_elem_offset(c, D) for c in range(8)]
    assert all(o % 4 == 0 for o in offs)
    assert all(b - a >= D for a, b in zip(offs, offs[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert udist.is_distributed() and udist.rank_world() == (rank, world)
        S, N, C = 10, 37, 5
        rng = np.random.RandomState(0)                         # same data on every rank
        p = rng.dirichlet(np.ones(C), size=(S, N)).astype(np.float32)
        e = rng.rand(S, N).astype(np.float32)
        lo, hi = udist.shard_range(S)                          # samples sharded across ranks
        local_p = torch.from_numpy(p[lo:hi].sum(0))
        local_e = torch.from_numpy(e[lo:hi].sum(0))
        before = local_p.clone()
        P, E, cnt = udist.allreduce_bma(local_p, local_e, hi - lo)
        assert torch.equal(local_p, before)                    # inputs untouched
        assert cnt == S
        np.testing.assert_allclose(P.numpy(), p.sum(0), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(E.numpy(), e.sum(0), rtol=1e-6, atol=1e-6)
        # every rank ends with the same reduced tensors -> identical metrics everywhere
        gathered = [torch.zeros_like(P) for _ in range(world)]
        dist.all_gather(gathered, P.contiguous())
        assert all(torch.equal(gathered[0], g) for g in gathered)
        assert udist.allreduce_max_scalar(float(rank + 1), torch.device("cpu")) == float(world)
        # hybrid sharding, S = 3 samples over 2 ranks: each rank sums its (sample, image-range) shares into a full-size
        # accumulator, counts the samples on rank 0 only, and the same single all-reduce completes the evaluation
        S3 = 3
        acc_p, acc_e = torch.zeros(N, C), torch.zeros(N)
        for s_i, lo_i, hi_i in udist.shard_pairs(S3, N, quantum=8):
            acc_p[lo_i:hi_i] += torch.from_numpy(p[s_i, lo_i:hi_i])
            acc_e[lo_i:hi_i] += torch.from_numpy(e[s_i, lo_i:hi_i])
        P3, E3, cnt3 = udist.allreduce_bma(acc_p, acc_e, S3 if rank == 0 else 0)
        assert cnt3 == S3
        np.testing.assert_allclose(P3.numpy(), p[:S3].sum(0), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(E3.numpy(), e[:S3].sum(0), rtol=1e-6, atol=1e-6)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_bma_allreduce_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_single_process_is_identity():
    P, E = torch.rand(4, 3), torch.rand(4)
    P2, E2, n = udist.allreduce_bma(P, E, 7)
    assert P2 is P and E2 is E and n == 7
