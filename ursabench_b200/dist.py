"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the box, gloo in
CPU tests).  The hot path shards without any data-path collective -- chains, posterior samples and draws are
independent -- except ONE exchange: the all-reduce(sum) of the summed predictive probabilities [N, C], the summed
entropies [N] and the sample count at the end of a BMA evaluation."""
import os

import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or (dist.is_available() and dist.is_initialized()):
        return rank_world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend)
    return rank_world()


def shard_range(n, rank=None, world=None):
    """Contiguous, balanced [lo, hi) share of ``n`` independent units (samples / chains / draws) for ``rank``."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_pairs(n_samples, n_images, rank=None, world=None, quantum=64):
    """Balanced share of the ``n_samples x n_images`` (sample, image) grid for ``rank`` as a list of
    ``(sample, img_lo, img_hi)``: whole samples when ``n_samples`` divides evenly over the ranks, otherwise a contiguous
    run of ``quantum``-image blocks in sample-major order, so that S < G or S mod G != 0 no longer leaves ranks idle
    (the partial [N, C] sums of all ranks still meet in the same single all-reduce).  Shares are equal to within
    one block, plus up to two blocks at either end where a boundary was moved onto a nearby sample boundary."""
    if rank is None or world is None:
        rank, world = rank_world()
    if n_samples <= 0 or n_images <= 0:
        return []
    if n_samples % world == 0:
        lo, hi = shard_range(n_samples, rank, world)
        return [(s, 0, n_images) for s in range(lo, hi)]
    nb = (n_images + quantum - 1) // quantum                 # image blocks per sample
    lo, hi = shard_range(n_samples * nb, rank, world)
    # a share boundary that falls within two blocks of a sample boundary is moved onto it: a 64- or 128-image stub of one
    # sample costs its rank a whole launch chain (and two re-zeroings of the engine's plane images) for 0.1 % of its work
    snap = min(2, nb // 32)

    def snapped(u):
        b = u % nb
        if 0 < u < n_samples * nb and snap:
            if b <= snap:
                return u - b
            if nb - b <= snap:
                return u + (nb - b)
        return u
    lo, hi = snapped(lo), snapped(hi)
    out = []
    u = lo
    while u < hi:
        s, b0 = divmod(u, nb)
        b1 = min(nb, b0 + (hi - u))
        out.append((s, b0 * quantum, min(n_images, b1 * quantum)))
        u += b1 - b0
    return out


def chain_elem_offset(chain_id, D):
    """Philox element base of chain ``chain_id`` so that chains on different ranks never share a noise stream
    (K1's counter is the global element index; multiple of 4 as the kernel requires)."""
    ld = (D + 3) // 4 * 4
    return chain_id * ld


def allreduce_bma(proba_sum, entropy_sum, num_samples):
    """Sum the BMA accumulators over ranks with ONE collective: [N*C + N + 1] fp32 packed into a single buffer.
    Returns (proba_sum, entropy_sum, num_samples) reduced; inputs are left untouched."""
    if not is_distributed():
        return proba_sum, entropy_sum, num_samples
    n, c = proba_sum.shape
    packed = torch.empty(n * c + n + 1, dtype=torch.float32, device=proba_sum.device)
    packed[:n * c].copy_(proba_sum.reshape(-1))
    packed[n * c:n * c + n].copy_(entropy_sum)
    packed[-1] = float(num_samples)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed[:n * c].view(n, c), packed[n * c:n * c + n], int(round(packed[-1].item()))


def allreduce_max_scalar(value, device):
    """Max over ranks of a python float (used for max-over-ranks timing)."""
    if not is_distributed():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
