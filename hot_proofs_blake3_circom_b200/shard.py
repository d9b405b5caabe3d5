"""shard.py -- multi-GPU plan: the batch is a range of independent instance indices, so rank r of R simply owns a
contiguous slice; no collective touches the data path.  The only cross-rank step is a reduction of the
per-shard counters / checksums, which is what reduce_summary() does over torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous, balanced partition of [0, n_total): the first n_total % world ranks get one extra instance."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(int(n_total), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def local_summary(status, sums=None):
    """Per-shard compact result: [instances, ok, failed, xor of witness checksums] as int64 (checksums are u64 bit patterns)."""
    status = np.asarray(status)
    x = np.uint64(0)
    if sums is not None and len(sums):
        x = np.bitwise_xor.reduce(np.asarray(sums, np.uint64))
    return np.array([status.size, int((status == 0).sum()), int((status != 0).sum()), np.int64(np.uint64(x).view(np.int64))],
                    np.int64)


def reduce_summary(local, dist=None, device=None):
    """Sum the counters and XOR the checksums over all ranks (no-op without an initialised process group)."""
    import torch
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local, np.int64).copy()
    t = torch.from_numpy(np.asarray(local, np.int64).copy())
    if device is not None:
        t = t.to(device)
    counts = t[:3].clone()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    gathered = [torch.zeros_like(t[3:4]) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, t[3:4].contiguous())
    x = np.uint64(0)
    for g in gathered:
        x ^= np.uint64(np.int64(g.cpu().item()).view(np.uint64))
    return np.array(list(counts.cpu().numpy()) + [np.int64(np.uint64(x).view(np.int64))], np.int64)
