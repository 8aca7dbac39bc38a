"""Multi-GPU plumbing for the decode path: images are independent, so ranks shard the batch by
image and the only cross-rank step is a final reduce of the timing.
There is no data-path collective."""


def shard_indices(n_items, rank, world):
    """Images {i : i mod world == rank} -- frame f of a sequence goes to GPU f mod world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    return list(range(rank, n_items, world))


def reduce_job(elapsed_ms, units, device=None):
    """Whole-job figures from per-rank ones: (max over ranks of elapsed_ms, sum over ranks of
    each entry of `units`).  Uses the default process group (NCCL on GPUs, gloo in the CPU
    tests); with no process group it returns its inputs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), [float(u) for u in units]
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(x) for x in units], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(x) for x in u.tolist()]
