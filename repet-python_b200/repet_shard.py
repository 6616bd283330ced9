"""Clip sharding across ranks (one process per GPU).  REPET clips never interact
(repet.py:125, 263, 483, 631, 771 are pure functions of one clip), so the only multi-GPU
logic is which rank owns which clips and a max-over-ranks reduction of the step time."""


def shard_range(number_items, rank, world_size):
    """Contiguous, balanced shard [lo, hi) of `number_items` for `rank` (the first
    number_items % world_size ranks get one extra item)."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size")
    base, extra = divmod(number_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks of the default process group
    (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    tensor = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
    return [float(v) for v in tensor]


def gather_int_arrays(array, device=None):
    """all_gather of equally shaped int32 arrays (e.g. per-rank periods); returns the
    concatenation in rank order on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist

    tensor = torch.as_tensor(np.ascontiguousarray(array, dtype=np.int32), device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor.cpu().numpy()
    parts = [torch.empty_like(tensor) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, tensor)
    return torch.cat(parts).cpu().numpy()
