"""Multi-GPU partitioning of the hot path: one process per GPU,
torch.distributed for the plumbing.

* ensembles (C3) shard by conformation, no data-path collective; an optional
  all-gather collects the small per-structure results;
* dense all-pairs systems (C4) and covariance products (C5) shard by row slabs
  of the output; each rank calls the C ABI with its [row0, row1) range.
"""

__all__ = ["shard_range", "row_slab", "gather_results", "world", "dcc_row_partitioned"]


def world():
    """(rank, world_size) from torch.distributed if initialised, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(total, rank, world_size):
    """Contiguous, balanced split of `total` units: the first `total % world`
    ranks get one extra unit.  Returns (start, stop)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def row_slab(n_rows, rank, world_size, align=1):
    """Row slab [row0, row1) of an n_rows output for this rank; slab boundaries
    are multiples of `align` (e.g. 2 so that residue pairs stay together)."""
    units = (n_rows + align - 1) // align
    a, b = shard_range(units, rank, world_size)
    return min(a * align, n_rows), min(b * align, n_rows)


def gather_results(local, total, dim=0):
    """All-gather per-rank result tensors that were produced by `shard_range`
    (ragged along `dim`) into the full tensor on every rank."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(total, r, ws) for r in range(ws)]
    longest = max(b - a for a, b in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = longest
    padded = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    parts = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(parts, padded)
    return torch.cat([p.narrow(dim, 0, b - a) for p, (a, b) in zip(parts, sizes)], dim=dim)


def dcc_row_partitioned(D, lam, modes, norm=True, scale=1.0, gather=False):
    """every rank computes the rows [row0,row1) of the DCC /
    covariance contraction sum_k U_k U_k^T / lam_k with the DMMA kernel
    (`scb_dcc` takes the row range); the (lam, modes) operands are replicated.
    No exchange is needed inside the contraction (the normalisation uses the
    locally recomputed diagonal).  Returns (row0, row1, slab) or, with
    gather=True, the full matrix on every rank."""
    from . import _engine
    rank, ws = world()
    n = int(modes.shape[1]) // D
    row0, row1 = row_slab(n, rank, ws)
    slab = _engine.modes_dcc(D, lam, modes, norm=norm, scale=scale, rows=(row0, row1))
    if not gather:
        return row0, row1, slab
    return gather_results(slab, n, dim=0)
