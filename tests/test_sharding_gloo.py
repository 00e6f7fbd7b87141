"""N>1 host logic on CPU: world_size-2 `gloo` run of the partitioning helpers
used by the multi-GPU paths (ensemble shards, row slabs, result gather)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from springcraft_b200.parallel import gather_results, row_slab, shard_range
from .conftest import ROOT


@pytest.mark.parametrize("total,ws", [(4096, 8), (4096, 3), (5, 8), (0, 2), (20000, 7)])
def test_shard_range_partitions(total, ws):
    spans = [shard_range(total, r, ws) for r in range(ws)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_row_slab_alignment():
    spans = [row_slab(301, r, 4, align=2) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == 301
    assert all(a % 2 == 0 for a, _ in spans)
    assert all(b == c for (_, b), (c, _) in zip(spans, spans[1:]))


def _worker(rank, ws, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        a, b = shard_range(total, rank, ws)
        # stand-in for the per-structure results each rank produces on its shard
        local = torch.arange(a, b, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0, 100.0]])
        full = gather_results(local, total)
        want = torch.arange(total, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0, 100.0]])
        ok = bool(torch.equal(full, want))
        # max-over-ranks timing reduction as used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t.item()) == ws
        torch.save({"ok": ok, "span": (a, b)}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [4096, 11])
def test_two_rank_gather(tmp_path, total):
    ws = 2
    port = 29500 + (os.getpid() % 500) + total % 7
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(ws, port, total, out), nprocs=ws, join=True)
    spans = []
    for r in range(ws):
        res = torch.load(f"{out}.{r}")
        assert res["ok"]
        spans.append(res["span"])
    assert spans[0][1] == spans[1][0] and spans[1][1] == total
