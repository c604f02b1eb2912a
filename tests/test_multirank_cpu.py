"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding, replicated plans,
post-run stat combination.  The data path itself has no collective."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kvcompress import _planner as P
from kvcompress import sharding


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_layers(32, 8, 3) == [12, 13, 14, 15]
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_job_units_cover_every_layer_block_pair_once():
    """bench.py's strong-scaling section: a global job of L layers x n_blocks stream blocks, cut by batch or by layer."""
    L, n_blocks = 32, 8
    for how in ("batch", "layer"):
        for world in (1, 2, 4, 8):
            seen = []
            for rank in range(world):
                units = sharding.job_units(how, L, n_blocks, world, rank)
                assert len(units) == 8 // world          # slabs per rank: 8 / 4 / 2 / 1
                for layer_ids, block_ids in units:
                    assert (len(layer_ids), len(block_ids)) == ((L, 1) if how == "batch" else (4, n_blocks))
                    seen += [(l, b) for l in layer_ids for b in block_ids]
            assert sorted(seen) == [(l, b) for l in range(L) for b in range(n_blocks)]
    with pytest.raises(ValueError):
        sharding.job_units("heads", L, n_blocks, 2, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kvcompress

        g = torch.Generator().manual_seed(0)  # every rank builds the same global cache, then takes its shard
        kv = [(torch.randn(6, 2, 700, 16, generator=g), torch.randn(6, 2, 700, 16, generator=g)) for _ in range(3)]
        mine = sharding.shard_batch(kv, world, rank)
        assert mine[0][0].size(0) == 3 and mine[0][0].data_ptr() == kv[0][0][3 * rank:].data_ptr()
        # plans are host arithmetic on sequence lengths: identical on every rank
        plans = P.plan_h2o([k.size(2) for k, _ in mine], 4, 64, 444, [])
        gathered = [None] * world
        dist.all_gather_object(gathered, [(p.kind, p.sink, p.sel_lo, p.sel_hi, p.k_sel, p.tail) for p in plans])
        assert gathered[0] == gathered[1]
        # a view-only method runs on CPU shards; its result is the shard of the global result
        out = kvcompress.recent_only_compress(mine, window_size=512, skip_layers=[0])
        full = kvcompress.recent_only_compress(kv, window_size=512, skip_layers=[0])
        assert all(torch.equal(o[0], f[0][3 * rank:3 * rank + 3]) for o, f in zip(out, full))
        # a global job cut by batch and by layer: per-(layer, block) terms add up to the same total on any cut
        term = lambda l, b: (l + 1) * 1000 + 7 * b
        for how in ("batch", "layer"):
            mine_sum = sum(term(l, b) for ls, bs in sharding.job_units(how, 8, 4, world, rank, layer_group=2) for l in ls for b in bs)
            total = sharding.combine_stats({"checksum": float(mine_sum)})["checksum"]
            assert total == float(sum(term(l, b) for l in range(8) for b in range(4)))
        local_bytes = P.algorithmic_bytes(plans, mine[0][0].size(0), 2, 16, 4)
        stats = sharding.combine_stats({"step_ms": 10.0 + rank, "bytes": float(local_bytes), "streams": 3.0})
        q.put((rank, stats, P.algorithmic_bytes(plans, 6, 2, 16, 4)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _rank, stats, global_bytes in results:
        assert stats["step_ms"] == 11.0            # MAX over ranks
        assert stats["streams"] == 6.0             # SUM over ranks
        assert stats["bytes"] == float(global_bytes)  # shards add up to the global job


def test_combine_stats_without_process_group():
    assert sharding.combine_stats({"a_ms": 1.0, "b": 2.0}) == {"a_ms": 1.0, "b": 2.0}
