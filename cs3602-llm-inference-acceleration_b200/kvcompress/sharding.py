"""Multi-GPU sharding of the compress path.

Every (layer, batch, head) row is compressed independently (the reference loops over layers and
reduces over dim=-1 only, e.g. h2o_l2.py:77,122-141), so the path shards with **no data-path
collective**: each rank owns a contiguous block of decode streams (batch) — or of layers — and runs
an ordinary single-GPU call.  Plans are pure host arithmetic and are identical on every rank.
`torch.distributed` is used only after the timed region, to combine timings / checksums.
"""

from __future__ import annotations

from typing import Dict, List, Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) block of `total` items owned by `rank` (sizes differ by at most 1)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad world/rank {world}/{rank}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(kv, world: int, rank: int):
    """This rank's decode streams of a [B, H, S, D] cache: views, no copy."""
    out = []
    for keys, values in kv:
        lo, hi = shard_range(keys.size(0), world, rank)
        out.append((keys[lo:hi], values[lo:hi]))
    return out


def shard_layers(n_layers: int, world: int, rank: int) -> List[int]:
    """Layer indices owned by `rank` under layer sharding (pipeline-style placement)."""
    lo, hi = shard_range(n_layers, world, rank)
    return list(range(lo, hi))


def job_units(how: str, n_layers: int, n_blocks: int, world: int, rank: int, layer_group: int = 4):
    """This rank's share of a GLOBAL compress job as a list of slabs ``(layer_ids, block_ids)``, a slab being what is
    resident at one time.  The job is ``n_layers`` layers x ``n_blocks`` blocks of decode streams; every
    (layer, block) pair is owned by exactly one rank, whichever way the job is cut:

    ``how="batch"``  rank owns a contiguous range of stream blocks, all layers; one slab per block
                     (reference: units are independent per (batch, head), h2o_l2.py:77,122-141);
    ``how="layer"``  rank owns a contiguous range of layers (``shard_layers``), every block; one slab per
                     ``layer_group`` layers.  Per-layer budgets (pyramid_kv.py:84-97) are a function of the GLOBAL
                     layer index, so plans are built for all layers on every rank and indexed by ``layer_ids``."""
    if how == "batch":
        lo, hi = shard_range(n_blocks, world, rank)
        return [(list(range(n_layers)), [blk]) for blk in range(lo, hi)]
    if how == "layer":
        mine = shard_layers(n_layers, world, rank)
        return [(mine[i:i + layer_group], list(range(n_blocks))) for i in range(0, len(mine), layer_group)]
    raise ValueError(f"unknown sharding {how!r}: 'batch' or 'layer'")


def combine_stats(local: Dict[str, float]) -> Dict[str, float]:
    """After the timed region: MAX of times, SUM of bytes / checksums across ranks.

    Keys ending in ``_ms`` / ``_s`` are reduced with MAX (the job is as slow as its slowest rank),
    everything else with SUM.  Works on any initialised backend (nccl on GPUs, gloo in CPU tests);
    without an initialised process group it returns `local` unchanged."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    keys = sorted(local)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    out = {}
    for op, pick in ((dist.ReduceOp.MAX, lambda k: k.endswith(("_ms", "_s"))),
                     (dist.ReduceOp.SUM, lambda k: not k.endswith(("_ms", "_s")))):
        names = [k for k in keys if pick(k)]
        if not names:
            continue
        t = torch.tensor([float(local[k]) for k in names], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        out.update({k: float(v) for k, v in zip(names, t.tolist())})
    return out
