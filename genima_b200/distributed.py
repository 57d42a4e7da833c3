"""Multi-GPU plumbing for the episode-parallel evaluation: one process per GPU, ONE broadcast of the
packed weight arena at start-up, then no per-step collective; (task, episode) units are sharded round-robin and the
per-episode records are gathered at the end.  Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests).

The reference evaluates serially in one process (controller/eval_genima.py:115-330; README.md:299 lists it as a
limitation); episodes are independent — each re-seeds its generator and resets its env (eval_genima.py:129-142)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def _is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def arena_layout(shapes_by_model: "OrderedDict[str, OrderedDict[str, Tuple[int, ...]]]"):
    """-> (index {(model, key): (offset, shape)}, total elements); every tensor starts on a 16-byte boundary."""
    index, off = OrderedDict(), 0
    for model, shapes in shapes_by_model.items():
        for key, shp in shapes.items():
            n = 1
            for v in shp:
                n *= v
            index[(model, key)] = (off, tuple(shp))
            off += (n + 7) // 8 * 8
    return index, off


def broadcast_weights(shapes_by_model, state_dicts=None, src: int = 0, device="cpu", timings: dict = None):
    """Rank `src` packs its fp16 state dicts into one flat arena; one broadcast; every rank returns state dicts whose
    tensors are views into its copy of the arena (so the device graphs bind them without another copy).
    `timings` (optional dict) receives pack_s (host state dicts -> device arena on the source rank) and broadcast_s (the
    collective alone, synchronised on both sides; the communicator must already be warm for this to be a bandwidth
    figure)."""
    import time

    index, total = arena_layout(shapes_by_model)
    arena = torch.empty(total, dtype=torch.float16, device=device)
    rank = dist.get_rank() if _is_dist() else 0
    is_cuda = torch.device(device).type == "cuda"
    t0 = time.perf_counter()
    if rank == src:
        if state_dicts is None:
            raise ValueError("the source rank must supply the state dicts")
        for (model, key), (off, shp) in index.items():
            t = state_dicts[model][key]
            arena[off:off + t.numel()].copy_(t.reshape(-1).to(torch.float16))
    if is_cuda:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    if _is_dist() and dist.get_world_size() > 1:
        dist.barrier()
        if is_cuda:
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        dist.broadcast(arena, src=src)
        if is_cuda:
            torch.cuda.synchronize()
    t2 = time.perf_counter()
    if timings is not None:
        timings["pack_s"] = t1 - t0 if rank == src else 0.0
        timings["broadcast_s"] = (t2 - t1) if (_is_dist() and dist.get_world_size() > 1) else 0.0
    out = OrderedDict((m, OrderedDict()) for m in shapes_by_model)
    for (model, key), (off, shp) in index.items():
        n = 1
        for v in shp:
            n *= v
        out[model][key] = arena[off:off + n].view(shp)
    return out, arena


def sync_tune_caches(ops_list, src: int = 0) -> int:
    """Rank `src` has measured its GEMM tile configurations (gn_set_autotune); every other rank adopts them
    (gn_tune_cache_export / gn_tune_cache_import) instead of timing its own, so that all ranks launch identical tile and
    split-K configurations and every episode gives bit-identical results wherever it is sharded.  `ops_list`: the Ops
    handles in the same order on every rank.  Returns the number of bytes exchanged."""
    if not _is_dist() or dist.get_world_size() == 1:
        return 0
    payload = [[o.tune_cache_export() for o in ops_list] if dist.get_rank() == src else None]
    dist.broadcast_object_list(payload, src=src)
    if dist.get_rank() != src:
        for o, blob in zip(ops_list, payload[0]):
            o.tune_cache_import(blob, replace=True)
    return sum(len(b) for b in payload[0])


def shard_units(tasks: Sequence[str], episodes_per_task: int, rank: int, world: int) -> List[Tuple[str, int]]:
    """Static round-robin split of the (task, episode index) units; the episode index is what the reference passes to
    reset_to_demo (controller/eval_genima.py:142), so a unit is self-describing."""
    units = [(t, e) for t in tasks for e in range(episodes_per_task)]
    return units[rank::world]


def gather_records(records: List[dict]) -> List[dict]:
    """All per-episode records on every rank (a few KB: no tensor collective needed)."""
    if not _is_dist() or dist.get_world_size() == 1:
        return list(records)
    bucket: List[List[dict]] = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, records)
    return [r for part in bucket for r in part]


def reduce_max(value: float, device="cpu") -> float:
    if not _is_dist() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device="cpu") -> float:
    if not _is_dist() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
