"""N > 1 plumbing on CPU: world_size-2 gloo processes exercise the weight-arena broadcast, the (task, episode) sharding
and the record gather used by bench.py / the episode-parallel evaluation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genima_b200 import distributed as gd
from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, UNetConfig


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = {"controlnet": W.controlnet_shapes(UNetConfig.tiny()), "act": W.act_shapes(ACTConfig.tiny())}
        sds = None
        if rank == 0:
            sds = {"controlnet": W.synth_state_dict(shapes["controlnet"], 1), "act": W.synth_state_dict(shapes["act"], 3)}
        got, arena = gd.broadcast_weights(shapes, sds, src=0)
        want = W.synth_state_dict(shapes["controlnet"], 1)
        ok = all(torch.equal(got["controlnet"][k], want[k]) for k in want)
        ok = ok and all(t.data_ptr() % 16 == 0 for t in got["act"].values())        # TMA-aligned views of ONE arena
        units = gd.shard_units(["open_box", "close_box", "push_button"], 5, rank, world)
        recs = gd.gather_records([{"task": t, "episode": e, "rank": rank} for t, e in units])
        # tile configurations measured on rank 0 are adopted by every other rank (bit-identical sharded results)
        class FakeOps:
            def __init__(self, blob):
                self.blob = blob

            def tune_cache_export(self):
                return self.blob

            def tune_cache_import(self, blob, replace=False):
                assert replace
                self.blob = blob

        fake = [FakeOps(b"1:32:320:45=160,1,5,256\n" if rank == 0 else b"other\n"), FakeOps(b"" if rank == 0 else b"x")]
        nbytes = gd.sync_tune_caches(fake, src=0)
        ok = ok and nbytes == 24 and fake[0].blob == b"1:32:320:45=160,1,5,256\n" and fake[1].blob == b""
        mx = gd.reduce_max(10.0 + rank)
        sm = gd.reduce_sum(float(len(units)))
        q.put((rank, ok, len(units), sorted((r["task"], r["episode"]) for r in recs), mx, sm))
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_units = sorted((t, e) for t in ["open_box", "close_box", "push_button"] for e in range(5))
    for rank, ok, n, recs, mx, sm in res:
        assert ok, f"rank {rank}: broadcast weights differ from the source"
        assert recs == all_units                                  # every unit exactly once across ranks
        assert mx == 11.0 and sm == 15.0
    assert sorted(r[2] for r in res) == [7, 8]


def test_single_process_paths_need_no_process_group():
    shapes = {"act": W.act_shapes(ACTConfig.tiny())}
    sds = {"act": W.synth_state_dict(shapes["act"], 3)}
    got, _ = gd.broadcast_weights(shapes, sds)
    assert all(torch.equal(got["act"][k], sds["act"][k]) for k in sds["act"])
    assert gd.shard_units(["a"], 3, 0, 1) == [("a", 0), ("a", 1), ("a", 2)]
    assert gd.gather_records([{"x": 1}]) == [{"x": 1}] and gd.reduce_max(3.0) == 3.0
