"""Host logic of the episode-parallel replay harness (genima_b200/eval_replay.py) with a fake agent: episode structure
follows controller/eval_genima.py (per-episode re-seed, advance by len(actions), stop past episode_length), and the
union of all ranks' shards covers every (task, episode) exactly once."""
import numpy as np

from genima_b200 import distributed as gd
from genima_b200.eval_replay import RLBENCH_25, StubEnv, run_units, summarize


def test_episode_loop_structure():
    seeds, calls = [], []

    def fake_agent(views, qpos, k):
        assert views.shape == (4, 256, 256, 3) and views.dtype == np.uint8 and qpos.shape == (1, 8)
        calls.append(k)
        return np.zeros((20, 8), dtype=np.float32)

    recs = run_units([("open_box", 0), ("open_box", 1)], fake_agent, episode_length=200, reseed=seeds.append)
    assert seeds == [2, 2]                                # diffusion_seed re-applied per episode (eval_genima.py:129-135)
    # 20 sim steps per agent step, terminate once t > 200 -> 11 agent steps (eval_genima.py:261-275)
    assert [r["agent_steps"] for r in recs] == [11, 11] and recs[0]["sim_steps"] == 220
    assert calls == list(range(11)) * 2


def test_stub_env_is_deterministic_per_unit():
    a, b = StubEnv("t", 3, size=64).observe(), StubEnv("t", 3, size=64).observe()
    c = StubEnv("t", 4, size=64).observe()
    assert np.array_equal(a[0], b[0]) and not np.array_equal(a[0], c[0])


def test_shards_cover_config5_exactly_once():
    world = 8
    seen = []
    for r in range(world):
        seen += gd.shard_units(RLBENCH_25, 25, r, world)
    assert len(seen) == 625 and len(set(seen)) == 625      # 25 tasks x 25 episodes (BASELINE configs[4])
    sizes = [len(gd.shard_units(RLBENCH_25, 25, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    s = summarize([{"agent_steps": 11, "mean_step_time": 0.03}] * 4, wall_s=2.0, world=2)
    assert s["agent_steps"] == 44 and s["agent_steps_per_sec"] == 22.0


def test_batched_episode_loop_matches_the_serial_one():
    """Episodes advanced in lock-step (episodes in flight) see the same observations, step counts and per-slot re-seeding
    as the serial loop; a ragged last group is padded, never stepped."""
    from genima_b200.eval_replay import run_units_batched

    units = [("open_box", 0), ("open_box", 1), ("push_button", 0)]
    seen_serial, seen_batched, seeds = {}, {}, []

    def fake_serial(views, qpos, k):
        seen_serial.setdefault(len(seen_serial) // 11, []).append(int(views.sum()))
        return np.full((20, 8), float(views[0, 0, 0, 0]), dtype=np.float32)

    def fake_batched(views, qpos, k, active):
        assert views.shape == (2, 4, 64, 64, 3) and qpos.shape == (2, 1, 8) and len(active) == 2
        for i, a in enumerate(active):
            if a:
                seen_batched.setdefault((len(seeds), i), []).append(int(views[i].sum()))
        return np.stack([np.full((20, 8), float(v[0, 0, 0, 0]), dtype=np.float32) for v in views])

    serial = run_units(units, fake_serial, size=64, episode_length=200)
    batched = run_units_batched(units, fake_batched, 2, size=64, episode_length=200,
                                reseed=lambda slot, s: seeds.append((slot, s)))
    assert seeds == [(0, 2), (1, 2), (0, 2)]
    assert [(r["task"], r["episode"], r["agent_steps"], r["sim_steps"], r["checksum"]) for r in serial] == \
           [(r["task"], r["episode"], r["agent_steps"], r["sim_steps"], r["checksum"]) for r in batched]


def test_async_sim_workers_give_the_serial_records():
    """Simulator worker threads + batching server (run_units_async): every episode sees its own observations in order,
    finishes with the serial loop's step counts and checksums, and is re-seeded when it starts in a slot -- also when one
    simulator is much slower than the others (its episode lags, the others are not held back)."""
    from genima_b200.eval_replay import run_units_async

    units = [("open_box", 0), ("open_box", 1), ("push_button", 0), ("push_button", 1), ("close_jar", 0)]
    seeds, batch_sizes = [], []

    def fake_serial(views, qpos, k):
        return np.full((20, 8), float(views[0, 0, 0, 0]), dtype=np.float32)

    def fake_batched(views, qpos, k, active):
        assert views.shape == (3, 4, 64, 64, 3) and qpos.shape == (3, 1, 8) and len(active) == 3 and any(active)
        batch_sizes.append(sum(active))
        return np.stack([np.full((20, 8), float(v[0, 0, 0, 0]), dtype=np.float32) for v in views])

    serial = run_units(units, fake_serial, size=64, episode_length=200)
    key = lambda r: (r["task"], r["episode"], r["agent_steps"], r["sim_steps"], r["checksum"])   # noqa: E731
    for delay in (None, lambda slot: 0.004 if slot == 0 else 0.0):
        seeds.clear()
        batch_sizes.clear()
        recs = run_units_async(units, fake_batched, 3, size=64, episode_length=200,
                               reseed=lambda slot, s: seeds.append((slot, s)), sim_delay=delay)
        assert [key(r) for r in recs] == [key(r) for r in serial]
        assert len(seeds) == len(units) and all(s == 2 for _, s in seeds)
        assert sum(batch_sizes) == sum(r["agent_steps"] for r in serial)
    # with one slow simulator the server did not wait for it: some calls ran with fewer than all slots active
    assert min(batch_sizes) < 3
