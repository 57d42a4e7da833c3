"""Episode-parallel replay harness: the reference's evaluation loop (controller/eval_genima.py:105-346) with the
simulator replaced by a stub env, sharded over GPUs.

The real loop is `for episode: reset generator(seed 2) -> reset env -> while not done: tile -> infer -> untile -> act ->
env.step(actions)`; RLBench / CoppeliaSim cannot run here, so `StubEnv` returns seeded uint8 camera frames and advances by
`len(actions)` sim steps per agent step (eval_genima.py:261-263) until `episode_length` (eval_genima.py:272-275).
Everything else — per-episode generator re-seeding, per-episode JSON records with mean gen/control time, the round-robin
(task, episode) sharding and the final gather — mirrors the reference so that the throughput it reports is the
throughput an actual evaluation would see on the GPU side.

    torchrun --nproc-per-node N -m genima_b200.eval_replay --tasks 25 --episodes 25 --preset sd-turbo
"""
from __future__ import annotations

import argparse
import json
import os
import time
import zlib
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import distributed as gd

RLBENCH_25 = [f"task_{i:02d}" for i in range(25)]   # the 25 task names only label the records here


class StubEnv:
    """Seeded stand-in for GenimaRLBenchEnv: 4 cameras x uint8 [3, S, S], 8-dim proprioception, fixed episode length."""

    def __init__(self, task: str, episode: int, size: int = 256, episode_length: int = 200, state_dim: int = 8):
        self.rng = np.random.RandomState((zlib.crc32(task.encode()) & 0xFFFF) * 1000 + episode)
        self.size, self.episode_length, self.state_dim = size, episode_length, state_dim
        self.t = 0

    def observe(self) -> Tuple[np.ndarray, np.ndarray]:
        views = self.rng.randint(0, 256, size=(4, self.size, self.size, 3), dtype=np.uint8)
        qpos = self.rng.randn(1, self.state_dim).astype(np.float32)
        return views, qpos

    def step(self, actions: np.ndarray) -> bool:
        self.t += len(actions)                       # eval_genima.py:261-263: advance by the whole action chunk
        return self.t > self.episode_length          # eval_genima.py:272-275


def run_units(units: Sequence[Tuple[str, int]], agent_step: Callable[[np.ndarray, np.ndarray, int], np.ndarray],
              size: int = 256, episode_length: int = 200, diffusion_seed: int = 2,
              reseed: Callable[[int], None] = lambda seed: None) -> List[dict]:
    """Runs the (task, episode) units assigned to this rank.  `agent_step(views_u8, qpos, step) -> actions [20, 8]`."""
    records = []
    for task, ep in units:
        reseed(diffusion_seed)                       # eval_genima.py:129-135: generator re-seeded per episode
        env = StubEnv(task, ep, size=size, episode_length=episode_length)
        steps, t_total, done = 0, 0.0, False
        while not done:
            views, qpos = env.observe()
            t0 = time.perf_counter()
            actions = agent_step(views, qpos, steps)
            t_total += time.perf_counter() - t0
            done = env.step(actions)
            steps += 1
        records.append({"task": task, "episode": ep, "agent_steps": steps, "sim_steps": env.t,
                        "mean_step_time": t_total / max(steps, 1), "checksum": float(np.abs(actions).sum())})
    return records


def run_units_batched(units: Sequence[Tuple[str, int]],
                      agent_step: Callable[[np.ndarray, np.ndarray, int, Sequence[bool]], np.ndarray], in_flight: int,
                      size: int = 256, episode_length: int = 200, diffusion_seed: int = 2,
                      reseed: Callable[[int, int], None] = lambda slot, seed: None) -> List[dict]:
    """Like run_units, with up to `in_flight` independent episodes advanced in lock-step and batched into ONE agent-step
    call (the reference loop is serial, README.md:299; episodes are independent, eval_genima.py:129-142, so batching them
    changes no episode's results).  `agent_step(views [E, 4, S, S, 3], qpos [E, 1, 8], step, active [E]) -> actions
    [E, 20, 8]`; `reseed(slot, seed)` re-seeds the generator of batch slot `slot` when a new episode starts there."""
    records = []
    for g0 in range(0, len(units), in_flight):
        group = list(units[g0:g0 + in_flight])
        envs = [StubEnv(t, e, size=size, episode_length=episode_length) for t, e in group]
        for slot in range(len(group)):
            reseed(slot, diffusion_seed)
        steps = [0] * len(group)
        done = [False] * len(group)
        last = [None] * len(group)
        t_total, k = 0.0, 0
        obs = [env.observe() for env in envs]
        while not all(done):
            for i, env in enumerate(envs):
                if not done[i] and k > 0:
                    obs[i] = env.observe()
            pad = in_flight - len(group)
            views = np.stack([o[0] for o in obs] + [obs[0][0]] * pad)
            qpos = np.stack([o[1] for o in obs] + [obs[0][1]] * pad)
            t0 = time.perf_counter()
            actions = agent_step(views, qpos, k, [not d for d in done] + [False] * pad)
            t_total += time.perf_counter() - t0
            for i, env in enumerate(envs):
                if not done[i]:
                    last[i] = actions[i]
                    done[i] = env.step(actions[i])
                    steps[i] += 1
            k += 1
        for i, (task, ep) in enumerate(group):
            records.append({"task": task, "episode": ep, "agent_steps": steps[i], "sim_steps": envs[i].t,
                            "mean_step_time": t_total / max(k, 1), "checksum": float(np.abs(last[i]).sum())})
    return records


def run_units_async(units: Sequence[Tuple[str, int]],
                    agent_step: Callable[[np.ndarray, np.ndarray, int, Sequence[bool]], np.ndarray], in_flight: int,
                    size: int = 256, episode_length: int = 200, diffusion_seed: int = 2,
                    reseed: Callable[[int, int], None] = lambda slot, seed: None, sim_delay: Callable[[int], float] = None,
                    gather_window_s: float = None) -> List[dict]:
    """Asynchronous variant of run_units_batched: `in_flight` simulator workers (threads; one episode each at a time, the
    next unit of the shard when it finishes) post their observations to a request queue, and ONE server loop batches
    whatever is ready (waiting at most `gather_window_s` for stragglers once a request is there; default: 30 % of the
    running mean of the agent-step time, so a full batch is preferred as long as waiting for it is cheap) into a single
    agent-step call.  A slow simulator step therefore delays only its own episode, never the batch (the lock-step loop waits for the
    slowest env every step).  Worker i always uses batch slot i and the generator of that slot, re-seeded when a new
    episode starts there, so an episode's noise stream does not depend on what its neighbours do.  `sim_delay(slot)`:
    optional extra seconds per env.step (tests / what-if runs)."""
    import queue
    import threading

    todo = list(units)
    todo_lock = threading.Lock()
    requests: "queue.Queue" = queue.Queue()
    replies = [queue.Queue(maxsize=1) for _ in range(in_flight)]
    records: List[dict] = []
    rec_lock = threading.Lock()
    live = [in_flight]

    def worker(slot: int):
        while True:
            with todo_lock:
                unit = todo.pop(0) if todo else None
            if unit is None:
                break
            task, ep = unit
            env = StubEnv(task, ep, size=size, episode_length=episode_length)
            steps, done, actions, t_wait = 0, False, None, 0.0
            first = True
            while not done:
                views, qpos = env.observe()
                t0 = time.perf_counter()
                requests.put((slot, views, qpos, first))
                actions = replies[slot].get()
                t_wait += time.perf_counter() - t0
                first = False
                if sim_delay is not None:
                    time.sleep(sim_delay(slot))
                done = env.step(actions)
                steps += 1
            with rec_lock:
                records.append({"task": task, "episode": ep, "agent_steps": steps, "sim_steps": env.t,
                                "mean_step_time": t_wait / max(steps, 1), "checksum": float(np.abs(actions).sum())})
        requests.put((slot, None, None, False))      # this worker is done

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(in_flight)]
    for t in threads:
        t.start()
    last_views, last_qpos = None, None
    k = 0
    step_ema = 0.015
    while live[0] > 0:
        batch = [requests.get()]
        deadline = time.perf_counter() + (gather_window_s if gather_window_s is not None else 0.3 * step_ema)
        while len(batch) < live[0]:
            try:
                batch.append(requests.get(timeout=max(0.0, deadline - time.perf_counter())))
            except queue.Empty:
                break
        work = []
        for slot, views, qpos, first in batch:
            if views is None:
                live[0] -= 1
            else:
                if first:
                    reseed(slot, diffusion_seed)
                work.append((slot, views, qpos))
        if not work:
            continue
        if last_views is None:
            last_views, last_qpos = work[0][1], work[0][2]
        v = [last_views] * in_flight
        q = [last_qpos] * in_flight
        active = [False] * in_flight
        for slot, views, qpos in work:
            v[slot], q[slot], active[slot] = views, qpos, True
        t0 = time.perf_counter()
        actions = agent_step(np.stack(v), np.stack(q), k, active)
        step_ema = 0.8 * step_ema + 0.2 * (time.perf_counter() - t0)
        k += 1
        for slot, _, _ in work:
            replies[slot].put(np.array(actions[slot]))
    for t in threads:
        t.join()
    order = {u: i for i, u in enumerate(units)}
    records.sort(key=lambda r: order[(r["task"], r["episode"])])
    return records


def summarize(records: List[dict], wall_s: float, world: int) -> dict:
    steps = sum(r["agent_steps"] for r in records)
    return {"episodes": len(records), "agent_steps": steps, "wall_s": wall_s, "n_gpus": world,
            "agent_steps_per_sec": steps / wall_s if wall_s > 0 else 0.0,
            "mean_step_time": float(np.mean([r["mean_step_time"] for r in records])) if records else 0.0}


def main(argv=None):
    import torch.distributed as dist

    from .act_policy import DeviceACT
    from .ops import Ops
    from .pipeline import B200ControlNetPipeline
    from .step import GenimaStep
    from . import weights as W
    from .configs import ACTConfig, UNetConfig, VAEConfig

    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=25)
    ap.add_argument("--episodes", type=int, default=25)
    ap.add_argument("--episode-length", type=int, default=200)
    ap.add_argument("--denoise-steps", type=int, default=5)
    ap.add_argument("--preset", default="sd-turbo", choices=["sd-turbo", "tiny"])
    ap.add_argument("--out", default="")
    ap.add_argument("--episodes-in-flight", type=int, default=1,
                    help="independent episodes batched into one agent-step call per GPU (1 = the reference's serial loop)")
    ap.add_argument("--async-sim", action="store_true",
                    help="simulator worker threads feed a batching GPU server (run_units_async) instead of advancing the "
                         "episodes in flight in lock-step")
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's INIT lines (communicator size) stay visible: they prove how many ranks joined; the JSON line is printed last
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):   # (the image presets VERSION)
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            # ... on stderr: NCCL also logs at teardown, and the JSON line must stay the last line of stdout
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    tiny = args.preset == "tiny"
    ucfg, vcfg, acfg = (UNetConfig.tiny(), VAEConfig.tiny(), ACTConfig.tiny()) if tiny else (UNetConfig(), VAEConfig(), ACTConfig())
    shapes = dict(unet=W.unet_shapes(ucfg), controlnet=W.controlnet_shapes(ucfg), vae=W.vae_decoder_shapes(vcfg),
                  act=W.act_shapes(acfg))
    host = None
    if rank == 0:
        host = dict(unet=W.synth_state_dict(shapes["unet"]), controlnet=W.synth_state_dict(shapes["controlnet"], 1),
                    vae=W.synth_state_dict(shapes["vae"], 2), act=W.synth_state_dict(shapes["act"], 3))
    sds, _ = gd.broadcast_weights(shapes, host, src=0, device=dev)      # the ONE collective of the whole evaluation
    ops = Ops(local)
    pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg)
    step = GenimaStep(pipe, DeviceACT(ops, sds["act"], acfg), num_inference_steps=args.denoise_steps)
    S = acfg.image_size
    ctx = torch.randn(1, 77, ucfg.cross_attention_dim, generator=torch.Generator().manual_seed(3)).half().to(dev)
    task_emb = torch.randn(1, acfg.task_emb_dim, generator=torch.Generator().manual_seed(4)).to(dev)
    E = max(1, args.episodes_in_flight)
    gens = [torch.Generator(device=dev) for _ in range(E)]            # one noise stream per in-flight episode
    pin_v = torch.empty(E, 4, S, S, 3, dtype=torch.uint8).pin_memory()
    pin_q = torch.empty(E, acfg.state_dim, dtype=torch.float32).pin_memory()
    ctx_e, task_e = ctx.repeat(E, 1, 1), task_emb.repeat(E, 1)

    zero_lat = torch.zeros((1, 4, S // 4, S // 4), device=dev, dtype=torch.float16)

    def agent_step_batched(views, qpos, _k, active):
        pin_v.copy_(torch.from_numpy(views))
        pin_q.copy_(torch.from_numpy(qpos).reshape(E, -1))
        # diffusers prepare_latents, one generator per episode: every episode sees the noise it would see on its own (an
        # idle slot draws nothing, so its generator does not move while its simulator is busy)
        lat = torch.cat([torch.randn((1, 4, S // 4, S // 4), generator=g, device=dev, dtype=torch.float16) if on else zero_lat
                         for g, on in zip(gens, active)])
        out = step(pin_v.to(dev, non_blocking=True), lat, pin_q.to(dev, non_blocking=True), task_e, prompt_embeds=ctx_e)
        return out["a_hat"].float().cpu().numpy()                                                    # the step's sync point

    def agent_step(views, qpos, k):
        return agent_step_batched(views[None], qpos[None], k, [True])[0]

    units = gd.shard_units(RLBENCH_25[:args.tasks], args.episodes, rank, world)
    w_views, w_qpos = StubEnv("warm", 0, S).observe()
    if world > 1:
        # rank 0 measures the GEMM tile configurations on its first pass, every other rank adopts them BEFORE it captures
        # its graph: all ranks then launch identical kernels and an episode gives the same bits wherever it is sharded
        if rank == 0:
            agent_step_batched(np.stack([w_views] * E), np.stack([w_qpos] * E), 0, [True] * E)
        dist.barrier(device_ids=[local])
        gd.sync_tune_caches(pipe.all_ops(), src=0)
    agent_step_batched(np.stack([w_views] * E), np.stack([w_qpos] * E), 0, [True] * E)   # graph capture, untimed
    if world > 1:
        dist.barrier(device_ids=[local])
    t0 = time.perf_counter()
    if E == 1:
        recs = run_units(units, agent_step, size=S, episode_length=args.episode_length,
                         reseed=lambda s: gens[0].manual_seed(s))
    elif args.async_sim:
        recs = run_units_async(units, agent_step_batched, E, size=S, episode_length=args.episode_length,
                               reseed=lambda slot, s: gens[slot].manual_seed(s))
    else:
        recs = run_units_batched(units, agent_step_batched, E, size=S, episode_length=args.episode_length,
                                 reseed=lambda slot, s: gens[slot].manual_seed(s))
    torch.cuda.synchronize()
    wall = gd.reduce_max(time.perf_counter() - t0, device=dev)
    allrecs = gd.gather_records(recs)
    # which physical GPUs took part (one distinct UUID per rank proves the sharding really used `world` devices)
    uuids = gd.gather_records([{"rank": rank, "gpu_uuid": str(torch.cuda.get_device_properties(local).uuid),
                                "episodes": len(recs)}])
    if rank == 0:
        out = summarize(allrecs, wall, world)
        out["gpus_active"] = len({u["gpu_uuid"] for u in uuids})
        out["episodes_per_rank"] = [u["episodes"] for u in sorted(uuids, key=lambda u: u["rank"])]
        out["episodes_in_flight_per_gpu"] = E
        out["scheduling"] = "serial" if E == 1 else ("async sim workers -> batching server" if args.async_sim else "lock-step")
        print(json.dumps(out), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump({"summary": out, "eval_episodes": allrecs}, f)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
