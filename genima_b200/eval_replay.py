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
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"
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
    gen = torch.Generator(device=dev)
    pin_v = torch.empty(1, 4, S, S, 3, dtype=torch.uint8).pin_memory()
    pin_q = torch.empty(1, acfg.state_dim, dtype=torch.float32).pin_memory()

    def agent_step(views, qpos, _k):
        pin_v.copy_(torch.from_numpy(views)[None])
        pin_q.copy_(torch.from_numpy(qpos))
        lat = torch.randn((1, 4, S // 4, S // 4), generator=gen, device=dev, dtype=torch.float16)   # diffusers prepare_latents
        out = step(pin_v.to(dev, non_blocking=True), lat, pin_q.to(dev, non_blocking=True), task_emb, prompt_embeds=ctx)
        return out["a_hat"][0].float().cpu().numpy()                                                 # the step's sync point

    units = gd.shard_units(RLBENCH_25[:args.tasks], args.episodes, rank, world)
    agent_step(*StubEnv("warm", 0, S).observe(), 0)                     # graph capture outside the timed region
    if world > 1:
        dist.barrier(device_ids=[local])
    t0 = time.perf_counter()
    recs = run_units(units, agent_step, size=S, episode_length=args.episode_length, reseed=lambda s: gen.manual_seed(s))
    torch.cuda.synchronize()
    wall = gd.reduce_max(time.perf_counter() - t0, device=dev)
    allrecs = gd.gather_records(recs)
    if rank == 0:
        out = summarize(allrecs, wall, world)
        print(json.dumps(out), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump({"summary": out, "eval_episodes": allrecs}, f)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
