"""Host-side time breakdown of the reference-facing step (B200ControlNetAgent.infer + untile_images + B200GenimaACT.act)
with cProfile: where the e2e figure of bench.py loses time against the device-resident graph replay."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genima_b200 import distributed as gd  # noqa: E402
from genima_b200.act_policy import DeviceACT  # noqa: E402
from genima_b200.agents import B200ControlNetAgent, B200GenimaACT, B200GenimaACTPolicy  # noqa: E402
from genima_b200.host_glue import tile_images, untile_images  # noqa: E402
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.pipeline import B200ControlNetPipeline  # noqa: E402
from PIL import Image  # noqa: E402

ucfg, vcfg, acfg = bench.presets("sd-turbo")
shapes = bench.model_shapes(ucfg, vcfg, acfg)
dev = torch.device("cuda", 0)
sds, arena = gd.broadcast_weights(shapes, bench.synth_all(shapes), device=dev)
ops = Ops(0)
pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg, use_cuda_graph=True)
act = DeviceACT(ops, sds["act"], acfg)
views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
S = acfg.image_size
d_task, d_ctx = task.to(dev), ctx.to(dev)
agent = B200ControlNetAgent.__new__(B200ControlNetAgent)
agent.eval_cfg = dict(image_resolution=2 * S, device="cuda:0")
agent.pipe, agent._ops = pipe, ops
agent.set_optimizations()
agent.common_setup()
policy = B200GenimaACTPolicy.__new__(B200GenimaACTPolicy)
policy.cfg, policy.ops, policy._sd, policy.impl, policy.training = acfg, ops, sds["act"], act, False
policy.use_cuda_graph = True
controller = B200GenimaACT(policy)
lang_tokens = torch.zeros(1, 1, 77, dtype=torch.int32)
controller._emb_cache[lang_tokens.reshape(-1, 77).numpy().tobytes()] = (d_task, None)
cameras = ["wrist", "front", "right_shoulder", "left_shoulder"]
obs_np = {f"{c}_rgb": views[i:i + 1].numpy() for i, c in enumerate(cameras)}
low_dim = qpos.numpy()[None]
gen = [torch.Generator(device=dev).manual_seed(2)]
T = {}


def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0


def e2e_step():
    t0 = time.perf_counter()
    rgbs = [Image.fromarray(np.transpose(obs_np[f"{c}_rgb"][0], (1, 2, 0))) for c in cameras]
    tiles = tile_images(rgbs, 1)
    tick("tile (PIL)", t0)
    t0 = time.perf_counter()
    target = agent.infer(images=tiles, prompts=None, negative_prompts=None, prompt_embeds=d_ctx, num_inference_steps=5,
                         guidance_scale=0.0, generator=gen * len(tiles))
    tick("infer (incl. GPU)", t0)
    t0 = time.perf_counter()
    un = untile_images(target[0], cameras, agent.transform_to_half_resolution)
    tick("untile (PIL)", t0)
    t0 = time.perf_counter()
    obs = {f"{c}_rgb": un[c] for c in cameras}
    obs["low_dim_state"] = low_dim
    obs = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev).unsqueeze(0) for k, v in obs.items()}
    obs["lang_tokens"] = lang_tokens
    tick("obs -> device", t0)
    t0 = time.perf_counter()
    actions = controller.act(obs, step=0, eval_mode=True)[0]
    actions = actions.detach().cpu().numpy()
    tick("act (incl. GPU)", t0)
    return actions


for _ in range(3):
    e2e_step()
T.clear()
n = 20
w0 = time.perf_counter()
for _ in range(n):
    e2e_step()
torch.cuda.synchronize()
tot = (time.perf_counter() - w0) / n * 1e3
print(f"e2e {tot:.2f} ms/step")
for k, v in T.items():
    print(f"  {k:22s} {v / n * 1e3:7.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    e2e_step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print("\n".join(s.getvalue().splitlines()[:40]))
