"""Runs ONE eager (non-graph) full-size agent step between cudaProfilerStart/Stop, for ncu launch lists:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/one_step.py [--denoise-steps 5] [--graph]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genima_b200 import distributed as gd  # noqa: E402
from genima_b200.act_policy import DeviceACT  # noqa: E402
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.pipeline import B200ControlNetPipeline  # noqa: E402
from genima_b200.step import GenimaStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--denoise-steps", type=int, default=5)
ap.add_argument("--preset", default="sd-turbo")
ap.add_argument("--graph", action="store_true")
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
ucfg, vcfg, acfg = bench.presets(args.preset)
shapes = bench.model_shapes(ucfg, vcfg, acfg)
dev = torch.device("cuda", 0)
sds, arena = gd.broadcast_weights(shapes, bench.synth_all(shapes), device=dev)
ops = Ops(0)
pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg)
act = DeviceACT(ops, sds["act"], acfg)
step = GenimaStep(pipe, act, num_inference_steps=args.denoise_steps, use_cuda_graph=args.graph)
views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
d = dict(views=views.permute(0, 2, 3, 1).contiguous()[None].to(dev), lat=lat.to(dev), qpos=qpos.to(dev),
         task=task.to(dev), ctx=ctx.to(dev))
for _ in range(2):
    step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(args.reps):
    out = step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches per step:", step.launches_per_step, "a_hat[0,0]:", out["a_hat"][0, 0].tolist())
