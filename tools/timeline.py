"""Per-kernel timeline of one graph-replayed agent step (kineto): start, duration, gap to the previous kernel of the
same stream.  Usage: python tools/timeline.py [--first N] [--skip K]  -> gpurun_out/timeline.txt"""
import argparse
import collections
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genima_b200 import distributed as gd  # noqa: E402
from genima_b200.act_policy import DeviceACT  # noqa: E402
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.pipeline import B200ControlNetPipeline  # noqa: E402
from genima_b200.step import GenimaStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=400)
ap.add_argument("--skip", type=int, default=0)
ap.add_argument("--out", default="gpurun_out/timeline.txt")
args = ap.parse_args()
ucfg, vcfg, acfg = bench.presets("sd-turbo")
shapes = bench.model_shapes(ucfg, vcfg, acfg)
dev = torch.device("cuda", 0)
sds, arena = gd.broadcast_weights(shapes, bench.synth_all(shapes), device=dev)
ops = Ops(0)
pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg)
act = DeviceACT(ops, sds["act"], acfg)
step = GenimaStep(pipe, act, num_inference_steps=5, use_cuda_graph=True)
views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
d = dict(views=views.permute(0, 2, 3, 1).contiguous()[None].to(dev), lat=lat.to(dev), qpos=qpos.to(dev),
         task=task.to(dev), ctx=ctx.to(dev))
for _ in range(3):
    step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
    torch.cuda.synchronize()
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
with open(tmp) as f:
    tr = json.load(f)
evs = [e for e in tr.get("traceEvents", []) if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
evs.sort(key=lambda e: e["ts"])
t0 = evs[0]["ts"]
last_end = collections.defaultdict(lambda: None)
lines = []
tot_gap = collections.defaultdict(float)
tot_dur = collections.defaultdict(float)
for i, e in enumerate(evs):
    a = e.get("args", {})
    st = a.get("stream", 0)
    gap = (e["ts"] - last_end[st]) if last_end[st] is not None else 0.0
    last_end[st] = e["ts"] + e.get("dur", 0.0)
    tot_gap[st] += max(gap, 0.0)
    tot_dur[st] += e.get("dur", 0.0)
    if args.skip <= i < args.skip + args.first:
        name = e["name"].split("(")[0].replace("void ", "").replace("gn::", "")[:28]
        lines.append(f"{i:5d} s{st:<3} t={e['ts'] - t0:9.1f} dur={e.get('dur', 0.0):7.2f} gap={gap:6.2f} "
                     f"{name:28s} grid={a.get('grid')} smem={a.get('shared memory', 0)}")
lines.append(f"total span {evs[-1]['ts'] + evs[-1].get('dur', 0) - t0:.1f} us; per stream: "
             + "; ".join(f"s{k}: busy {tot_dur[k]:.0f} us, gaps {tot_gap[k]:.0f} us" for k in tot_dur))
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
