"""In-situ (warm, CUDA-graph replay) per-kernel durations of the agent step through torch.profiler (CUPTI).
Usage: python tools/profile_step.py [--eager] [--reps 5]  -> prints a table and writes gpurun_out/kineto_step.txt"""
import argparse
import collections
import os
import sys

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
ap.add_argument("--eager", action="store_true")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/kineto_step.txt")
args = ap.parse_args()
w = bench.build_world(argparse.Namespace(preset="sd-turbo", autoencoder="", denoise_steps=5))
pipe, act = w.pipe, w.act
step = GenimaStep(pipe, act, num_inference_steps=5, use_cuda_graph=not args.eager)
d = dict(views=w.d_views, lat=w.d_lat, qpos=w.d_qpos, task=w.d_task, ctx=w.d_ctx)
for _ in range(3):
    step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.reps):
        step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        name = ev.name.split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        tr = ev.time_range
        t_min = tr.start if t_min is None else min(t_min, tr.start)
        t_max = tr.end if t_max is None else max(t_max, tr.end)
tot = sum(v[1] for v in agg.values())
lines = [f"per step: kernel time {tot / args.reps / 1e3:.2f} ms over {sum(v[0] for v in agg.values()) // args.reps} kernels; "
         f"span {(t_max - t_min) / args.reps / 1e3:.2f} ms ({'eager' if args.eager else 'CUDA graph replay'}, warm)"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k[:60]:60s} n/step={v[0] // args.reps:5d} us/step={v[1] / args.reps:10.1f} share={v[1] / tot:.3f} avg={v[1] / v[0]:.2f}")
# per-grid breakdown of the GEMM kernel from the chrome trace (kineto records grid / block per kernel)
import json, tempfile
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
with open(tmp) as f:
    tr = json.load(f)
g = collections.defaultdict(lambda: [0, 0.0])
for ev in tr.get("traceEvents", []):
    if ev.get("cat") == "kernel" and "gemm_tc" in ev.get("name", ""):
        a = ev.get("args", {})
        key = (tuple(a.get("grid", [])), a.get("shared memory", 0))
        g[key][0] += 1
        g[key][1] += ev.get("dur", 0.0)
lines.append("--- gemm_tc_kernel by (grid, dynamic smem): n/step, us/step, avg us")
for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append(f"{str(k):40s} n={v[0] // args.reps:4d} us={v[1] / args.reps:9.1f} avg={v[1] / v[0]:7.2f}")
os.remove(tmp)
print("\n".join(lines))
os.makedirs(os.path.dirname(args.out), exist_ok=True)
with open(args.out, "w") as f:
    f.write("\n".join(lines) + "\n")
