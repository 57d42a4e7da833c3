"""Host-side time breakdown of the reference-facing step (the loop body of controller/eval_genima.py:163-249 as
bench.py's e2e leg runs it) with phase timers and cProfile: where the e2e figure loses time against the device-resident
graph replay.  Usage (GPU box): python tools/e2e_breakdown.py > gpurun_out/e2e.txt"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = argparse.Namespace(preset="sd-turbo", autoencoder="", denoise_steps=5)
w = bench.build_world(args)
T = {}


def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0


e2e_step, _ = bench.make_e2e_step(w.pipe, w.ops, w.controller, w.views, w.qpos, w.prompts, w.negative_prompts, w.lang_np,
                                  w.acfg, 5, w.dev, 0, tick=tick)
with torch.inference_mode():
    for _ in range(3):
        e2e_step()
    T.clear()
    n = 20
    w0 = time.perf_counter()
    for _ in range(n):
        e2e_step()
    torch.cuda.synchronize()
    tot = (time.perf_counter() - w0) / n * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        w.step(w.d_views, w.d_lat, w.d_qpos, w.d_task, prompt_embeds=w.d_ctx)
    e1.record()
    torch.cuda.synchronize()
    print(f"e2e {tot:.2f} ms/step; device-resident fused step {e0.elapsed_time(e1) / n:.2f} ms/step")
    for k, v in T.items():
        print(f"  {k:36s} {v / n * 1e3:7.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        e2e_step()
    pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats("tottime").print_stats(30)
print("\n".join(s.getvalue().splitlines()[:60]))
s2 = io.StringIO()
pstats.Stats(pr, stream=s2).print_callers("torch.empty")   # who allocates on the timed path
print("\n".join(s2.getvalue().splitlines()[:30]))
