"""Throughput of the fused agent step when B independent episodes are batched into one denoise call (SURVEY.md §8d
config 2 'B_tile' variant).  Usage: python tools/batch_step.py [B ...]"""
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

Bs = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
ucfg, vcfg, acfg = bench.presets("sd-turbo")
shapes = bench.model_shapes(ucfg, vcfg, acfg)
dev = torch.device("cuda", 0)
sds, arena = gd.broadcast_weights(shapes, bench.synth_all(shapes), device=dev)
ops = Ops(0)
pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg)
act = DeviceACT(ops, sds["act"], acfg)
step = GenimaStep(pipe, act, num_inference_steps=5, use_cuda_graph=True)
views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
for B in Bs:
    v = views.permute(0, 2, 3, 1).contiguous()[None].repeat(B, 1, 1, 1, 1).to(dev)
    args = (v, lat.repeat(B, 1, 1, 1).to(dev), qpos.repeat(B, 1).to(dev), task.repeat(B, 1).to(dev))
    c = ctx.repeat(B, 1, 1).to(dev)
    try:
        for _ in range(3):
            out = step(*args, prompt_embeds=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            out = step(*args, prompt_embeds=c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        a = out["a_hat"].float()
        same = float((a - a[0:1]).abs().max())
        print(f"B={B}: {ms:.2f} ms per call -> {B * 1e3 / ms:.1f} agent steps/s; max |a_hat[b] - a_hat[0]| = {same:.3e}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"B={B}: FAILED {type(ex).__name__}: {str(ex)[:300]}", flush=True)
