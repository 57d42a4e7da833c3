"""Device time of the agent step with the InstructPix2Pix sibling (controller/agent/sd_pix2pix_agent.py)
in place of the ControlNet pipeline: VAE encode + 5 x U-Net (8-channel input) + KL-VAE decode + untile + ACT, one CUDA
graph.  Prints a breakdown (VAE encode alone, whole step).  Usage: python tools/pix2pix_step.py"""
import dataclasses
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genima_b200 import weights as W  # noqa: E402
from genima_b200.act_policy import DeviceACT  # noqa: E402
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.pipeline import B200Pix2PixPipeline  # noqa: E402
from genima_b200.step import GenimaStep  # noqa: E402

ucfg, vcfg, acfg = bench.presets("sd-turbo")
ucfg = dataclasses.replace(ucfg, in_channels=8)
dev = torch.device("cuda", 0)
vsd = W.synth_state_dict(W.vae_decoder_shapes(vcfg), salt=2)
vsd.update(W.synth_state_dict(W.vae_encoder_shapes(vcfg), salt=2))
ops = Ops(0)
pipe = B200Pix2PixPipeline(ops, W.synth_state_dict(W.unet_shapes(ucfg), salt=5), vsd, None, ucfg, vcfg)
act = DeviceACT(ops, W.synth_state_dict(W.act_shapes(acfg), salt=3), acfg)
step = GenimaStep(pipe, act, num_inference_steps=5, use_cuda_graph=True)
views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
v = views.permute(0, 2, 3, 1).contiguous()[None].to(dev)
args = (v, lat.to(dev), qpos.to(dev), task.to(dev))
c = ctx.to(dev)


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(lambda: step(*args, prompt_embeds=c))
tile = ops.tile_views(v)
g = torch.cuda.CUDAGraph()
pipe.vae_enc_impl.encode(tile)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    pipe.vae_enc_impl.encode(tile)
enc_ms = timed(g.replay)
out = step(*args, prompt_embeds=c)
print(json.dumps({"pipeline": "instruct-pix2pix (no ControlNet, 8-channel U-Net, KL-VAE encode + decode)",
                  "ms_per_agent_step": round(ms, 3), "agent_steps_per_sec": round(1e3 / ms, 2),
                  "vae_encode_ms": round(enc_ms, 3), "launches_per_step": step.launches_per_step,
                  "a_hat_finite": bool(torch.isfinite(out["a_hat"]).all())}), flush=True)
