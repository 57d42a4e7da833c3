"""Full-size SDXL-ControlNet sibling (controller/agent/sdxl_controlnet_agent.py) on synthetic weights:
  1. parity of one ControlNet + U-Net evaluation (2.57 B + 1.25 B parameters) against the oracle graph (oracle/sd_models.py)
     executed in fp32 ON THE GPU (the CPU would need minutes per evaluation; TF32 disabled), with stock torch fp16
     running the same graph as the yardstick;
  2. device time of the pipeline chain (ControlNet-conditioned denoise loop with Euler-ancestral steps + KL-VAE decode),
     one CUDA graph, 512 x 512 tile.
Usage: python tools/sdxl_step.py [denoise_steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200 import weights as W  # noqa: E402
from genima_b200.configs import UNetConfig, VAEConfig  # noqa: E402
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.pipeline import B200SDXLControlNetPipeline  # noqa: E402
from oracle import sd_models  # noqa: E402  (checker only: this is a measurement / parity tool, not the product path)

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
ucfg, vcfg = UNetConfig.sdxl(), VAEConfig(scaling_factor=0.13025)
usd = W.synth_state_dict(W.unet_shapes(ucfg))
csd = W.synth_state_dict(W.controlnet_shapes(ucfg), salt=1)
vsd = W.synth_state_dict(W.vae_decoder_shapes(vcfg), salt=2)
ops = Ops(0)
pipe = B200SDXLControlNetPipeline(ops, usd, csd, vsd, None, None, ucfg, vcfg, use_cuda_graph=True)
g = torch.Generator().manual_seed(0)
cond = torch.randint(0, 256, (1, 512, 512, 3), generator=g, dtype=torch.uint8)
ctx = torch.randn(1, 77, 2048, generator=g).half()
pooled = torch.randn(1, 1280, generator=g).half()
x = torch.randn(1, 4, 64, 64, generator=g).half()


class _DevSD(dict):
    def __init__(self, sd, dtype):
        super().__init__()
        self._src, self._dtype = sd, dtype

    def __contains__(self, k):
        return k in self._src

    def __getitem__(self, k):
        t = self._src[k].to("cuda", self._dtype)

        class _T:
            def to(self_, *a, **kw):
                return t
        return _T()


def oracle_eps(dtype):
    added = dict(text_embeds=pooled.to("cuda", dtype), time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]], device="cuda"))
    tt = torch.tensor([999.0], device="cuda")
    c = (cond.float() / 255.0).permute(0, 3, 1, 2).to("cuda", dtype)
    xx, cc = x.to("cuda", dtype), ctx.to("cuda", dtype)
    down, mid = sd_models.controlnet_forward(_DevSD(csd, dtype), ucfg, xx, tt, cc, c, 1.0, added)
    return sd_models.unet_forward(_DevSD(usd, dtype), ucfg, xx, tt, cc, down, mid, added)


with torch.no_grad():
    ref = oracle_eps(torch.float32).float().cpu()
    stock = oracle_eps(torch.float16).float().cpu()
    unet, cn = pipe.unet_impl, pipe.controlnet_impl
    dadd = dict(text_embeds=pooled.cuda(), time_ids=[512.0, 512.0, 0.0, 0.0, 512.0, 512.0])
    ctx_d = ctx.cuda()
    kv_u = {tr.prefix: tr.project_context(ops, ctx_d) for tr in unet.transformers()}
    kv_c = {tr.prefix: tr.project_context(pipe.ops_side, ctx_d) for tr in cn.transformers()}
    tu = unet.temb_rows(unet.resblocks(), unet.time_embedding(999.0, dadd), 1)
    tc = cn.temb_rows(cn.resblocks(), cn.time_embedding(999.0, dadd), 1)
    for o in pipe.all_ops():
        o.gn_stats_reset()
    xs = ops.nchw_to_nhwc(x.cuda(), cpad=8)
    cmid, cskips = cn.encode(xs, cn.cond_embedding(pipe.ops_side.u8_to_nhwc(cond.cuda(), cpad=64)), tc, kv_c, 77)
    umid, uskips = unet.encode(xs, tu, kv_u, 77)
    torch.cuda.synchronize()
    skips, mid = cn.zero_convs(cmid, cskips, uskips, umid, 1.0)
    eps = torch.zeros_like(xs)
    unet.decode(mid, skips, tu, kv_u, 77, eps)
    out = eps[..., :4].permute(0, 3, 1, 2).float().cpu()
err = float((out - ref).abs().max() / ref.abs().max())
err_stock = float((stock - ref).abs().max() / ref.abs().max())

kw = dict(prompt_embeds=ctx, pooled_prompt_embeds=pooled, image=cond.cuda(), num_inference_steps=n_steps, guidance_scale=0.0,
          output_type="u8")
gen = torch.Generator(device="cuda").manual_seed(2)
for _ in range(3):
    img = pipe(generator=gen, **kw).images
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    img = pipe(generator=gen, **kw).images
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"pipeline": "sdxl-controlnet (2.57 B U-Net + 1.25 B ControlNet, Euler ancestral, KL-VAE), synthetic weights",
                  "denoise_steps": n_steps, "eps_err_vs_fp32_oracle_on_gpu": err, "stock_torch_fp16_err": err_stock,
                  "ms_per_call_incl_noise_draws": round(ms, 3), "launches": pipe.last_graph_launches,
                  "image_shape": list(img.shape)}), flush=True)
