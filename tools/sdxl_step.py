"""Full-size SDXL-ControlNet sibling (controller/agent/sdxl_controlnet_agent.py) on synthetic weights:
device time of the pipeline call (ControlNet-conditioned denoise loop with Euler-ancestral steps + KL-VAE decode), one
CUDA graph, 512 x 512 tile.  (Parity of the same model: tests/test_gpu_networks.py::test_sdxl_unet_controlnet_full_size_one_step.)
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

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
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
                  "denoise_steps": n_steps,
                  "ms_per_call_incl_noise_draws": round(ms, 3), "launches": pipe.last_graph_launches,
                  "image_shape": list(img.shape)}), flush=True)
