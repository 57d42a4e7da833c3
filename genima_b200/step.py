"""One Genima agent step, device-resident: the body of the reference's hot loop between `obs` and `actions`
(controller/eval_genima.py:163-249) with both PIL / numpy round trips removed:

    views u8 [B, 4, S, S, 3] --gn_tile_views--> tile u8 [B, 2S, 2S, 3]          (tile_images, controller/utils/misc.py:6-19)
      --B200ControlNetPipeline device chain--> generated tile u8                 (diffusion_agent.infer, eval_genima.py:203-210)
      --gn_untile_views--> generated views u8 [B, 4, S, S, 3]                    (untile_images, misc.py:22-47; the
                                                                                  Resize+CenterCrop is an identity at S)
      --DeviceACT.forward--> a_hat fp32 [B, 20, 8]                               (controller_agent.act, eval_genima.py:243-247)

The whole chain (about 1.5 k kernel launches at 5 denoise steps) is captured once into a CUDA graph and replayed per
step; inputs are copied into static buffers, so a step costs one graph launch.  Results are identical to calling the
pipeline and the policy through their reference-facing signatures (tests/test_gpu_step.py checks bit-equality).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .act_policy import DeviceACT
from .pipeline import B200ControlNetPipeline
from .unet import LATENT_CPAD, tensor_key


class GenimaStep:
    def __init__(self, pipe: B200ControlNetPipeline, act: DeviceACT, num_inference_steps: int = 5,
                 use_cuda_graph: bool = True):
        self.pipe, self.act, self.ops = pipe, act, pipe.ops
        self.n_steps = int(num_inference_steps)
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[tuple, dict] = {}
        self.launches_per_step = 0

    def _chain(self, views_u8, lat_nchw, qpos, task_emb, kv, tk):
        ops, pipe = self.ops, self.pipe
        tile = ops.tile_views(views_u8)
        lat_in = ops.nchw_to_nhwc(lat_nchw, cpad=LATENT_CPAD)
        _, img = pipe._denoise_and_decode(tile, lat_in, kv, tk, self.n_steps, True, 1.0)
        gen_tile = ops.nhwc_to_u8(img)
        gen_views = ops.untile_views(gen_tile)
        a_hat, is_pad = self.act.forward(qpos, gen_views, task_emb)
        return a_hat, is_pad, gen_tile

    @torch.no_grad()
    def __call__(self, views_u8: torch.Tensor, latents: torch.Tensor, qpos: torch.Tensor, task_emb: torch.Tensor,
                 prompt_embeds: Optional[torch.Tensor] = None, prompt=None):
        """views_u8 [B, 4, S, S, 3] uint8, latents [B, 4, S/4, S/4] fp16/fp32 unit-variance noise, qpos [B, state] fp32,
        task_emb [B, E] fp32 — all on the device.  Returns dict(a_hat, is_pad_hat, tile_u8), device tensors that are
        overwritten by the next call when the CUDA graph is in use."""
        ops, pipe = self.ops, self.pipe
        for name, t in (("views_u8", views_u8), ("latents", latents), ("qpos", qpos), ("task_emb", task_emb)):
            if not t.is_cuda:
                raise TypeError(f"{name} must already be on the device (this is the device-resident step)")
        B = views_u8.shape[0]
        ctx = pipe.encode_prompt(prompt, prompt_embeds)
        if ctx.shape[0] != B:
            ctx = pipe._expand_ctx(ctx, B)
        kv = pipe._context_kv(ctx)
        tk = ctx.shape[1]
        pipe.schedule.set_timesteps(self.n_steps)
        if latents.dtype not in (torch.float16, torch.float32):
            raise TypeError("latents must be fp16 or fp32")
        if not self.use_cuda_graph:
            l0 = pipe.launch_count()
            a_hat, is_pad, gen_tile = self._chain(views_u8.contiguous(), latents.contiguous(), qpos, task_emb, kv, tk)
            self.launches_per_step = pipe.launch_count() - l0
            return dict(a_hat=a_hat, is_pad_hat=is_pad, tile_u8=gen_tile)

        key = (tuple(views_u8.shape), tuple(latents.shape), latents.dtype, id(kv), tensor_key(task_emb), self.n_steps)
        g = self._graphs.get(key)
        if g is None:
            st = dict(views=views_u8.clone(), lat=latents.clone(), qpos=qpos.to(torch.float32).clone(), task=task_emb)
            temb = pipe._time_rows(self.n_steps, B)
            film = self.act.film_affines(task_emb)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):   # warm-up outside capture: smem attributes, per-shape caches, allocator
                self._chain(st["views"], st["lat"], st["qpos"], st["task"], kv, tk)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = pipe.launch_count()
            with torch.cuda.graph(graph):
                a_hat, is_pad, gen_tile = self._chain(st["views"], st["lat"], st["qpos"], st["task"], kv, tk)
            # the entry owns every buffer the graph reads by raw pointer (static inputs, cross-attention K/V and their
            # context, per-step time-embedding rows, FiLM affines): cache evictions elsewhere cannot free them
            g = dict(graph=graph, st=st, a_hat=a_hat, is_pad=is_pad, tile=gen_tile, kv=kv, ctx=ctx, temb=temb, film=film,
                     launches=pipe.launch_count() - l0)
            if len(self._graphs) > 4:
                self._graphs.clear()
            self._graphs[key] = g
        g["st"]["views"].copy_(views_u8, non_blocking=True)
        g["st"]["lat"].copy_(latents, non_blocking=True)
        g["st"]["qpos"].copy_(qpos, non_blocking=True)
        g["graph"].replay()
        self.launches_per_step = g["launches"]
        return dict(a_hat=g["a_hat"], is_pad_hat=g["is_pad"], tile_u8=g["tile"])
