"""B200ControlNetPipeline — drop-in for diffusers' StableDiffusionControlNetPipeline on Genima's eval path.

The reference calls (controller/agent/sd_controlnet_agent.py:67-76)

    self.pipe(prompt=..., image=..., negative_prompt=..., num_inference_steps=..., guidance_scale=..., generator=...)

and reads `out[0]` as a list of 512x512 PIL images (controller/eval_genima.py:215,225).  `__call__` below keeps the
diffusers 0.29.0 signature, implements the arguments the reference uses plus `latents`,
`prompt_embeds`, `output_type` and RAISES for everything it does not implement (CFG, ip-adapter, guess mode, control
guidance windows, multiple images per prompt) instead of silently ignoring it.

Execution (restated for the device):
  control image  u8 -> [B, 512, 512, 64] fp16 -> ControlNet conditioning embedding          (once per call)
  prompt ids     -> text encoder -> cross-attention K/V of all 23 transformer layers        (cached per prompt)
  timesteps      -> time-embedding MLP + per-ResBlock projections                           (cached per step count)
  for each sigma: U-Net encoder -> ControlNet encoder (+zero-convs fused with the skip adds) -> U-Net decoder ->
                  Euler update (+ next step's input scaling) — all on one stream, no host sync inside the loop
  latents / 0.18215 -> VAE decoder -> (x / 2 + 0.5).clamp -> u8
The whole post-upload chain can be captured once into a CUDA graph (`use_cuda_graph=True`) and replayed per call.

Sibling pipelines on the same device graphs, further down in this file:
  B200SDXLControlNetPipeline   diffusers StableDiffusionXLControlNetPipeline (controller/agent/sdxl_controlnet_agent.py)
  B200Pix2PixPipeline          diffusers StableDiffusionInstructPix2PixPipeline (controller/agent/sd_pix2pix_agent.py)
Both Euler schedulers the upstream snapshots ship are implemented (EulerDiscrete for sd-turbo, EulerAncestral for
sdxl-turbo: its per-step noise is drawn from the caller's generator in diffusers' order before the graph is replayed).
"""
from __future__ import annotations

import hashlib
import os
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .configs import CLIPTextConfig, SchedulerConfig, TAESDConfig, UNetConfig, VAEConfig
from .ops import Ops
from .scheduler import EulerDiscreteSchedule
from .text_encoder import DeviceCLIPText
from .unet import LATENT_CPAD, DeviceControlNet, DeviceUNet, tensor_key
from .vae import DeviceTAESDDecoder, DeviceVAEDecoder, DeviceVAEEncoder

try:  # PIL is only needed for the "pil" input/output types the reference uses
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None


class PipelineOutput:
    """StableDiffusionPipelineOutput look-alike: `.images`, `.nsfw_content_detected`, tuple-style indexing."""

    def __init__(self, images, nsfw_content_detected=None):
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected

    def __getitem__(self, i):
        return (self.images, self.nsfw_content_detected)[i]

    def __iter__(self):
        return iter((self.images, self.nsfw_content_detected))

    def __len__(self):
        return 2


class _ModuleShim:
    """Accepts the attribute pokes DiffusionAgent.set_optimizations makes on pipe.vae / pipe.unet (no-ops here: the
    kernels are already fused / sliced as needed)."""

    def __init__(self, impl=None):
        self.impl = impl

    def enable_slicing(self):
        return None

    def to(self, *a, **k):
        return self


class B200ControlNetPipeline:
    def __init__(self, ops: Ops, unet_sd, controlnet_sd, vae_sd, text_sd=None,
                 unet_cfg: UNetConfig = UNetConfig(), vae_cfg: VAEConfig = VAEConfig(),
                 text_cfg: CLIPTextConfig = CLIPTextConfig.sd_turbo(), scheduler_cfg: SchedulerConfig = SchedulerConfig(),
                 tokenizer: Optional[Callable[[Sequence[str]], torch.Tensor]] = None, use_cuda_graph: bool = False,
                 concurrent_controlnet: bool = True):
        self.ops = ops
        self.unet_cfg, self.vae_cfg, self.text_cfg = unet_cfg, vae_cfg, text_cfg
        # The ControlNet encoder and the U-Net encoder are independent until the zero-conv adds: they run concurrently
        # on two streams (both are chains of small latency-bound kernels at batch 1).  The ControlNet gets its own
        # gn_handle so that the per-handle GroupNorm scratch / grid-barrier words are never shared between streams.
        self.concurrent_controlnet = bool(concurrent_controlnet) and os.environ.get("GENIMA_B200_CONCURRENT", "1") != "0"   # A/B
        self.ops_side = Ops(ops.device.index, autotune=ops.autotune) if self.concurrent_controlnet else ops
        if self.concurrent_controlnet and os.environ.get("GENIMA_B200_SIDE_OCC"):   # A/B: ring sizing of the ControlNet's kernels
            self.ops_side.handle.check(self.ops_side.lib.gn_set_gemm_occupancy(
                self.ops_side.h, int(os.environ["GENIMA_B200_SIDE_OCC"])), "gn_set_gemm_occupancy")
        # third handle / stream: the 13 zero-convs start as soon as both encoders have produced their skip tensor,
        # instead of running one after the other once the two encoders have joined
        self.ops_zero = Ops(ops.device.index, autotune=ops.autotune) if self.concurrent_controlnet else ops
        self.zero_stream = torch.cuda.Stream(device=ops.device) if self.concurrent_controlnet else None
        self.overlap_zero_convs = os.environ.get("GENIMA_B200_ZERO_OVERLAP", "1") != "0"
        # scheduler step + input scaling fused into conv_out / conv_in epilogues (A/B: GENIMA_B200_FUSE_SCHED=0)
        self.fuse_scheduler = os.environ.get("GENIMA_B200_FUSE_SCHED", "1") != "0"
        self.side_stream = (torch.cuda.Stream(device=ops.device, priority=int(os.environ.get("GENIMA_B200_SIDE_PRIO", "0")))
                            if self.concurrent_controlnet else None)
        self.unet_impl = DeviceUNet(ops, unet_sd, unet_cfg)
        self.controlnet_impl = DeviceControlNet(self.ops_side, controlnet_sd, unet_cfg)
        # AutoencoderKL decoder, or AutoencoderTiny when the caller selected TAESD (vae_cfg is a TAESDConfig)
        self.vae_impl = (DeviceTAESDDecoder(ops, vae_sd, vae_cfg) if isinstance(vae_cfg, TAESDConfig)
                         else DeviceVAEDecoder(ops, vae_sd, vae_cfg))
        self.text_impl = DeviceCLIPText(ops, text_sd, text_cfg) if text_sd is not None else None
        self.schedule = EulerDiscreteSchedule(scheduler_cfg)
        self.tokenizer = tokenizer
        self.use_cuda_graph = use_cuda_graph
        self.vae_scale_factor = 2 ** (len(vae_cfg.block_out_channels) - 1)
        # attributes the reference pokes (controller/agent/diffusion_agent.py:21-42, sd_controlnet_agent.py:45-65)
        self.vae = _ModuleShim(self.vae_impl)
        self.unet = _ModuleShim(self.unet_impl)
        self.controlnet = _ModuleShim(self.controlnet_impl)
        self.text_encoder = _ModuleShim(self.text_impl)
        self.scheduler = self.schedule
        self._ctx_cache: Dict = {}
        self._kv_cache: Dict[str, Dict[str, torch.Tensor]] = {}
        self._temb_cache: Dict[tuple, list] = {}
        self._graphs: Dict[tuple, dict] = {}
        self._pinned: Dict[tuple, torch.Tensor] = {}
        self._tuned_shapes = set()
        self.progress_bar_disabled = True
        self._added = None      # SDXL added conditioning of the current call (text_embeds, time_ids); None for SD-2.x

    def launch_count(self) -> int:
        """Kernels launched through this pipeline's handle(s) since creation."""
        return sum(o.launch_count() for o in self.all_ops())

    def all_ops(self):
        return [self.ops] if self.ops_side is self.ops else [self.ops, self.ops_side, self.ops_zero]

    def tune_cache_export(self) -> List[bytes]:
        """Measured GEMM tile configurations of this pipeline's handles (see genima_b200.distributed.sync_tune_caches)."""
        return [o.tune_cache_export() for o in self.all_ops()]

    def tune_cache_import(self, blobs: Sequence[bytes]) -> None:
        for o, b in zip(self.all_ops(), blobs):
            o.tune_cache_import(b, replace=True)

    # ------------------------------------------------------------------ diffusers API surface used by the reference
    def to(self, *args, **kwargs):
        return self

    def set_progress_bar_config(self, **kwargs):
        self.progress_bar_disabled = bool(kwargs.get("disable", True))

    def upcast_vae(self):
        return None

    def fuse_qkv_projections(self, unet: bool = True, vae: bool = True):
        return None  # Q/K/V projections are always one GEMM here

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None  # attention is always the tcgen05 flash kernel

    # ------------------------------------------------------------------ prompt handling
    def _prompt_ids(self, prompt) -> torch.Tensor:
        if isinstance(prompt, torch.Tensor):
            ids = prompt
        elif isinstance(prompt, (list, tuple)) and len(prompt) and not isinstance(prompt[0], str):
            ids = torch.as_tensor(np.asarray(prompt))
        else:
            if isinstance(prompt, str):
                prompt = [prompt]
            if self.tokenizer is None:
                raise RuntimeError(
                    "string prompts need a CLIP tokenizer, and no BPE vocabulary is available offline: pass "
                    "`tokenizer=` to the pipeline, or token ids [B, 77] as `prompt`, or `prompt_embeds`")
            ids = self.tokenizer(list(prompt))
        ids = ids.to(torch.int64)
        if ids.dim() == 1:
            ids = ids[None]
        return ids

    def encode_prompt(self, prompt=None, prompt_embeds=None) -> torch.Tensor:
        """-> [B, 77, D] fp16 on the device (diffusers encode_prompt without CFG, clip_skip=None)."""
        if prompt_embeds is not None:
            # cached by tensor identity so that repeated calls with the same embeddings reuse the cross-attention K/V
            key = ("embeds",) + tensor_key(prompt_embeds)
            hit = self._ctx_cache.get(key)
            if hit is None:
                if len(self._ctx_cache) > 64:
                    self._ctx_cache.clear()
                hit = (prompt_embeds.to(self.ops.device, torch.float16).contiguous(), prompt_embeds)
                self._ctx_cache[key] = hit
            return hit[0]
        ids = self._prompt_ids(prompt)
        key = hashlib.sha1(ids.cpu().numpy().tobytes()).hexdigest()
        if key not in self._ctx_cache:
            if self.text_impl is None:
                raise RuntimeError("pipeline was built without a text encoder: pass prompt_embeds")
            if len(self._ctx_cache) > 64:
                self._ctx_cache.clear()
            self._ctx_cache[key] = self.text_impl(ids.to(self.ops.device))[0]
        return self._ctx_cache[key]

    def _expand_ctx(self, ctx: torch.Tensor, B: int) -> torch.Tensor:
        if ctx.shape[0] != 1:
            raise ValueError(f"{ctx.shape[0]} prompts for {B} control images")
        key = ("expand", B) + tensor_key(ctx)
        if key not in self._ctx_cache:
            self._ctx_cache[key] = (ctx.expand(B, -1, -1).contiguous(), ctx)
        return self._ctx_cache[key][0]

    def _context_kv(self, ctx: torch.Tensor) -> Dict[str, torch.Tensor]:
        key = tensor_key(ctx)
        hit = self._kv_cache.get(key)
        if hit is None or hit[1] is not ctx:
            if len(self._kv_cache) > 8:
                self._kv_cache.clear()      # (graphs that captured an evicted K/V set hold their own reference to it)
            kv = {}
            for net in (self.unet_impl, self.controlnet_impl):
                for tr in net.transformers():
                    kv[("u:" if net is self.unet_impl else "c:") + tr.prefix] = tr.project_context(self.ops, ctx)
            kv["__ctx__"] = ctx             # the entry keeps its context alive: the data_ptr in `key` cannot be recycled
            hit = self._kv_cache[key] = (kv, ctx)
        return hit[0]

    def _added_key(self) -> tuple:
        a = getattr(self, "_added", None)
        return () if a is None else tensor_key(a["text_embeds"]) + tuple(float(v) for v in a["time_ids"])

    def _time_rows(self, n_steps: int, batch: int):
        key = (n_steps, batch) + self._added_key()
        if key not in self._temb_cache:
            if len(self._temb_cache) > 16:
                self._temb_cache.clear()
            ts, _ = self.schedule.set_timesteps(n_steps)
            per_step = []
            for t in ts:
                su = self.unet_impl.time_embedding(float(t), self._added)
                sc = self.controlnet_impl.time_embedding(float(t), self._added)
                per_step.append((self.unet_impl.temb_rows(self.unet_impl.resblocks(), su, batch),
                                 self.controlnet_impl.temb_rows(self.controlnet_impl.resblocks(), sc, batch)))
            # scheduler.scale_model_input folded into the two conv_in epilogues: one constant fp32 vector per step
            _, sig = self.schedule.set_timesteps(n_steps)
            c0 = self.unet_cfg.block_out_channels[0]
            per_step = [(tu, tc, torch.full((c0,), 1.0 / float(np.sqrt(float(sig[i]) ** 2 + 1.0)), dtype=torch.float32,
                                            device=self.ops.device)) for i, (tu, tc) in enumerate(per_step)]
            self._temb_cache[key] = per_step
        return self._temb_cache[key]

    # ------------------------------------------------------------------ image handling
    def _control_image_u8(self, image) -> torch.Tensor:
        """-> uint8 [B, H, W, 3] on the device (VaeImageProcessor.preprocess without resize/normalise)."""
        if isinstance(image, torch.Tensor):
            t = image
            if t.dtype != torch.uint8:
                raise TypeError("tensor control images must be uint8 [B, H, W, 3]")
            if t.dim() == 3:
                t = t[None]
        else:
            if not isinstance(image, (list, tuple)):
                image = [image]
            arrs = []
            for im in image:
                if Image is not None and isinstance(im, Image.Image):
                    im = np.asarray(im if im.mode == "RGB" else im.convert("RGB"))
                arrs.append(np.asarray(im, dtype=np.uint8))
            shape = (len(arrs),) + arrs[0].shape
            if len(shape) != 4 or shape[-1] != 3 or any(a.shape != arrs[0].shape for a in arrs):
                raise ValueError(f"control images must all be [H, W, 3], got {[a.shape for a in arrs]}")
            # host images: written once, straight into a persistent pinned buffer (asynchronous DMA from there)
            buf = self._pinned_buffer("in", shape)
            dst = buf.numpy()
            for i, a in enumerate(arrs):
                np.copyto(dst[i], a)
            return buf.to(self.ops.device, non_blocking=True)
        if t.shape[-1] != 3:
            raise ValueError(f"control image must be [B, H, W, 3], got {tuple(t.shape)}")
        if t.is_cuda:
            return t.to(self.ops.device).contiguous()
        buf = self._pinned_buffer("in", tuple(t.shape))
        buf.copy_(t)
        return buf.to(self.ops.device, non_blocking=True)

    def _pinned_buffer(self, kind: str, shape) -> torch.Tensor:
        key = (kind, tuple(shape))
        buf = self._pinned.get(key)
        if buf is None:
            buf = self._pinned[key] = torch.empty(tuple(shape), dtype=torch.uint8).pin_memory()
        return buf

    def _to_host_u8(self, u8: torch.Tensor) -> np.ndarray:
        """Device uint8 image -> numpy through a pinned buffer; the stream synchronise here is the step's sync point."""
        key = ("out", tuple(u8.shape))
        buf = self._pinned.get(key)
        if buf is None:
            buf = self._pinned[key] = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
        buf.copy_(u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return buf.numpy().copy()

    # ------------------------------------------------------------------ the device chain
    def _scheduler_step(self, x: torch.Tensor, eps: torch.Tensor, i: int, noise: Optional[torch.Tensor]):
        """scheduler.step + the next step's scale_model_input in one kernel -> (x_next, x_scaled_next)."""
        sig = self.schedule.sigmas
        x_next, xs_next = torch.empty_like(x), torch.empty_like(x)
        if self.schedule.ddim:
            # DDIM (eta = 0): x' = a x + b eps, written as x + b eps + (a - 1) x on the ancestral-step kernel with the
            # sample itself in the noise slot (sigmas are all zero here, so x_scaled = x')
            a, b = self.schedule.ddim_coeffs(i)
            self.ops.euler_ancestral_step(x, eps, x, 0.0, b, a - 1.0, 0.0, x_next=x_next, x_scaled=xs_next)
        elif self.schedule.ancestral:
            if noise is None:
                raise RuntimeError("EulerAncestralDiscreteScheduler re-injects noise at every step: per-step noise "
                                   "[n_steps, B, h, w, 8] is required (the public __call__ draws it from `generator`)")
            up, down = self.schedule.ancestral_sigmas(i)
            self.ops.euler_ancestral_step(x, eps, noise[i], float(sig[i]), down, up, float(sig[i + 1]),
                                          x_next=x_next, x_scaled=xs_next)
        else:
            self.ops.euler_step(x, eps, float(sig[i]), float(sig[i + 1]), x_next=x_next, x_scaled=xs_next)
        return x_next, xs_next

    def _denoise_and_decode(self, cond_u8: torch.Tensor, lat_in: torch.Tensor, kv, tk: int, n_steps: int,
                            want_image: bool, cond_scale: float, noise: Optional[torch.Tensor] = None):
        """cond_u8 [B, H, W, 3]; lat_in [B, h, w, 8] fp16 (unit-variance noise).  Returns (latents, image fp16|None)."""
        ops = self.ops
        B = lat_in.shape[0]
        temb = self._time_rows(n_steps, B)
        # GroupNorm statistics are accumulated by the GEMM epilogues into per-handle arenas: one memset per forward pass
        # (issued here, before the two encoder streams fork)
        for o in self.all_ops():
            o.gn_stats_reset()
        sig = self.schedule.sigmas
        kv_u = {k[2:]: v for k, v in kv.items() if k.startswith("u:")}
        kv_c = {k[2:]: v for k, v in kv.items() if k.startswith("c:")}
        cond = ops.u8_to_nhwc(cond_u8, cpad=64)
        cond_emb = self.controlnet_impl.cond_embedding(cond)
        x = ops.scale(lat_in, self.schedule.init_noise_sigma)                       # latents * init_noise_sigma
        # Euler / DDIM: the whole scheduler lives in epilogues -- scale_model_input in the two conv_in's (per-step scale
        # vector), the update x' = a x + b eps in the U-Net's conv_out (alpha / beta of its residual add); the ancestral
        # variant re-injects noise and keeps its own kernel
        fused = self.fuse_scheduler and not self.schedule.ancestral and self.unet_impl.w_out8 is not None
        xs = eps = None
        if not fused:
            xs = ops.scale(x, 1.0 / float(np.sqrt(float(sig[0]) ** 2 + 1.0)))       # scale_model_input, step 0
            eps = torch.zeros_like(x)
        half_sms = max(1, ops.num_sms() // 2)
        # the first eager pass for a new shape runs the two encoders one after the other: that is when the GEMM tile
        # configurations are measured (gn_set_autotune), and a concurrent neighbour would disturb the timings
        shape_key = (tuple(lat_in.shape), n_steps)
        concurrent = self.concurrent_controlnet and (torch.cuda.is_current_stream_capturing()
                                                     or shape_key in self._tuned_shapes)
        self._tuned_shapes.add(shape_key)
        # two grid-barrier GroupNorm kernels may be in flight at once (one per stream): cap each at half the SMs so
        # that all their CTAs are always co-resident (no barrier deadlock).  The cap is applied in the sequential
        # mode too, so that results do not depend on the execution mode (the CTA count fixes the summation order).
        for o in self.all_ops():
            o.set_gn_max_ctas(half_sms)
        for i in range(n_steps):
            tu, tc, svec = temb[i]
            if fused:
                xs = x                                   # conv_in reads the unscaled sample, its epilogue scales
                if self.schedule.ddim:
                    svec, (sa, sb) = None, self.schedule.ddim_coeffs(i)
                else:
                    sa, sb = 1.0, float(sig[i + 1]) - float(sig[i])
                eps = torch.empty_like(x)                # receives x' directly
                sched = (x, sa, sb)
            else:
                svec, sched = None, None
            if concurrent and self.overlap_zero_convs:
                main = torch.cuda.current_stream()
                n_ev = len(self.controlnet_impl.zero_w) + 1
                ev_u = [torch.cuda.Event() for _ in range(n_ev)]
                ev_c = [torch.cuda.Event() for _ in range(n_ev)]
                self.side_stream.wait_stream(main)
                self.zero_stream.wait_stream(main)
                with torch.cuda.stream(self.side_stream):
                    cn_mid, cn_skips = self.controlnet_impl.encode(
                        xs, cond_emb, tc, kv_c, tk, on_skip=lambda j: ev_c[j].record(self.side_stream), in_scale=svec)
                mid, skips = self.unet_impl.encode(xs, tu, kv_u, tk, on_skip=lambda j: ev_u[j].record(main),
                                                   in_scale=svec)
                with torch.cuda.stream(self.zero_stream):
                    outs = []
                    for j in range(n_ev):
                        self.zero_stream.wait_event(ev_u[j])
                        self.zero_stream.wait_event(ev_c[j])
                        src_c, src_u = (cn_skips[j], skips[j]) if j < n_ev - 1 else (cn_mid, mid)
                        outs.append(self.controlnet_impl.zero_conv(j, src_c, src_u, cond_scale, ops=self.ops_zero))
                main.wait_stream(self.side_stream)
                main.wait_stream(self.zero_stream)
                skips, mid = outs[:-1], outs[-1]
                del cn_mid, cn_skips, outs
                self.unet_impl.decode(mid, skips, tu, kv_u, tk, eps, step=sched)
                x, xs = (eps, None) if fused else self._scheduler_step(x, eps, i, noise)
                continue
            if concurrent:
                main = torch.cuda.current_stream()
                self.side_stream.wait_stream(main)
                with torch.cuda.stream(self.side_stream):
                    cn_mid, cn_skips = self.controlnet_impl.encode(xs, cond_emb, tc, kv_c, tk, in_scale=svec)
                mid, skips = self.unet_impl.encode(xs, tu, kv_u, tk, in_scale=svec)
                main.wait_stream(self.side_stream)
            else:
                mid, skips = self.unet_impl.encode(xs, tu, kv_u, tk, in_scale=svec)
                cn_mid, cn_skips = self.controlnet_impl.encode(xs, cond_emb, tc, kv_c, tk, in_scale=svec)
            # (same handle as the overlapped path, so that its tile configurations are the ones measured here)
            skips, mid = self.controlnet_impl.zero_convs(cn_mid, cn_skips, skips, mid, cond_scale,
                                                         ops=self.ops_zero if self.overlap_zero_convs else None)
            del cn_mid, cn_skips
            self.unet_impl.decode(mid, skips, tu, kv_u, tk, eps, step=sched)
            x, xs = (eps, None) if fused else self._scheduler_step(x, eps, i, noise)
        for o in self.all_ops():
            o.set_gn_max_ctas(0)
        img = None
        if want_image:
            if isinstance(self.vae_impl, DeviceVAEDecoder):
                # `latents / scaling_factor` rides in post_quant_conv's epilogue
                img = self.vae_impl.decode(x, in_scale=1.0 / self.vae_cfg.scaling_factor)
            else:
                img = self.vae_impl.decode(ops.scale(x, 1.0 / self.vae_cfg.scaling_factor))
        return x, img

    def _run(self, cond_u8, lat_in, kv, tk, n_steps, want_image, cond_scale, noise=None):
        if not self.use_cuda_graph:
            return self._denoise_and_decode(cond_u8, lat_in, kv, tk, n_steps, want_image, cond_scale, noise)
        key = (tuple(cond_u8.shape), tuple(lat_in.shape), id(kv), tk, n_steps, want_image, cond_scale,
               noise is not None) + self._added_key()
        g = self._graphs.get(key)
        if g is None:
            static_cond = cond_u8.clone()
            static_lat = lat_in.clone()
            static_noise = noise.clone() if noise is not None else None
            temb = self._time_rows(n_steps, lat_in.shape[0])          # hoisted work must exist before capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                             # warm-up: tensor maps, smem attributes, allocator
                self._denoise_and_decode(static_cond, static_lat, kv, tk, n_steps, want_image, cond_scale, static_noise)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            launches0 = ops_launches = self.launch_count()
            with torch.cuda.graph(graph):
                out_lat, out_img = self._denoise_and_decode(static_cond, static_lat, kv, tk, n_steps, want_image,
                                                            cond_scale, static_noise)
            ops_launches = self.launch_count() - launches0
            # the entry owns everything the graph reads by raw pointer: static inputs, the K/V set (which holds its
            # context), the per-step time-embedding rows and the SDXL added conditioning -- clearing _temb_cache /
            # _kv_cache / _ctx_cache later cannot free memory under a graph that is still cached
            g = dict(graph=graph, cond=static_cond, lat=static_lat, noise=static_noise, out_lat=out_lat,
                     out_img=out_img, kv=kv, temb=temb, added=self._added, launches=ops_launches)
            if len(self._graphs) > 4:
                self._graphs.clear()
            self._graphs[key] = g
        g["cond"].copy_(cond_u8, non_blocking=True)
        g["lat"].copy_(lat_in, non_blocking=True)
        if noise is not None:
            g["noise"].copy_(noise, non_blocking=True)
        g["graph"].replay()
        self.last_graph_launches = g["launches"]
        return g["out_lat"], g["out_img"]

    # ------------------------------------------------------------------ public call
    @torch.no_grad()
    def __call__(self, prompt=None, image=None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, timesteps: Optional[List[int]] = None, guidance_scale: float = 7.5,
                 negative_prompt=None, num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None,
                 latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, ip_adapter_image=None,
                 ip_adapter_image_embeds=None, output_type: Optional[str] = "pil", return_dict: bool = True,
                 cross_attention_kwargs=None, controlnet_conditioning_scale: Union[float, List[float]] = 1.0,
                 guess_mode: bool = False, control_guidance_start: Union[float, List[float]] = 0.0,
                 control_guidance_end: Union[float, List[float]] = 1.0, clip_skip: Optional[int] = None,
                 callback_on_step_end=None, callback_on_step_end_tensor_inputs: List[str] = ["latents"], **kwargs):
        # ---- refuse what is not implemented (never silently ignore an argument that changes the result)
        if guidance_scale is not None and guidance_scale > 1.0:
            raise NotImplementedError("classifier-free guidance (guidance_scale > 1) is not implemented; Genima runs 0.0")
        if num_images_per_prompt not in (None, 1):
            raise NotImplementedError("num_images_per_prompt != 1 is not implemented")
        if guess_mode:
            raise NotImplementedError("guess_mode is not implemented")
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            raise NotImplementedError("ip-adapter inputs are not implemented")
        if control_guidance_start != 0.0 or control_guidance_end != 1.0:
            raise NotImplementedError("control_guidance_start/end windows are not implemented")
        if isinstance(controlnet_conditioning_scale, (list, tuple)):
            raise NotImplementedError("multi-ControlNet conditioning scales are not implemented")
        if timesteps is not None:
            raise NotImplementedError("custom `timesteps` are not implemented")
        if clip_skip is not None:
            raise NotImplementedError("clip_skip is not implemented")
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs are not implemented")
        if callback_on_step_end is not None:
            raise NotImplementedError("per-step callbacks would force a host sync inside the loop; not implemented")
        if kwargs:
            raise TypeError(f"unexpected arguments: {sorted(kwargs)}")
        if output_type not in ("pil", "np", "pt", "latent", "u8"):
            raise ValueError(f"output_type {output_type!r} is not supported")
        if image is None:
            raise ValueError("`image` (the ControlNet conditioning image) is required")
        # guidance_scale <= 1 -> no CFG: negative prompts are never encoded; eta is unused by Euler.
        return self._generate(prompt, image, height, width, num_inference_steps, generator, latents, prompt_embeds,
                              output_type, return_dict, float(controlnet_conditioning_scale))

    def _generate(self, prompt, image, height, width, num_inference_steps, generator, latents, prompt_embeds,
                  output_type, return_dict, cond_scale: float):
        """Shared body of the public calls: image / prompt / latents preparation -> device chain -> output conversion."""
        ops = self.ops
        cond_u8 = self._control_image_u8(image)
        B, H, W, _ = cond_u8.shape
        f = self.vae_scale_factor
        if (height not in (None, H)) or (width not in (None, W)):
            raise NotImplementedError("resizing the control image is not implemented: pass it at the target size")
        if H % (f * 8) or W % (f * 8):
            raise ValueError(f"image size {H}x{W} must be a multiple of {f * 8}")
        ctx = self.encode_prompt(prompt, prompt_embeds)
        if ctx.shape[0] != B:
            ctx = self._expand_ctx(ctx, B)
        kv = self._context_kv(ctx)
        tk = ctx.shape[1]

        h, w = H // f, W // f
        lc = self.vae_cfg.latent_channels
        if latents is None:
            gens = generator if isinstance(generator, (list, tuple)) else [generator] * B
            if len(gens) != B:
                raise ValueError(f"{len(gens)} generators for batch {B}")
            parts = []
            for g in gens:  # diffusers randn_tensor: one draw per sample on the generator's device, model dtype
                gdev = g.device if g is not None else ops.device
                parts.append(torch.randn((1, lc, h, w), generator=g, device=gdev, dtype=torch.float16).to(ops.device))
            latents = torch.cat(parts, dim=0)
        else:
            if tuple(latents.shape) != (B, lc, h, w):
                raise ValueError(f"latents must be {(B, lc, h, w)}, got {tuple(latents.shape)}")
            latents = latents.to(ops.device)
            if latents.dtype not in (torch.float16, torch.float32):
                latents = latents.float()
        lat_in = ops.nchw_to_nhwc(latents.contiguous(), cpad=LATENT_CPAD)

        n_steps = int(num_inference_steps)
        self.schedule.set_timesteps(n_steps)
        noise = None
        if self.schedule.ancestral:
            # EulerAncestralDiscreteScheduler.step draws randn_tensor(model_output.shape, dtype=model dtype,
            # generator=generator) at EVERY step (also the last one, where sigma_up = 0): same draws, same order, so a
            # generator shared across calls stays aligned with the reference's
            gens = generator if isinstance(generator, (list, tuple)) else [generator] * B
            if len(gens) != B:
                raise ValueError(f"{len(gens)} generators for batch {B}")
            draws = []
            for _ in range(n_steps):
                for g in gens:
                    gdev = g.device if g is not None else ops.device
                    draws.append(torch.randn((1, lc, h, w), generator=g, device=gdev, dtype=torch.float16).to(ops.device))
            noise = ops.nchw_to_nhwc(torch.cat(draws, dim=0).contiguous(), cpad=LATENT_CPAD)
            noise = noise.reshape(n_steps, B, h, w, LATENT_CPAD)
        x, img = self._run(cond_u8, lat_in, kv, tk, n_steps, output_type != "latent", cond_scale, noise)
        if img is not None and getattr(self.vae_cfg, "force_upcast", False):
            # this VAE asks for an fp32 decode upstream (config.json force_upcast: the stock SDXL VAE overflows fp16);
            # the fp16 decoder ran anyway -- refuse to hand back a NaN / black image
            if not bool(torch.isfinite(img).all()):
                raise FloatingPointError(
                    "the fp16 VAE decode produced non-finite values and the snapshot's VAE sets force_upcast=true: use an "
                    "fp16-safe VAE (madebyollin/sdxl-vae-fp16-fix, which the reference trains with, or taesdxl via "
                    "`autoencoder`)")

        if output_type == "latent":
            images = ops.nhwc_to_nchw(x, channels=lc)
        elif output_type == "u8":
            images = ops.nhwc_to_u8(img)                                    # device uint8 [B, H, W, 3]
        elif output_type == "pt":
            images = (ops.nhwc_to_nchw(img, channels=3, fp32=True) / 2 + 0.5).clamp(0, 1)
        else:
            # persistent device buffer: a fresh allocation here would sit behind the replayed graph in the allocator
            key = ("u8dev", tuple(img.shape[:3]))
            dbuf = self._pinned.get(key)
            if dbuf is None:
                dbuf = self._pinned[key] = torch.empty(tuple(img.shape[:3]) + (3,), dtype=torch.uint8, device=img.device)
            u8 = self._to_host_u8(ops.nhwc_to_u8(img, out=dbuf))
            if output_type == "np":
                images = u8.astype(np.float32) / 255.0
            else:
                if Image is None:
                    raise RuntimeError("PIL is not available for output_type='pil'")
                images = [Image.fromarray(a) for a in u8]
        if not return_dict:
            return (images, None)
        return PipelineOutput(images=images, nsfw_content_detected=None)


class B200SDXLControlNetPipeline(B200ControlNetPipeline):
    """Drop-in for diffusers' StableDiffusionXLControlNetPipeline as controller/agent/sdxl_controlnet_agent.py:66-75 calls
    it.  Same device chain as the SD-Turbo pipeline on the SDXL topology (UNetConfig.sdxl(): three levels, 1 / 2 / 10
    transformer blocks, 2048-wide context) plus what SDXL adds [upstream, from memory: diffusers 0.29.0
    pipelines/controlnet/pipeline_controlnet_sd_xl.py]:
      * two text encoders: prompt_embeds = cat(hidden_states[-2] of CLIP-L, hidden_states[-2] of OpenCLIP-bigG),
        pooled_prompt_embeds = text_encoder_2's projected EOT embedding;
      * added conditioning of U-Net and ControlNet: emb += add_embedding(cat(pooled, sinusoid(time_ids))),
        time_ids = original_size + crops_coords_top_left + target_size (hoisted: constant per call);
      * sdxl-turbo's EulerAncestralDiscreteScheduler (per-step noise drawn from `generator`, gn_euler_ancestral_step).
    The VAE runs in fp16 on the tensor cores.  Upstream upcasts an fp16 SDXL VAE to fp32 when its config.json sets
    force_upcast (stabilityai/sdxl-turbo's does; madebyollin/sdxl-vae-fp16-fix, which the reference trains with,
    diffusion/README.md:68, does not need it): checkpoint.load_sd_turbo carries the flag into VAEConfig.force_upcast, the
    constructor warns, and every decode is then checked for non-finite values (FloatingPointError instead of a silent
    black image)."""

    def __init__(self, ops: Ops, unet_sd, controlnet_sd, vae_sd, text_sd=None, text2_sd=None,
                 unet_cfg: UNetConfig = UNetConfig.sdxl(), vae_cfg: VAEConfig = VAEConfig(scaling_factor=0.13025),
                 text_cfg: CLIPTextConfig = CLIPTextConfig.sdxl_clip_l(),
                 text2_cfg: CLIPTextConfig = CLIPTextConfig.sdxl_open_clip_bigg(),
                 scheduler_cfg: SchedulerConfig = SchedulerConfig(class_name="EulerAncestralDiscreteScheduler"),
                 tokenizer=None, tokenizer_2=None, use_cuda_graph: bool = False, concurrent_controlnet: bool = True):
        if not unet_cfg.addition_embed:
            raise ValueError("the SDXL pipeline needs a U-Net with text_time added conditioning (UNetConfig.sdxl())")
        super().__init__(ops, unet_sd, controlnet_sd, vae_sd, text_sd, unet_cfg, vae_cfg, text_cfg, scheduler_cfg,
                         tokenizer=tokenizer, use_cuda_graph=use_cuda_graph, concurrent_controlnet=concurrent_controlnet)
        if getattr(vae_cfg, "force_upcast", False):
            import warnings

            warnings.warn("this SDXL snapshot's VAE sets force_upcast=true (diffusers would decode in fp32); decoding in "
                          "fp16 with a non-finite check — prefer sdxl-vae-fp16-fix or taesdxl", RuntimeWarning)
        self.text2_cfg = text2_cfg
        self.text2_impl = DeviceCLIPText(ops, text2_sd, text2_cfg) if text2_sd is not None else None
        self.text_encoder_2 = _ModuleShim(self.text2_impl)
        self.tokenizer_2 = tokenizer_2 or tokenizer
        self._pooled_cache: Dict = {}

    def encode_prompt_sdxl(self, prompt=None, prompt_2=None, prompt_embeds=None, pooled_prompt_embeds=None):
        """-> (prompt_embeds [B, 77, D1 + D2] fp16, pooled [B, P] fp16) on the device.  `prompt` / `prompt_2`: strings
        (tokenizer / tokenizer_2) or token ids [B, 77] for the respective encoder; prompt_2 defaults to prompt."""
        if prompt_embeds is not None:
            if pooled_prompt_embeds is None:
                raise ValueError("If `prompt_embeds` are provided, `pooled_prompt_embeds` also have to be passed.")
            ctx = self.encode_prompt(None, prompt_embeds)
            key = ("pooled",) + tensor_key(pooled_prompt_embeds)
            hit = self._pooled_cache.get(key)
            if hit is None:
                if len(self._pooled_cache) > 64:
                    self._pooled_cache.clear()
                hit = (pooled_prompt_embeds.to(self.ops.device, torch.float16).contiguous(), pooled_prompt_embeds)
                self._pooled_cache[key] = hit
            return ctx, hit[0]
        if self.text_impl is None or self.text2_impl is None:
            raise RuntimeError("pipeline was built without text encoders: pass prompt_embeds and pooled_prompt_embeds")
        ids1 = self._prompt_ids(prompt)
        tok = self.tokenizer
        self.tokenizer = self.tokenizer_2
        try:
            ids2 = self._prompt_ids(prompt if prompt_2 is None else prompt_2)
        finally:
            self.tokenizer = tok
        key = hashlib.sha1(ids1.cpu().numpy().tobytes() + b"|" + ids2.cpu().numpy().tobytes()).hexdigest()
        if key not in self._ctx_cache:
            if len(self._ctx_cache) > 64:
                self._ctx_cache.clear()
            h1, _ = self.text_impl(ids1.to(self.ops.device), penultimate=True)
            h2, pooled = self.text2_impl(ids2.to(self.ops.device), penultimate=True)
            self._ctx_cache[key] = (torch.cat([h1, h2], dim=-1).contiguous(), pooled.to(torch.float16))
        return self._ctx_cache[key]

    @torch.no_grad()
    def __call__(self, prompt=None, prompt_2=None, image=None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, denoising_end: Optional[float] = None,
                 guidance_scale: float = 5.0, negative_prompt=None, negative_prompt_2=None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None,
                 latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None,
                 pooled_prompt_embeds: Optional[torch.Tensor] = None,
                 negative_pooled_prompt_embeds: Optional[torch.Tensor] = None, ip_adapter_image=None,
                 ip_adapter_image_embeds=None, output_type: Optional[str] = "pil", return_dict: bool = True,
                 cross_attention_kwargs=None, controlnet_conditioning_scale: Union[float, List[float]] = 1.0,
                 guess_mode: bool = False, control_guidance_start: Union[float, List[float]] = 0.0,
                 control_guidance_end: Union[float, List[float]] = 1.0, original_size=None,
                 crops_coords_top_left=(0, 0), target_size=None, negative_original_size=None,
                 negative_crops_coords_top_left=(0, 0), negative_target_size=None, clip_skip: Optional[int] = None,
                 callback_on_step_end=None, callback_on_step_end_tensor_inputs: List[str] = ["latents"], **kwargs):
        if guidance_scale is not None and guidance_scale > 1.0:
            raise NotImplementedError("classifier-free guidance (guidance_scale > 1) is not implemented; Genima runs 0.0")
        if num_images_per_prompt not in (None, 1):
            raise NotImplementedError("num_images_per_prompt != 1 is not implemented")
        if guess_mode:
            raise NotImplementedError("guess_mode is not implemented")
        if denoising_end is not None:
            raise NotImplementedError("denoising_end (base / refiner split) is not implemented")
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            raise NotImplementedError("ip-adapter inputs are not implemented")
        if control_guidance_start != 0.0 or control_guidance_end != 1.0:
            raise NotImplementedError("control_guidance_start/end windows are not implemented")
        if isinstance(controlnet_conditioning_scale, (list, tuple)):
            raise NotImplementedError("multi-ControlNet conditioning scales are not implemented")
        if clip_skip is not None:
            raise NotImplementedError("clip_skip is not implemented")
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs are not implemented")
        if callback_on_step_end is not None:
            raise NotImplementedError("per-step callbacks would force a host sync inside the loop; not implemented")
        if kwargs:
            raise TypeError(f"unexpected arguments: {sorted(kwargs)}")
        if output_type not in ("pil", "np", "pt", "latent", "u8"):
            raise ValueError(f"output_type {output_type!r} is not supported")
        if image is None:
            raise ValueError("`image` (the ControlNet conditioning image) is required")
        # the eval loop passes `frame_stack` copies of one prompt string, one per tiled image (eval_genima.py:176-183):
        # identical rows collapse to one conditioning row that is broadcast over the batch of control images
        if isinstance(prompt, (list, tuple)) and len(prompt) > 1 and all(isinstance(p, str) for p in prompt):
            if len(set(prompt)) == 1 and (prompt_2 is None or isinstance(prompt_2, str) or len(set(prompt_2)) == 1):
                prompt = [prompt[0]]
                if isinstance(prompt_2, (list, tuple)):
                    prompt_2 = [prompt_2[0]]
        ctx, pooled = self.encode_prompt_sdxl(prompt, prompt_2, prompt_embeds, pooled_prompt_embeds)
        if pooled.shape[0] != 1:
            if bool((pooled == pooled[:1]).all()) and bool((ctx == ctx[:1]).all()):
                ctx, pooled = ctx[:1].contiguous(), pooled[:1].contiguous()
            else:
                raise NotImplementedError("different prompts within one call are not implemented for SDXL (the added "
                                          "text_time conditioning is hoisted per call): pass one prompt, or identical rows")
        cond_u8 = self._control_image_u8(image)
        H, W = int(cond_u8.shape[1]), int(cond_u8.shape[2])
        # _get_add_time_ids: original_size + crops_coords_top_left + target_size, each (height, width)
        osz = tuple(original_size) if original_size is not None else (H, W)
        tsz = tuple(target_size) if target_size is not None else (height or H, width or W)
        self._added = dict(text_embeds=pooled, time_ids=[float(v) for v in (*osz, *crops_coords_top_left, *tsz)])
        return self._generate(None, cond_u8, height, width, num_inference_steps, generator, latents, ctx,
                              output_type, return_dict, float(controlnet_conditioning_scale))


class B200Pix2PixPipeline(B200ControlNetPipeline):
    """Drop-in for diffusers' StableDiffusionInstructPix2PixPipeline as controller/agent/sd_pix2pix_agent.py:52-60 calls
    it (same six keyword arguments as the ControlNet agent).  No ControlNet: the input image is VAE-encoded once per
    call (posterior mode, not scaled) and its 4 latent channels ride in channels 4..7 of the U-Net's 8-channel conv_in
    input at every step (`torch.cat([scaled_latents, image_latents], dim=1)` upstream).  With the reference's
    guidance_scale 0.0 `do_classifier_free_guidance` is False upstream, so there is one U-Net evaluation per step;
    guidance (guidance_scale > 1 and image_guidance_scale >= 1: the three-way batch) raises NotImplementedError."""

    def __init__(self, ops: Ops, unet_sd, vae_sd, text_sd=None, unet_cfg: UNetConfig = UNetConfig(in_channels=8),
                 vae_cfg: VAEConfig = VAEConfig(), text_cfg: CLIPTextConfig = CLIPTextConfig.sd_turbo(),
                 scheduler_cfg: SchedulerConfig = SchedulerConfig(),
                 tokenizer: Optional[Callable[[Sequence[str]], torch.Tensor]] = None, use_cuda_graph: bool = False):
        if isinstance(vae_cfg, TAESDConfig):
            raise NotImplementedError("the InstructPix2Pix pipeline needs the AutoencoderKL encoder; TAESD is decode-only here")
        if unet_cfg.in_channels != 2 * vae_cfg.latent_channels or 2 * vae_cfg.latent_channels != LATENT_CPAD:
            raise ValueError(f"InstructPix2Pix U-Net must take {2 * vae_cfg.latent_channels} input channels "
                             f"(latents + image latents), config says {unet_cfg.in_channels}")
        self.ops = ops
        self.unet_cfg, self.vae_cfg, self.text_cfg = unet_cfg, vae_cfg, text_cfg
        self.concurrent_controlnet = False
        self.ops_side = self.ops_zero = ops
        self.side_stream = self.zero_stream = None
        self.overlap_zero_convs = False
        self.fuse_scheduler = False          # the image latents ride in channels 4..7 of the scaled input: not foldable
        self.unet_impl = DeviceUNet(ops, unet_sd, unet_cfg)
        self.controlnet_impl = None
        self.vae_impl = DeviceVAEDecoder(ops, vae_sd, vae_cfg)
        self.vae_enc_impl = DeviceVAEEncoder(ops, vae_sd, vae_cfg)
        self.text_impl = DeviceCLIPText(ops, text_sd, text_cfg) if text_sd is not None else None
        self.schedule = EulerDiscreteSchedule(scheduler_cfg)
        self.tokenizer = tokenizer
        self.use_cuda_graph = use_cuda_graph
        self.vae_scale_factor = 2 ** (len(vae_cfg.block_out_channels) - 1)
        self.vae = _ModuleShim(self.vae_impl)
        self.unet = _ModuleShim(self.unet_impl)
        self.text_encoder = _ModuleShim(self.text_impl)
        self.scheduler = self.schedule
        self._ctx_cache, self._kv_cache, self._temb_cache, self._graphs, self._pinned = {}, {}, {}, {}, {}
        self._tuned_shapes = set()
        self.progress_bar_disabled = True
        self._added = None

    def _context_kv(self, ctx: torch.Tensor) -> Dict[str, torch.Tensor]:
        key = tensor_key(ctx)
        hit = self._kv_cache.get(key)
        if hit is None or hit[1] is not ctx:
            if len(self._kv_cache) > 8:
                self._kv_cache.clear()
            kv = {"u:" + tr.prefix: tr.project_context(self.ops, ctx) for tr in self.unet_impl.transformers()}
            kv["__ctx__"] = ctx
            hit = self._kv_cache[key] = (kv, ctx)
        return hit[0]

    def _time_rows(self, n_steps: int, batch: int):
        key = (n_steps, batch)
        if key not in self._temb_cache:
            ts, _ = self.schedule.set_timesteps(n_steps)
            self._temb_cache[key] = [self.unet_impl.temb_rows(self.unet_impl.resblocks(),
                                                              self.unet_impl.time_embedding(float(t)), batch) for t in ts]
        return self._temb_cache[key]

    def _denoise_and_decode(self, cond_u8: torch.Tensor, lat_in: torch.Tensor, kv, tk: int, n_steps: int,
                            want_image: bool, cond_scale: float, noise: Optional[torch.Tensor] = None):
        """cond_u8 [B, H, W, 3] (the image to edit); lat_in [B, h, w, 8] fp16 (unit-variance noise in channels 0..3)."""
        ops = self.ops
        temb = self._time_rows(n_steps, lat_in.shape[0])
        ops.gn_stats_reset()
        sig = self.schedule.sigmas
        kv_u = {k[2:]: v for k, v in kv.items() if k.startswith("u:")}
        lc = self.vae_cfg.latent_channels
        img_lat = self.vae_enc_impl.encode(cond_u8)                                  # [B, h, w, 8], mean in 0..3
        x = ops.scale(lat_in, self.schedule.init_noise_sigma)
        xs = ops.scale(x, 1.0 / float(np.sqrt(float(sig[0]) ** 2 + 1.0)))
        eps = torch.zeros_like(x)
        for i in range(n_steps):
            xs[..., lc:] = img_lat[..., :lc]                                         # cat([scaled latents, image latents])
            mid, skips = self.unet_impl.encode(xs, temb[i], kv_u, tk)
            self.unet_impl.decode(mid, skips, temb[i], kv_u, tk, eps)
            x, xs = self._scheduler_step(x, eps, i, noise)
        img = None
        if want_image:
            img = self.vae_impl.decode(ops.scale(x, 1.0 / self.vae_cfg.scaling_factor))
        return x, img

    @torch.no_grad()
    def __call__(self, prompt=None, image=None, num_inference_steps: int = 100, guidance_scale: float = 7.5,
                 image_guidance_scale: float = 1.5, negative_prompt=None, num_images_per_prompt: Optional[int] = 1,
                 eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 ip_adapter_image=None, ip_adapter_image_embeds=None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback_on_step_end=None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], cross_attention_kwargs=None, **kwargs):
        # upstream: do_classifier_free_guidance = guidance_scale > 1.0 and image_guidance_scale >= 1.0
        if guidance_scale is not None and guidance_scale > 1.0 and image_guidance_scale >= 1.0:
            raise NotImplementedError("text / image classifier-free guidance (three-way batch) is not implemented; "
                                      "Genima runs guidance_scale 0.0")
        if num_images_per_prompt not in (None, 1):
            raise NotImplementedError("num_images_per_prompt != 1 is not implemented")
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            raise NotImplementedError("ip-adapter inputs are not implemented")
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs are not implemented")
        if callback_on_step_end is not None:
            raise NotImplementedError("per-step callbacks would force a host sync inside the loop; not implemented")
        if kwargs:
            raise TypeError(f"unexpected arguments: {sorted(kwargs)}")
        if output_type not in ("pil", "np", "pt", "latent", "u8"):
            raise ValueError(f"output_type {output_type!r} is not supported")
        if image is None:
            raise ValueError("`image` cannot be undefined.")
        return self._generate(prompt, image, None, None, num_inference_steps, generator, latents, prompt_embeds,
                              output_type, return_dict, 1.0)
