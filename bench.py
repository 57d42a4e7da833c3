#!/usr/bin/env python
"""Genima agent-step benchmark (BASELINE.json: "Genima agent steps/sec (5 denoise, 4x256^2 views)").

    python bench.py --gpus N --steps K --warmup W                      own arm (libgenima_b200.so on B200)
    python bench.py --impl reference --gpus N --steps K --warmup W     CPU arm: the fp32 oracle on the host cores
    torchrun ... bench.py --gpus N ...                                 one rank per GPU (N > 1), episode-parallel

One "step" = one agent step of controller/eval_genima.py:162-275 restricted to the hot path: tile 4x256^2 views ->
ControlNet + SD-Turbo U-Net, 5 Euler-trailing denoise steps on one 512^2 tile (latents 1x4x64x64) -> KL-VAE decode ->
untile -> ACT controller -> a_hat [1, 20, 8]  (BASELINE.json configs[2]).  Synthetic seeded
weights of the real architectures (no checkpoints offline), synthetic observations.
  value   device-resident throughput: inputs already in HBM, the whole step is one CUDA-graph replay, K steps timed
          with CUDA events between barriers, max over ranks; whole-job = N x per-rank (weak scaling: every rank runs its
          own episodes, no per-step collective — SURVEY.md §8e).
  e2e     the same step through the reference-facing plugin API with HOST buffers, exactly as the reference loop calls
          it: PIL views -> tile_images -> agent.infer(...)[0] (PIL) -> untile_images -> obs tensors .to(device) ->
          controller.act(obs) -> .cpu().numpy(); H2D / D2H copies and PIL conversions are inside the timed region.
L2: every step streams ~2.7 GB of weights (>> 126 MB L2), so no explicit flush is needed between timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from genima_b200 import weights as W  # noqa: E402
from genima_b200.configs import ACTConfig, CLIPTextConfig, UNetConfig, VAEConfig  # noqa: E402

METRIC = "genima_agent_steps_per_sec"
UNIT = "steps/s"
# algorithmic FLOPs (2 x MAC, dense) per agent step at 5 denoise steps, B = 1 tile
FLOPS_DENOISE_ITER = 1.088e12
FLOPS_VAE = 2.515e12
FLOPS_ACT = 0.0228e12


def presets(name: str, autoencoder: str = ""):
    """autoencoder containing 'taesd' selects AutoencoderTiny, like eval_cfg.autoencoder in the reference
    (controller/agent/sd_controlnet_agent.py:45-49); the default ('') is the KL-VAE of BASELINE configs[2]."""
    from genima_b200.configs import TAESDConfig

    taesd = "taesd" in (autoencoder or "")
    if name == "tiny":
        return UNetConfig.tiny(), (TAESDConfig.tiny() if taesd else VAEConfig.tiny()), ACTConfig.tiny()
    return UNetConfig(), (TAESDConfig() if taesd else VAEConfig()), ACTConfig()


def model_shapes(ucfg, vcfg, acfg):
    vae = W.taesd_decoder_shapes(vcfg) if hasattr(vcfg, "num_blocks") else W.vae_decoder_shapes(vcfg)
    return OrderedDict(unet=W.unet_shapes(ucfg), controlnet=W.controlnet_shapes(ucfg), vae=vae, act=W.act_shapes(acfg))


def synth_all(shapes):
    salts = dict(unet=0, controlnet=1, vae=2, act=3)
    return OrderedDict((m, W.synth_state_dict(s, salt=salts[m])) for m, s in shapes.items())


def make_inputs(ucfg, acfg, seed=0):
    """SURVEY.md §8d config 3 inputs: views randint seed 0, qpos randn seed 1, task_emb randn seed 4, prompt embeddings
    randn seed 3 (the text encoders are cached per episode, so their output is an input of the step)."""
    S = acfg.image_size
    views = torch.randint(0, 256, (4, 3, S, S), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
    qpos = torch.randn(1, acfg.state_dim, generator=torch.Generator().manual_seed(1))
    task = torch.randn(1, acfg.task_emb_dim, generator=torch.Generator().manual_seed(4))
    ctx = torch.randn(1, 77, ucfg.cross_attention_dim, generator=torch.Generator().manual_seed(3)).half()
    lat = torch.randn(1, 4, S // 4, S // 4, generator=torch.Generator().manual_seed(2))
    return views, qpos, task, ctx, lat


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [ln for (ts, ln) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for _, ln in self.lines]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_agent_step_fn(ucfg, vcfg, acfg, sds, n_denoise):
    """Returns a closure running one full agent step with the fp32 CPU oracle (weights pre-converted to fp32 once)."""
    from oracle.pipeline import agent_step

    w32 = {m: {k: v.float() for k, v in sd.items()} for m, sd in sds.items()}
    views, qpos, task, ctx, lat = make_inputs(ucfg, acfg)
    views_hwc = views.permute(0, 2, 3, 1).contiguous().numpy()

    def step():
        with torch.no_grad():
            return agent_step(w32, ucfg, vcfg, acfg, views_hwc, ctx.float(), lat, qpos, task, n_denoise)

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ucfg, vcfg, acfg = presets(args.preset, args.autoencoder)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sds = synth_all(model_shapes(ucfg, vcfg, acfg))
    step = cpu_agent_step_fn(ucfg, vcfg, acfg, sds, args.denoise_steps)
    budget = float(args.cpu_budget_s)
    t0 = time.perf_counter()
    step()                                            # warm-up (also sizes the run)
    t_one = time.perf_counter() - t0
    warm = 1
    while warm < args.warmup and (warm + 1) * t_one < 0.25 * budget:
        step()
        warm += 1
    k = max(1, min(args.steps, int((budget - warm * t_one) / max(t_one, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = time.perf_counter() - t0
    value = k / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * dt / k, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, ucfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{k} full agent steps (time-budgeted to {budget:.0f} s; {args.steps} requested) of "
                                   "the fp32 PyTorch-CPU oracle (oracle/pipeline.py::agent_step): the reference's "
                                   "diffusers/RoboBase stack is not installable offline and has no CPU fp16 path "
                                   "(SURVEY.md F7, F10)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, ucfg):
    s = ucfg.sample_size * 8
    return {"workload": f"full Genima agent step (BASELINE configs[2]): 4x{s // 2}^2 views -> {s}^2 tile -> "
                        f"ControlNet+SD-Turbo U-Net x{args.denoise_steps} Euler-trailing steps -> KL-VAE decode -> "
                        "untile -> ACT (ResNet18-FiLM x4 + 4enc/6dec transformer) -> a_hat[1,20,8]",
            "preset": args.preset, "denoise_steps": args.denoise_steps, "tile_batch": 1, "guidance_scale": 0.0,
            "autoencoder": ("taesd (AutoencoderTiny)" if "taesd" in (args.autoencoder or "") else "AutoencoderKL"),
            "weights": "synthetic seeded (real SD-2.1/SD-Turbo + ACT topologies)", "parallelism": f"episode-dp{args.gpus}",
            "l2": "inputs larger than L2: ~2.7 GB of fp16 weights streamed per step vs 126 MB L2"}


# ------------------------------------------------------------------------------------------------ own arm
def stub_tokenizer(prompts):
    """Deterministic stand-in for the CLIP BPE tokenizer (no vocabulary exists offline): BOS 49406, one id
    per whitespace-separated word (crc32 -> [1000, 41000)), EOS 49407, zero padding to 77 (the SD-2.x tokenizer pads
    with id 0).  Lets the e2e leg hand the pipeline the STRING prompts the reference loop builds
    (controller/eval_genima.py:178), so tokenisation, the prompt-cache lookup and its host-side hashing are timed."""
    import zlib

    ids = torch.zeros(len(prompts), 77, dtype=torch.int64)
    for r, p in enumerate(prompts):
        toks = [49406] + [1000 + zlib.crc32(w.encode()) % 40000 for w in p.split()][:75] + [49407]
        ids[r, :len(toks)] = torch.tensor(toks)
    return ids


def graph_kernel_times(fn, reps: int = 3, groups=("gemm_tc_kernel", "attn_tc_kernel")):
    """Per-kernel device intervals of `fn()` (a CUDA-graph replay) from CUPTI through torch.profiler: returns
    {kernel name: dict(n=launches per call, sum_ms=sum of durations per call, busy_ms=length of the UNION of the kernel's
    intervals per call)} plus "__span_ms__".  The union is the time during which at least one instance of the kernel is
    executing: it can never exceed the step time, whatever overlaps on the other streams."""
    from torch.profiler import ProfilerActivity, profile

    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    by = {}
    t_lo, t_hi = None, None
    for ev in prof.events():
        if ev.device_type is None or "cuda" not in str(ev.device_type).lower():
            continue
        name = ev.name.split("(")[0].replace("void ", "").strip()
        tr = ev.time_range
        by.setdefault(name, []).append((tr.start, tr.end))
        t_lo = tr.start if t_lo is None else min(t_lo, tr.start)
        t_hi = tr.end if t_hi is None else max(t_hi, tr.end)
    def union(iv):
        iv = sorted(iv)
        busy, cur_s, cur_e = 0.0, None, None
        for a, b in iv:
            if cur_e is None or a > cur_e:
                if cur_e is not None:
                    busy += cur_e - cur_s
                cur_s, cur_e = a, b
            else:
                cur_e = max(cur_e, b)
        if cur_e is not None:
            busy += cur_e - cur_s
        return busy

    out = {}
    for name, iv in by.items():
        out[name] = {"n": len(iv) / reps, "sum_ms": sum(b - a for a, b in iv) / reps / 1e3,
                     "busy_ms": union(iv) / reps / 1e3}
    for grp in groups:          # all template flavours of one kernel together
        iv = [x for name, v in by.items() if grp in name for x in v]
        out["__group__" + grp] = {"n": len(iv) / reps, "sum_ms": sum(b - a for a, b in iv) / reps / 1e3,
                                  "busy_ms": union(iv) / reps / 1e3}
    out["__span_ms__"] = (t_hi - t_lo) / reps / 1e3 if t_lo is not None else None
    return out


def gpu_baseline_leg(args):
    """Stock PyTorch fp16 on the same GPU (what the reference's diffusers / RoboBase stack would run here if it were
    installable: cuDNN / cuBLAS / SDPA library kernels on the oracle's restatement of the graph), in a subprocess:
    eager and CUDA-graphed always; torch.compile in the reference's own mode (`reduce-overhead`,
    controller/agent/sd_controlnet_agent.py:52-62) and `max-autotune` when asked for (--gpu-baseline all) or from the
    committed capture of the same script (profiles/) otherwise.  A BASELINE leg like cpu_baseline: never the product."""
    modes = {"eager": "eager,graph", "all": "eager,graph,compile-reduce-overhead,compile-max-autotune"}[args.gpu_baseline]
    res = {}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "stock_torch_gpu_baseline.py"), "--modes", modes,
                            "--denoise-steps", str(args.denoise_steps)], capture_output=True, text=True,
                           timeout=3000 if args.gpu_baseline == "all" else 240)
        last = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        res = json.loads(last[-1]) if last else {"error": (r.stderr or "no output")[-400:]}
    except Exception as ex:  # noqa: BLE001
        res = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    if args.gpu_baseline != "all":
        try:
            with open(os.path.join(ROOT, "profiles", "r2_gpu_baseline_compile.json")) as f:
                cap = json.load(f)
            res["committed_capture"] = {k: cap.get(k) for k in ("compile_reduce_overhead_ms", "compile_max_autotune_ms",
                                                                "compile_reduce_overhead_error",
                                                                "compile_max_autotune_error", "torch", "gpu")}
            res["committed_capture"]["source"] = ("profiles/r2_gpu_baseline_compile.json: the same script run once with "
                                                  "--modes all on a B200 of this pool (compilation takes minutes)")
        except Exception:
            pass
    return res


def build_world(args, rank: int = 0, local: int = 0):
    """Weights (rank 0 synthesises, one arena broadcast), the pipeline, the hydra-style controller, the fused step and
    the inputs of the workload: everything run_ours and the tools/ scripts share."""
    from types import SimpleNamespace

    from genima_b200 import distributed as gd
    from genima_b200.agents import B200GenimaACT
    from genima_b200.ops import Ops
    from genima_b200.pipeline import B200ControlNetPipeline
    from genima_b200.step import GenimaStep

    dev = torch.device("cuda", local)
    ucfg, vcfg, acfg = presets(args.preset, args.autoencoder)
    tiny = args.preset == "tiny"
    tcfg = CLIPTextConfig.tiny() if tiny else CLIPTextConfig.sd_turbo()
    ccfg = CLIPTextConfig.tiny(projection_dim=acfg.task_emb_dim) if tiny else CLIPTextConfig.vit_b32()
    shapes = model_shapes(ucfg, vcfg, acfg)
    shapes["text"] = W.clip_text_shapes(tcfg)       # SD text encoder (prompt -> context, cached per episode)
    shapes["clip"] = W.clip_text_shapes(ccfg)       # controller's CLIP ViT-B/32 text tower (task embedding, cached)
    # ---- weights: rank 0 synthesises, ONE broadcast of the packed arena, every rank binds views into its copy
    t_w = time.perf_counter()
    sds_host = None
    if rank == 0:
        sds_host = synth_all(model_shapes(ucfg, vcfg, acfg))
        sds_host["text"] = W.synth_state_dict(shapes["text"], salt=6)
        sds_host["clip"] = W.synth_state_dict(shapes["clip"], salt=7)
    synth_s = time.perf_counter() - t_w
    timings = {}
    sds, arena = gd.broadcast_weights(shapes, sds_host, src=0, device=dev, timings=timings)

    ops = Ops(local)
    pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], sds["text"], ucfg, vcfg, tcfg,
                                  tokenizer=stub_tokenizer, use_cuda_graph=True)
    # plugin point 2, built the way GenimaEvalWorkspace builds it (eval_genima.py:55-66, 91-103)
    S = acfg.image_size
    cameras = ["wrist", "front", "right_shoulder", "left_shoulder"]
    controller = B200GenimaACT(device=dev, observation_space=None, action_space=None, act_cfg=acfg, ops=ops,
                               clip_state_dict=sds["clip"], clip_cfg=ccfg, num_train_envs=1)
    controller.train(False)
    controller.load_state_dict({f"actor.{k}": v for k, v in sds["act"].items()}, strict=False)
    act = controller.actor.impl                                    # the DeviceACT both legs share
    step = GenimaStep(pipe, act, num_inference_steps=args.denoise_steps, use_cuda_graph=True)

    views, qpos, _task, _ctx, lat = make_inputs(ucfg, acfg, seed=rank)
    goal = "open the box"
    prompts = [f"tiled perspectives of a robot arm executing '{goal}'"]                  # eval_genima.py:178
    negative_prompts = ["monochrome, lowres, bad anatomy, worst quality, low quality"]   # eval_genima.py:181-183
    lang_np = stub_tokenizer([goal]).numpy().astype(np.int32)[:, None, :]                # obs["lang_tokens"] [T=1,77]->[1,1,77]
    d_ctx = pipe.encode_prompt(prompts)                                                  # text encoder: once per episode
    d_task = controller.encode_clip_text(torch.from_numpy(lang_np).to(dev))[0]           # CLIP text tower: once per episode
    d_views = views.permute(0, 2, 3, 1).contiguous()[None].to(dev)          # [1, 4, S, S, 3] u8
    d_lat, d_qpos = lat.to(dev), qpos.to(dev)

    return SimpleNamespace(**{k: v for k, v in locals().items() if k not in ("args", "SimpleNamespace")})


def make_e2e_step(pipe, ops, controller, views, qpos, prompts, negative_prompts, lang_np, acfg, denoise_steps, dev, local,
                  tick=None):
    """-> (e2e_step, bytes): `e2e_step()` is one iteration of the reference loop body (controller/eval_genima.py:163-249)
    on host buffers; `bytes` = dict(h2d=, d2h=) filled by the first call.  tick(name, t0): optional phase timer hook."""
    from PIL import Image

    from genima_b200.agents import B200ControlNetAgent
    from genima_b200.host_glue import tile_images, untile_images

    S = acfg.image_size
    cameras = ["wrist", "front", "right_shoulder", "left_shoulder"]
    agent = B200ControlNetAgent.__new__(B200ControlNetAgent)       # bind the already-built pipeline (no second copy)
    agent.eval_cfg = dict(image_resolution=2 * S, device=f"cuda:{local}")
    agent.pipe, agent._ops = pipe, ops
    agent.set_optimizations()
    agent.common_setup()
    obs_np = {f"{c}_rgb": views[i:i + 1].numpy() for i, c in enumerate(cameras)}             # [T=1, 3, S, S] u8
    low_dim = qpos.numpy()                                                                   # [T=1, 8]
    gen = [torch.Generator(device=dev).manual_seed(2)]                                       # eval_genima.py:129-135
    nbytes = {"h2d": 0, "d2h": 0}
    tick = tick or (lambda name, t0: None)

    def e2e_step():
        t0 = time.perf_counter()
        # eval_genima.py:163-186: env observation (uint8 CHW numpy) -> PIL -> tile
        rgbs = [Image.fromarray(np.transpose(obs_np[f"{c}_rgb"][0], (1, 2, 0))) for c in cameras]
        tiles = tile_images(rgbs, 1) if S == 256 else [Image.fromarray(
            np.concatenate([np.concatenate([np.asarray(rgbs[0]), np.asarray(rgbs[1])], 1),
                            np.concatenate([np.asarray(rgbs[2]), np.asarray(rgbs[3])], 1)], 0))]
        tick("obs -> PIL -> tile_images", t0)
        t0 = time.perf_counter()
        # :199-210 (string prompts -> tokenizer -> prompt cache; CUDA generator shared across the batch)
        target = agent.infer(images=tiles, prompts=prompts, negative_prompts=negative_prompts,
                             num_inference_steps=denoise_steps, guidance_scale=0.0, generator=gen * len(tiles))
        tick("diffusion_agent.infer (incl. GPU)", t0)
        t0 = time.perf_counter()
        # :224-234
        if S == 256:
            un = untile_images(target[0], cameras, agent.transform_to_half_resolution)
        else:
            g = np.asarray(target[0][0])
            quads = [g[:S, :S], g[:S, S:], g[S:, :S], g[S:, S:]]
            un = {c: np.ascontiguousarray(np.transpose(q, (2, 0, 1))[None]) for c, q in zip(cameras, quads)}
        obs = {f"{c}_rgb": un[c] for c in cameras}
        obs["low_dim_state"] = low_dim
        obs["lang_tokens"] = lang_np[0]
        tick("untile_images", t0)
        t0 = time.perf_counter()
        # :237-249
        obs = {k: torch.from_numpy(v).to(dev).unsqueeze(0) for k, v in obs.items()}
        tick("obs -> device", t0)
        t0 = time.perf_counter()
        actions = controller.act(obs, step=0, eval_mode=True)[0]
        actions = actions.detach().cpu().numpy()
        tick("controller_agent.act (incl. GPU)", t0)
        nbytes["h2d"] = (tiles[0].size[0] * tiles[0].size[1] * 3 + sum(v.nbytes for v in un.values()) + low_dim.nbytes
                         + lang_np[0].nbytes)
        nbytes["d2h"] = tiles[0].size[0] * tiles[0].size[1] * 3 + actions.nbytes + lang_np[0].nbytes  # (ids read back: cache key)
        return actions

    return e2e_step, nbytes


def run_ours(args):
    import torch.distributed as dist

    from genima_b200 import distributed as gd
    from genima_b200.agents import B200GenimaACT
    from genima_b200.ops import Ops
    from genima_b200.pipeline import B200ControlNetPipeline
    from genima_b200.step import GenimaStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a B200: genima_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm_init_s = 0.0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own INIT lines (communicator size, transports) are left visible: they prove how many ranks joined.
        # They go to stdout BEFORE the JSON line, which is always the LAST line this program prints.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):   # (the image presets VERSION)
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            # ... on stderr: NCCL also logs at teardown, and the JSON line must stay the last line of stdout
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        t_c = time.perf_counter()
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.ones(1, device=dev)
        dist.all_reduce(warm)                       # communicator bring-up (rings / NVLS set-up) happens here, once
        torch.cuda.synchronize()
        comm_init_s = time.perf_counter() - t_c
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    w = build_world(args, rank, local)
    ucfg, vcfg, acfg, sds, arena, ops, pipe, controller, act, step = (w.ucfg, w.vcfg, w.acfg, w.sds, w.arena, w.ops,
                                                                      w.pipe, w.controller, w.act, w.step)
    views, qpos, lat, prompts, negative_prompts, lang_np = w.views, w.qpos, w.lat, w.prompts, w.negative_prompts, w.lang_np
    d_ctx, d_task, d_views, d_lat, d_qpos, synth_s, timings = (w.d_ctx, w.d_task, w.d_views, w.d_lat, w.d_qpos, w.synth_s,
                                                                w.timings)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    # ---- tile configurations: rank 0 measures them on its first (eager) pass, every other rank adopts them, so all
    # ranks launch identical kernels (bit-identical results wherever an episode is sharded)
    tune_bytes = 0
    if world > 1:
        if rank == 0:
            step(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx)
        barrier()
        tune_bytes = gd.sync_tune_caches(pipe.all_ops(), src=0)

    # ---- device-resident timed region
    for _ in range(max(args.warmup, 3)):
        out = step(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx)
    e1.record()
    barrier()
    t1 = time.time()
    ms = gd.reduce_max(e0.elapsed_time(e1), device=dev)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    launches = step.launches_per_step * args.steps
    value = world * args.steps / (ms * 1e-3)
    a_hat = out["a_hat"].float().cpu()
    tile_dev = out["tile_u8"].clone()

    # ---- U-Net + ControlNet only (sub-metric "U-Net ms/step"): 5-step loop to latents, graph replay
    for _ in range(3):
        pipe(prompt_embeds=d_ctx, image=tile_dev, num_inference_steps=args.denoise_steps, guidance_scale=0.0,
             latents=d_lat, output_type="latent")
    torch.cuda.synchronize()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nrep = max(5, min(args.steps, 30))
    u0.record()
    for _ in range(nrep):
        pipe(prompt_embeds=d_ctx, image=tile_dev, num_inference_steps=args.denoise_steps, guidance_scale=0.0,
             latents=d_lat, output_type="latent")
    u1.record()
    torch.cuda.synchronize()
    unet_ms = u0.elapsed_time(u1) / nrep / args.denoise_steps

    # ---- throughput variant: B independent episodes batched into one denoise call.
    # Informational only: `value` above is the batch-1 step.  Results are bit-identical across the batch entries.
    batched = None
    if rank == 0 and not args.no_batched:
        batched = []
        for Bt in (2, 4):
            bv = d_views.repeat(Bt, 1, 1, 1, 1)
            bargs = (bv, d_lat.repeat(Bt, 1, 1, 1), d_qpos.repeat(Bt, 1), d_task.repeat(Bt, 1))
            bctx = d_ctx.repeat(Bt, 1, 1)
            for _ in range(3):
                bout = step(*bargs, prompt_embeds=bctx)
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nb = max(5, min(args.steps, 10))
            b0.record()
            for _ in range(nb):
                bout = step(*bargs, prompt_embeds=bctx)
            b1.record()
            torch.cuda.synchronize()
            bms = b0.elapsed_time(b1) / nb
            ba = bout["a_hat"].float()
            batched.append({"tile_batch": Bt, "ms_per_call": bms, "value": Bt * 1e3 / bms, "unit": UNIT,
                            "max_abs_diff_between_batch_entries": float((ba - ba[0:1]).abs().max()),
                            "max_abs_diff_vs_batch_1": float((ba[0].cpu() - a_hat[0]).abs().max())})
        out = step(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx)   # back to the batch-1 graph / buffers

    # ---- e2e through the reference-facing API with host buffers: the loop body of controller/eval_genima.py:163-249
    e2e_step, e2e_bytes = make_e2e_step(pipe, ops, controller, views, qpos, prompts, negative_prompts, lang_np, acfg,
                                        args.denoise_steps, dev, local)
    n_e2e = max(3, min(args.steps, args.e2e_steps))

    with torch.inference_mode():                                                             # eval_genima.py:199
        for _ in range(3):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(n_e2e):
            actions = e2e_step()
        torch.cuda.synchronize()
    e2e_ms = gd.reduce_max((time.perf_counter() - w0) * 1e3, device=dev)
    e2e_value = world * n_e2e / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (gemm_tc_kernel: every convolution and linear layer)
    #   FLOPs / bytes: counted by the library per gn_linear / gn_conv2d call during ONE eager step (gn_profile_*);
    #   time: CUPTI intervals of the kernel's launches inside the REPLAYED graph (torch.profiler), as the length of their
    #   union -- the time during which a gemm_tc_kernel is executing, <= the step time by construction.
    roofline, classes, kernels = None, None, None
    if rank == 0 and not args.no_roofline:
        eager = GenimaStep(pipe, act, num_inference_steps=args.denoise_steps, use_cuda_graph=False)
        eager(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx)
        torch.cuda.synchronize()
        for o in pipe.all_ops():
            o.profile_begin()
        eager(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx)
        for o in pipe.all_ops():           # the ControlNet encoder runs through its own handle / stream
            part = o.profile_end()
            if classes is None:
                classes = part
            else:
                for k, v in part.items():
                    for f in v:
                        classes[k][f] += v[f]
        try:
            kernels = graph_kernel_times(lambda: step(d_views, d_lat, d_qpos, d_task, prompt_embeds=d_ctx))
        except Exception as ex:  # noqa: BLE001
            kernels = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        roofline = make_roofline(classes, kernels, ms / args.steps)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        from oracle.pipeline import agent_step as oracle_agent_step

        w32 = {m: {k: v.float().cpu() for k, v in sds[m].items()} for m in ("unet", "controlnet", "vae", "act")}
        views_hwc = views.permute(0, 2, 3, 1).contiguous().numpy()
        c0 = time.perf_counter()
        with torch.no_grad():
            ref = oracle_agent_step(w32, ucfg, vcfg, acfg, views_hwc, d_ctx.float().cpu(), lat, qpos,
                                    d_task.float().cpu(), args.denoise_steps)
        cdt = time.perf_counter() - c0
        err = float((a_hat - ref["a_hat"]).abs().max() / ref["a_hat"].abs().max())
        dd = np.abs(tile_dev.cpu().numpy()[0].astype(np.int32) - ref["tile_u8"][0].astype(np.int32))
        cpu_baseline = {"value": 1.0 / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "1 full agent step (same inputs, same synthetic weights) of the fp32 PyTorch-CPU "
                                  "oracle, oracle/pipeline.py::agent_step, no warm-up",
                        "a_hat_normalised_max_err_vs_device": err, "tile_max_abs_diff_levels": int(dd.max()),
                        "tile_frac_within_1_level": float((dd <= 1).mean())}
        del w32

    gpu_baseline = None
    if rank == 0 and world == 1 and args.gpu_baseline != "none":
        gpu_baseline = gpu_baseline_leg(args)

    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic", "config": workload_config(args, ucfg),
            "unet_ms_per_step": unet_ms, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_bytes["h2d"]), "d2h_bytes_per_step": int(e2e_bytes["d2h"]),
                    "steps": n_e2e, "ms_per_step": e2e_ms / n_e2e,
                    "api": "the loop body of controller/eval_genima.py:163-249: uint8 CHW observations -> PIL -> tile_images "
                           "-> B200ControlNetAgent.infer(string prompts through a stub tokenizer + prompt cache, CUDA "
                           "generator) -> untile_images -> obs tensors .to(device) (fresh lang_tokens every step) -> "
                           "B200GenimaACT.act -> .cpu().numpy(), under torch.inference_mode()"},
            "gpu_launches": int(launches), "launches_per_step": int(step.launches_per_step),
            "roofline": roofline, "kernel_classes": classes, "cpu_baseline": cpu_baseline, "gpu_baseline": gpu_baseline,
            "batched": batched,
            "weights": {"gb": arena.numel() * 2 / 1e9, "synthesis_s_rank0": synth_s, "pack_h2d_s": timings.get("pack_s"),
                        "broadcast_s": timings.get("broadcast_s"), "comm_init_s": comm_init_s,
                        "broadcast_gb_per_s": (arena.numel() * 2 / 1e9 / timings["broadcast_s"]
                                               if timings.get("broadcast_s") else None),
                        "tune_cache_bytes_broadcast": tune_bytes},
            "agent_step_tflop": (args.denoise_steps * FLOPS_DENOISE_ITER
                                 + (0.14e12 if "taesd" in (args.autoencoder or "") else FLOPS_VAE) + FLOPS_ACT) / 1e12
            if args.preset != "tiny" else None,
        }
        print(json.dumps(line), flush=True)          # ALWAYS the last line on stdout
    return 0


def make_roofline(classes, kernels, step_ms):
    """Dominant kernel = gemm_tc_kernel (gn_linear + gn_conv2d launches: every convolution and linear layer)."""
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step); fp16 uses the same pipe"
    if not peak:
        peak, src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
    g_fl = classes["linear"]["flops"] + classes["conv"]["flops"]
    g_by = classes["linear"]["bytes"] + classes["conv"]["bytes"]
    g_calls = classes["linear"]["calls"] + classes["conv"]["calls"]
    gk = (kernels or {}).get("__group__gemm_tc_kernel")
    busy = sum_ms = n = None
    if gk and gk["n"] > 0:
        sum_ms, n, busy = gk["sum_ms"], gk["n"], gk["busy_ms"]     # union over ALL flavours of the kernel template
    t_ms = busy if busy else step_ms
    achieved = g_fl / (t_ms * 1e-3) / 1e12
    # DRAM bytes the GEMM launches of one step really moved, from the committed ncu capture (default workload only)
    traffic = traffic_step = traffic_src = None
    for cap_name in ("r2_dram_traffic.json", "r1i_dram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", cap_name)) as f:
                cap = json.load(f)
            k = cap["kernels"]["gn::gemm_tc_kernel"]
            traffic = k["dram_bytes_per_launch"]
            traffic_step = k["dram_read_bytes_per_step"] + k["dram_write_bytes_per_step"]
            traffic_src = (f"profiles/{cap_name}: dram__bytes_read.sum + dram__bytes_write.sum of the {k['launches']} "
                           "gemm_tc_kernel launches of one agent step (ncu, cold cache per launch), average per launch"
                           f"; this run launches {g_calls} per step; algorithmic bytes of these launches: "
                           f"{g_by / 1e9:.2f} GB per step")
            break
        except Exception:
            continue
    return {"bound": "tensor", "kernel": "gemm_tc_kernel (gn_conv2d implicit GEMM + gn_linear)", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_bytes_per_step": traffic_step, "traffic_source": traffic_src, "peak_source": src,
            "launches_per_step": g_calls, "graph_launches_per_step": n,
            "kernel_ms_per_step": t_ms, "kernel_ms_sum_of_durations": sum_ms, "step_ms": step_ms,
            "graph_span_ms": (kernels or {}).get("__span_ms__"),
            "algorithmic_tflop_per_step": g_fl / 1e12, "algorithmic_gb_per_step": g_by / 1e9,
            "hbm_frac_of_measured": (g_by / (t_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
            "eager_event_pair_ms": classes["linear"]["ms"] + classes["conv"]["ms"],
            "kernels_in_graph": {k: v for k, v in sorted((kernels or {}).items(), key=lambda kv: -(kv[1]["busy_ms"] if isinstance(kv[1], dict) else 0))[:12]
                                 if isinstance(v, dict)},
            "how": "achieved = algorithmic FLOPs of the gn_linear / gn_conv2d calls of one agent step (counted by the "
                   "library, 2*M*N*K_real) / kernel_ms_per_step; kernel_ms_per_step = length of the union of the "
                   "gemm_tc_kernel intervals CUPTI records inside one REPLAYED step graph (torch.profiler on the graph "
                   "replay, warm, on the streams the graph launches on): the time a gemm_tc_kernel is executing, <= the "
                   f"step's {step_ms:.3f} ms"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--denoise-steps", type=int, default=5)
    ap.add_argument("--preset", default="sd-turbo", choices=["sd-turbo", "tiny"])
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the tile_batch 2 / 4 throughput variant")
    ap.add_argument("--no-roofline", action="store_true", help="skip the kernel-class profile / CUPTI pass")
    ap.add_argument("--gpu-baseline", default="eager", choices=["none", "eager", "all"],
                    help="stock PyTorch fp16 on the same GPU: 'eager' = eager + CUDA graph (seconds); 'all' adds "
                         "torch.compile reduce-overhead / max-autotune (minutes)")
    ap.add_argument("--autoencoder", default="", help="'taesd': decode with AutoencoderTiny (reference option "
                                                      "eval_cfg.autoencoder); default: the KL-VAE of the headline config")
    ap.add_argument("--replay", default="", metavar="TASKSxEPISODES[xIN_FLIGHT]",
                    help="BASELINE configs[3]/[4] instead of the step benchmark: the episode-parallel evaluation replay "
                         "(genima_b200.eval_replay) sharded over the launched ranks, e.g. 25x25x4 under torchrun "
                         "--nproc-per-node 8; prints its summary as the JSON line")
    args = ap.parse_args()
    if args.replay:
        from genima_b200 import eval_replay

        parts = [int(v) for v in args.replay.lower().split("x")]
        argv = ["--tasks", str(parts[0]), "--episodes", str(parts[1]), "--preset", args.preset, "--denoise-steps",
                str(args.denoise_steps), "--episodes-in-flight", str(parts[2] if len(parts) > 2 else 1)]
        return eval_replay.main(argv)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
