"""GPU baseline beside the product path (SURVEY.md §8d "GPU baseline beside it", BASELINE.md "Stock PyTorch on the same
B200"): the oracle's restatement of the reference's graphs (oracle/sd_models.py, oracle/act.py) executed by STOCK PyTorch
on the GPU — cuDNN / cuBLAS / F.scaled_dot_product_attention library kernels, i.e. what the reference's pipeline would
run on this box if its stack could be installed.  Precision as in the reference: the diffusion models in fp16
(controller/agent/sd_controlnet_agent.py:32-42), the ACT controller in fp32.  Same synthetic
weights and inputs as bench.py; one "step" = the full agent step: 5 x (ControlNet + U-Net) + Euler updates + KL-VAE
decode + postprocess + untile + ACT.

Modes (`--modes`, comma separated):
  eager                     plain eager execution
  graph                     the whole step as one CUDA-graph replay (launch overhead removed)
  compile-reduce-overhead   torch.compile(mode="reduce-overhead") on the U-Net and the ControlNet — the reference's own
                            fast path (README.md:260, sd_controlnet_agent.py:52-62; the text encoder is cached here)
  compile-max-autotune      torch.compile(mode="max-autotune") on the same two networks
Lives under tests/ because it executes oracle/ code; it is a measurement script (bench.py's gpu_baseline leg runs it in
a subprocess), not a pytest module.
Usage (GPU box): python tests/stock_torch_gpu_baseline.py --modes eager,graph > gpurun_out/stock_torch.json"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genima_b200 import weights as W  # noqa: E402
from oracle import act as act_oracle  # noqa: E402
from oracle import sd_models  # noqa: E402
from oracle.scheduler import EulerDiscreteOracle  # noqa: E402


class _T:
    def __init__(self, t):
        self.t = t

    def to(self, *a, **k):
        return self.t

    def float(self):
        return self.t


class DeviceSD(dict):
    """State dict whose `.to(torch.float32)` / `.float()` requests yield the CUDA tensors of the wanted precision: the
    oracle graph then runs in that precision on the GPU."""

    def __init__(self, sd, dtype):
        super().__init__({k: _T(v.to("cuda", dtype)) for k, v in sd.items()})


def attention_sdpa(sd, p, x, ctx, heads):
    """oracle.sd_models.attention with the softmax(QK^T)V core replaced by torch's fused SDPA (what diffusers'
    AttnProcessor2_0 calls)."""
    b, n, c = x.shape
    w = lambda k: sd[k].to(torch.float32)  # noqa: E731
    d = c // heads
    q = F.linear(x, w(f"{p}.to_q.weight")).reshape(b, n, heads, d).transpose(1, 2)
    k = F.linear(ctx, w(f"{p}.to_k.weight")).reshape(b, -1, heads, d).transpose(1, 2)
    v = F.linear(ctx, w(f"{p}.to_v.weight")).reshape(b, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
    return F.linear(o, w(f"{p}.to_out.0.weight"), w(f"{p}.to_out.0.bias"))


sd_models.attention = attention_sdpa


def make_chain(n_steps: int, inputs=None):
    """-> (chain, nets): `chain()` runs one full agent step with stock PyTorch on the GPU and returns
    (a_hat [1, nq, A] fp32, image [1, 3, 2S, 2S] fp16 in [-1, 1], tile uint8 [3, 2S, 2S]); `nets` holds the two
    callables torch.compile may replace.  inputs: optional (views, qpos, task, ctx, lat) instead of bench.make_inputs."""
    ucfg, vcfg, acfg = bench.presets("sd-turbo")
    usd = DeviceSD(W.synth_state_dict(W.unet_shapes(ucfg)), torch.float16)
    csd = DeviceSD(W.synth_state_dict(W.controlnet_shapes(ucfg), salt=1), torch.float16)
    vsd = DeviceSD(W.synth_state_dict(W.vae_decoder_shapes(vcfg), salt=2), torch.float16)
    asd = {k: v.to("cuda", torch.float32) for k, v in W.synth_state_dict(W.act_shapes(acfg), salt=3).items()}
    views, qpos, task, ctx, lat = inputs if inputs is not None else bench.make_inputs(ucfg, acfg)
    S = acfg.image_size
    v = views.permute(0, 2, 3, 1)
    tile = torch.cat([torch.cat([v[0], v[1]], 1), torch.cat([v[2], v[3]], 1)], 0)[None]         # tile_images
    cond = (tile.float() / 255.0).permute(0, 3, 1, 2).contiguous().to("cuda", torch.float16)
    ctx = ctx.to("cuda", torch.float16)
    lat = lat.to("cuda", torch.float16)
    qpos, task = qpos.cuda().float(), task.cuda().float()
    sched = EulerDiscreteOracle()
    ts, sig = sched.set_timesteps(n_steps)
    tts = [torch.tensor([float(t)], device="cuda") for t in ts]

    def controlnet(xs, t):
        return sd_models.controlnet_forward(csd, ucfg, xs, t, ctx, cond)

    def unet(xs, t, down, mid):
        return sd_models.unet_forward(usd, ucfg, xs, t, ctx, down, mid)

    nets = {"controlnet": controlnet, "unet": unet, "controlnet_eager": controlnet, "unet_eager": unet}

    def chain():
        x = lat * sched.init_noise_sigma
        for i in range(n_steps):
            if nets.get("compiled"):
                # CUDA-graph trees of torch.compile (reduce-overhead / max-autotune): a new denoise iteration may reuse
                # the memory of the previous iteration's graph outputs (eps and the residuals were consumed eagerly)
                torch.compiler.cudagraph_mark_step_begin()
            xs = sched.scale_model_input(x, i)
            down, mid = nets["controlnet"](xs, tts[i])
            eps = nets["unet"](xs, tts[i], down, mid)
            x = sched.step(eps, i, x).to(torch.float16)
        img = sd_models.vae_decode(vsd, vcfg, x / vcfg.scaling_factor)
        u8 = ((img / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)[0]                   # [3, 2S, 2S]
        gen = torch.stack([u8[:, :S, :S], u8[:, :S, S:], u8[:, S:, :S], u8[:, S:, S:]], 0)    # untile_images
        with torch.device("cuda"):
            a_hat, _ = act_oracle.act_forward(asd, acfg, qpos, gen[None].float(), task)
        return a_hat, img, u8

    return chain, nets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="eager,graph")
    ap.add_argument("--denoise-steps", type=int, default=5)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    modes = [m for m in args.modes.split(",") if m]
    n_steps = args.denoise_steps
    full_chain, nets = make_chain(n_steps)
    controlnet, unet = nets["controlnet_eager"], nets["unet_eager"]

    def chain():
        return full_chain()[0]

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    out = {"what": "oracle graphs (diffusers / RoboBase restatement) run by stock PyTorch on the GPU: full agent step = "
                   f"{n_steps} x (ControlNet + U-Net, fp16, SDPA) + Euler + KL-VAE decode + untile + ACT (fp32), batch 1, "
                   "512x512 tile, synthetic weights, NCHW",
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "unit": "ms per agent step"}
    with torch.no_grad():
        ref = None
        if "eager" in modes:
            ms, ref = timed(chain, args.reps, 2)
            out["eager_ms"] = round(ms, 2)
            out["eager_steps_per_s"] = round(1e3 / ms, 2)
        if "graph" in modes:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side), torch.device("cuda"):
                    chain()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph), torch.device("cuda"):
                    chain()
                ms_g, _ = timed(graph.replay, args.reps, 2)
                out["cuda_graph_ms"] = round(ms_g, 2)
                out["cuda_graph_steps_per_s"] = round(1e3 / ms_g, 2)
                del graph
            except Exception as ex:  # noqa: BLE001
                out["cuda_graph_ms"] = None
                out["cuda_graph_error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
        for mode in ("reduce-overhead", "max-autotune"):
            key = "compile_" + mode.replace("-", "_")
            if f"compile-{mode}" not in modes:
                continue
            try:
                torch._dynamo.reset()
                t0 = time.time()
                nets["controlnet"] = torch.compile(controlnet, mode=mode, fullgraph=False)
                nets["unet"] = torch.compile(unet, mode=mode, fullgraph=False)
                nets["compiled"] = True
                ms_c, got = timed(chain, args.reps, 3)
                out[key + "_ms"] = round(ms_c, 2)
                out[key + "_steps_per_s"] = round(1e3 / ms_c, 2)
                out[key + "_compile_s"] = round(time.time() - t0 - ms_c * args.reps / 1e3, 1)
                if ref is not None:
                    out[key + "_max_abs_diff_vs_eager"] = float((got - ref).abs().max())
            except Exception as ex:  # noqa: BLE001
                out[key + "_ms"] = None
                out[key + "_error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
            finally:
                nets["controlnet"], nets["unet"] = controlnet, unet
                nets["compiled"] = False
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
