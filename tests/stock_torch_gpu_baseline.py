"""GPU baseline beside the product path (SURVEY.md §8d "GPU baseline beside it", BASELINE.md "Stock PyTorch on the same
B200"): the oracle's restatement of the diffusers graph (oracle/sd_models.py) executed by STOCK PyTorch in fp16 on the
GPU — cuDNN / cuBLAS / F.scaled_dot_product_attention library kernels, i.e. what the reference's fp16 pipeline would run
on this box if its stack could be installed — eagerly and as one CUDA-graph replay (launch overhead removed, the part
of `torch.compile(mode="reduce-overhead")` that matters at batch 1).  Same synthetic weights and inputs as bench.py;
covers the diffusion part of the agent step (5 ControlNet + U-Net evaluations, Euler updates, KL-VAE decode).
Lives under tests/ because it executes oracle/ code; it is a measurement script, not a pytest module.
Usage (GPU box): python tests/stock_torch_gpu_baseline.py > gpurun_out/stock_torch.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genima_b200 import weights as W  # noqa: E402
from oracle import sd_models  # noqa: E402
from oracle.scheduler import EulerDiscreteOracle  # noqa: E402


class HalfSD(dict):
    """State dict whose `.to(torch.float32)` requests yield CUDA fp16 tensors: the oracle graph then runs in fp16."""

    class _T:
        def __init__(self, t):
            self.t = t

        def to(self, *a, **k):
            return self.t

    def __init__(self, sd):
        super().__init__({k: HalfSD._T(v.to("cuda", torch.float16)) for k, v in sd.items()})


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


def main():
    n_steps = 5
    ucfg, vcfg, acfg = bench.presets("sd-turbo")
    usd = HalfSD(W.synth_state_dict(W.unet_shapes(ucfg)))
    csd = HalfSD(W.synth_state_dict(W.controlnet_shapes(ucfg), salt=1))
    vsd = HalfSD(W.synth_state_dict(W.vae_decoder_shapes(vcfg), salt=2))
    views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
    g = torch.Generator().manual_seed(0)
    cond = (torch.randint(0, 256, (1, 512, 512, 3), generator=g, dtype=torch.uint8).float() / 255.0)
    cond = cond.permute(0, 3, 1, 2).contiguous().to("cuda", torch.float16)
    ctx = ctx.to("cuda", torch.float16)
    lat = lat.to("cuda", torch.float16)
    sched = EulerDiscreteOracle()
    ts, sig = sched.set_timesteps(n_steps)
    tts = [torch.tensor([float(t)], device="cuda") for t in ts]

    def chain():
        x = lat * sched.init_noise_sigma
        for i in range(n_steps):
            xs = sched.scale_model_input(x, i)
            down, mid = sd_models.controlnet_forward(csd, ucfg, xs, tts[i], ctx, cond)
            eps = sd_models.unet_forward(usd, ucfg, xs, tts[i], ctx, down, mid)
            x = sched.step(eps, i, x).to(torch.float16)
        img = sd_models.vae_decode(vsd, vcfg, x / vcfg.scaling_factor)
        return ((img / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)

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

    out = {"what": "oracle graph (diffusers restatement) run by stock PyTorch fp16 on the GPU: 5 x (ControlNet + U-Net) + "
                   "Euler + KL-VAE decode, batch 1, 512x512 tile, synthetic weights; NCHW, SDPA attention",
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}
    with torch.no_grad():
        ms, img = timed(chain, 5, 2)
        out["eager_ms"] = round(ms, 2)
        out["image_ok"] = bool(img.shape == (1, 3, 512, 512))
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                chain()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                chain()
            ms_g, _ = timed(graph.replay, 10, 2)
            out["cuda_graph_ms"] = round(ms_g, 2)
        except Exception as ex:  # noqa: BLE001
            out["cuda_graph_ms"] = None
            out["cuda_graph_error"] = f"{type(ex).__name__}: {str(ex)[:200]}"
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
