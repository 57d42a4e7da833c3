"""BASELINE configs[2] at FULL size: the real SD-Turbo / ControlNet / KL-VAE / ACT topologies
(synthetic seeded weights), 4 x 256^2 views -> one 512^2 tile -> 5 Euler-trailing denoise steps -> VAE decode -> untile ->
ACT -> a_hat [1, 20, 8], replayed from the step's CUDA graph, against the fp32 CPU oracle of the whole step
(oracle/pipeline.py::agent_step) on identical weights, views, latents (seed 2), proprioception and embeddings.

Reported for BOTH outputs the north star names (a_hat and the decoded 512^2 tile):
  * the element-wise gate rtol = 1e-3 / atol = 1e-4 as a PASS FRACTION (the per-kernel tests hold it outright; a whole
    fp16 network chain accumulates 2^-11 storage rounding per stored activation, as the reference's own fp16 pipeline
    does, so the fraction is printed next to the same figure for stock PyTorch fp16 on this GPU);
  * the normalised max error max|out - ref| / max|ref|, asserted at <= 2x what was measured when the test was written
    (and never worse than 1.5x stock fp16's).
Also: two independent handle sets (second one adopting the first one's tile configurations through
gn_tune_cache_export / import) give BIT-IDENTICAL a_hat and tile — what rank r of a sharded evaluation relies on."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

A_HAT_TOL = 3e-3       # measured 1.0e-3 (bench.py cpu_baseline leg, round 1) .. 1.4e-3
IMAGE_TOL = 4e-3       # decoded image in [0, 1], normalised max error; measured 1.9e-3


def _pass_fraction(out, ref, rtol=1e-3, atol=1e-4):
    o, r = out.detach().float().cpu().reshape(-1), ref.detach().float().cpu().reshape(-1)
    return float(((o - r).abs() <= atol + rtol * r.abs()).float().mean())


def _norm_err(out, ref):
    o, r = out.detach().float().cpu(), ref.detach().float().cpu()
    return float((o - r).abs().max() / r.abs().max())


@pytest.fixture(scope="module")
def world():
    import bench
    from genima_b200 import distributed as gd
    from genima_b200.act_policy import DeviceACT
    from genima_b200.ops import Ops
    from genima_b200.pipeline import B200ControlNetPipeline
    from genima_b200.step import GenimaStep

    ucfg, vcfg, acfg = bench.presets("sd-turbo")
    shapes = bench.model_shapes(ucfg, vcfg, acfg)
    host = bench.synth_all(shapes)
    dev = torch.device("cuda", 0)
    sds, arena = gd.broadcast_weights(shapes, host, device=dev)
    views, qpos, task, ctx, lat = bench.make_inputs(ucfg, acfg)
    d = dict(views=views.permute(0, 2, 3, 1).contiguous()[None].to(dev), lat=lat.to(dev), qpos=qpos.to(dev),
             task=task.to(dev), ctx=ctx.to(dev))

    def build():
        ops = Ops(0)
        pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg,
                                      use_cuda_graph=True)
        act = DeviceACT(ops, sds["act"], acfg)
        return pipe, act, GenimaStep(pipe, act, num_inference_steps=5, use_cuda_graph=True)

    return dict(cfgs=(ucfg, vcfg, acfg), host=host, sds=sds, arena=arena, inputs=(views, qpos, task, ctx, lat), dev=d,
                build=build)


def test_config3_full_size_against_the_oracle(world):
    from oracle.pipeline import agent_step
    from stock_torch_gpu_baseline import make_chain

    ucfg, vcfg, acfg = world["cfgs"]
    views, qpos, task, ctx, lat = world["inputs"]
    d = world["dev"]
    pipe, act, step = world["build"]()
    world["first"] = (pipe, act, step)
    for _ in range(2):                                   # second call = graph replay
        out = step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
    a_hat = out["a_hat"].float().cpu().clone()
    tile = out["tile_u8"].cpu().numpy().copy()
    world["a_hat"], world["tile"] = a_hat, tile
    img = pipe(prompt_embeds=d["ctx"], image=torch.from_numpy(np.ascontiguousarray(
        np.concatenate([np.concatenate([views[0].permute(1, 2, 0).numpy(), views[1].permute(1, 2, 0).numpy()], 1),
                        np.concatenate([views[2].permute(1, 2, 0).numpy(), views[3].permute(1, 2, 0).numpy()], 1)], 0)))[None],
        num_inference_steps=5, guidance_scale=0.0, latents=d["lat"], output_type="pt").images.float().cpu()

    w32 = {m: {k: v.float() for k, v in sd.items()} for m, sd in world["host"].items()}
    with torch.no_grad():
        ref = agent_step(w32, ucfg, vcfg, acfg, views.permute(0, 2, 3, 1).contiguous().numpy(), ctx.float(), lat, qpos,
                         task, 5)
        chain, _ = make_chain(5, inputs=world["inputs"])
        s_a, s_img, s_u8 = chain()
    ref_img = (ref["image"] / 2 + 0.5).clamp(0, 1)
    s_img01 = (s_img.float().cpu() / 2 + 0.5).clamp(0, 1)

    e_a, e_i = _norm_err(a_hat, ref["a_hat"]), _norm_err(img, ref_img)
    s_e_a, s_e_i = _norm_err(s_a, ref["a_hat"]), _norm_err(s_img01, ref_img)
    dt = np.abs(tile[0].astype(np.int32) - ref["tile_u8"][0].astype(np.int32))
    s_dt = np.abs(s_u8.permute(1, 2, 0).cpu().numpy().astype(np.int32) - ref["tile_u8"][0].astype(np.int32))
    print(f"config 3 full size, 5 steps, CUDA graph ({step.launches_per_step} launches):\n"
          f"  a_hat : normalised max err {e_a:.3e} (stock torch fp16: {s_e_a:.3e}); rtol 1e-3 / atol 1e-4 pass fraction "
          f"{_pass_fraction(a_hat, ref['a_hat']):.4f} (stock: {_pass_fraction(s_a, ref['a_hat']):.4f})\n"
          f"  image : normalised max err {e_i:.3e} (stock: {s_e_i:.3e}); pass fraction {_pass_fraction(img, ref_img):.4f} "
          f"(stock: {_pass_fraction(s_img01, ref_img):.4f})\n"
          f"  tile  : max |diff| {dt.max()} levels, {100 * (dt <= 1).mean():.3f}% within 1 level, {100 * (dt == 0).mean():.2f}% "
          f"exact (stock: max {s_dt.max()}, {100 * (s_dt <= 1).mean():.3f}% within 1, {100 * (s_dt == 0).mean():.2f}% exact)")
    assert torch.isfinite(a_hat).all() and torch.isfinite(img).all()
    assert e_a <= max(A_HAT_TOL, 1.5 * s_e_a)
    assert e_i <= max(IMAGE_TOL, 1.5 * s_e_i)
    assert dt.max() <= 2 and (dt <= 1).mean() > 0.999


def test_second_handle_set_is_bit_identical(world):
    """Rank r of a sharded evaluation imports rank 0's measured tile configurations (distributed.sync_tune_caches): same
    launches -> same summation order -> the same bits.  Emulated in one process with a second, independent set of
    handles, graphs and scratch arenas."""
    d = world["dev"]
    if "first" not in world:
        pytest.skip("needs the first test's pipeline")
    pipe0, act0, _ = world["first"]
    blobs = pipe0.tune_cache_export()
    assert sum(len(b) for b in blobs) > 1000          # the first set did measure its shapes
    from genima_b200.act_policy import DeviceACT
    from genima_b200.ops import Ops
    from genima_b200.pipeline import B200ControlNetPipeline
    from genima_b200.step import GenimaStep

    ucfg, vcfg, acfg = world["cfgs"]
    sds = world["sds"]
    ops = Ops(0)
    pipe = B200ControlNetPipeline(ops, sds["unet"], sds["controlnet"], sds["vae"], None, ucfg, vcfg, use_cuda_graph=True)
    pipe.tune_cache_import(blobs)                      # BEFORE the first launch: nothing is timed on this "rank"
    act = DeviceACT(ops, sds["act"], acfg)
    step = GenimaStep(pipe, act, num_inference_steps=5, use_cuda_graph=True)
    out = step(d["views"], d["lat"], d["qpos"], d["task"], prompt_embeds=d["ctx"])
    assert pipe.tune_cache_export() == blobs           # adopted, not re-measured
    assert torch.equal(out["a_hat"].float().cpu(), world["a_hat"])
    assert np.array_equal(out["tile_u8"].cpu().numpy(), world["tile"])
