"""The fused device-resident agent step (genima_b200/step.py) and the reference-facing plugin classes
(genima_b200/agents.py) against the CPU oracle of the whole step (oracle/pipeline.py::agent_step), on weights bound as
views into the broadcast arena (the layout bench.py and the multi-GPU driver use)."""
import numpy as np
import pytest
import torch

from genima_b200 import distributed as gd
from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, UNetConfig, VAEConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny(ops):
    from genima_b200.act_policy import DeviceACT
    from genima_b200.pipeline import B200ControlNetPipeline

    ucfg, vcfg, acfg = UNetConfig.tiny(), VAEConfig.tiny(), ACTConfig.tiny()
    shapes = dict(unet=W.unet_shapes(ucfg), controlnet=W.controlnet_shapes(ucfg), vae=W.vae_decoder_shapes(vcfg),
                  act=W.act_shapes(acfg))
    host = dict(unet=W.synth_state_dict(shapes["unet"]), controlnet=W.synth_state_dict(shapes["controlnet"], 1),
                vae=W.synth_state_dict(shapes["vae"], 2), act=W.synth_state_dict(shapes["act"], 3))
    dev_sds, arena = gd.broadcast_weights(shapes, host, device=ops.device)
    pipe = B200ControlNetPipeline(ops, dev_sds["unet"], dev_sds["controlnet"], dev_sds["vae"], None, ucfg, vcfg)
    act = DeviceACT(ops, dev_sds["act"], acfg)
    g = torch.Generator().manual_seed(0)
    S = acfg.image_size
    inputs = dict(views=torch.randint(0, 256, (4, S, S, 3), generator=g, dtype=torch.uint8),
                  ctx=torch.randn(1, 77, ucfg.cross_attention_dim, generator=g).half().float(),
                  lat=torch.randn(1, 4, S // 4, S // 4, generator=torch.Generator().manual_seed(2)).half().float(),
                  qpos=torch.randn(1, acfg.state_dim, generator=g), task=torch.randn(1, acfg.task_emb_dim, generator=g))
    return dict(pipe=pipe, act=act, host=host, cfgs=(ucfg, vcfg, acfg), inputs=inputs, arena=arena)


@pytest.mark.parametrize("graph", [False, True])
def test_agent_step_matches_oracle(tiny, graph):
    from genima_b200.step import GenimaStep
    from oracle.pipeline import agent_step

    ucfg, vcfg, acfg = tiny["cfgs"]
    i = tiny["inputs"]
    ref = agent_step(tiny["host"], ucfg, vcfg, acfg, i["views"].numpy(), i["ctx"], i["lat"], i["qpos"], i["task"], 3)
    step = GenimaStep(tiny["pipe"], tiny["act"], num_inference_steps=3, use_cuda_graph=graph)
    task_dev = i["task"].cuda()
    for rep in range(2):      # second call replays the captured graph
        out = step(i["views"][None].cuda(), i["lat"].cuda(), i["qpos"].cuda(), task_dev, prompt_embeds=i["ctx"])
        a = out["a_hat"].float().cpu()
        err = float((a - ref["a_hat"]).abs().max() / ref["a_hat"].abs().max())
        d = np.abs(out["tile_u8"].cpu().numpy().astype(np.int32) - ref["tile_u8"].astype(np.int32))
        print(f"agent step (graph={graph}, call {rep}): a_hat normalised max err {err:.3e}; tile max |diff| {d.max()}; "
              f"{step.launches_per_step} launches")
        assert err < 2e-3 and d.max() <= 3 and (d <= 1).mean() > 0.995   # measured 5.5e-4
    assert step.launches_per_step > 100


def test_plugin_api_path_equals_fused_step(tiny):
    """agent.infer(PIL) -> untile (host) -> controller.act(obs) must give the fused step's actions (same kernels)."""
    from PIL import Image

    from genima_b200.agents import B200ControlNetAgent, B200GenimaACT, B200GenimaACTPolicy
    from genima_b200.step import GenimaStep

    ucfg, vcfg, acfg = tiny["cfgs"]
    i = tiny["inputs"]
    S = acfg.image_size
    agent = B200ControlNetAgent.__new__(B200ControlNetAgent)
    agent.eval_cfg = dict(image_resolution=2 * S, device="cuda")
    agent.pipe, agent._ops = tiny["pipe"], tiny["pipe"].ops
    agent.set_optimizations()
    agent.common_setup()
    policy = B200GenimaACTPolicy.__new__(B200GenimaACTPolicy)
    policy.cfg, policy.ops, policy._sd, policy.impl, policy.training = acfg, tiny["pipe"].ops, None, tiny["act"], False
    ctrl = B200GenimaACT(policy)
    tokens = torch.zeros(1, 1, 77, dtype=torch.int32)
    task_dev = i["task"].cuda()
    ctrl._emb_cache[tokens.reshape(-1, 77).numpy().tobytes() + repr(tuple(tokens.shape)).encode()] = (task_dev, None)
    v = i["views"].numpy()
    tile = np.concatenate([np.concatenate([v[0], v[1]], 1), np.concatenate([v[2], v[3]], 1)], 0)
    out = agent.infer(images=[Image.fromarray(tile)], prompts=None, negative_prompts=None, num_inference_steps=3,
                      guidance_scale=0.0, generator=None, latents=i["lat"], prompt_embeds=i["ctx"])
    assert isinstance(out[0], list) and out[0][0].size == (2 * S, 2 * S) and out.nsfw_content_detected is None
    g = np.asarray(out[0][0])
    quads = [g[:S, :S], g[:S, S:], g[S:, :S], g[S:, S:]]
    obs = {f"cam{k}_rgb": torch.from_numpy(np.ascontiguousarray(np.transpose(q, (2, 0, 1))[None])).cuda().unsqueeze(0)
           for k, q in enumerate(quads)}
    obs["low_dim_state"] = i["qpos"][None].cuda()
    obs["lang_tokens"] = tokens
    a_api = ctrl.act(obs, step=0, eval_mode=True)
    assert tuple(a_api.shape) == (1, acfg.num_queries, acfg.action_dim)
    fused = GenimaStep(tiny["pipe"], tiny["act"], num_inference_steps=3, use_cuda_graph=False)
    a_fused = fused(i["views"][None].cuda(), i["lat"].cuda(), i["qpos"].cuda(), task_dev, prompt_embeds=i["ctx"])
    err = float((a_api.float().cpu() - a_fused["a_hat"].float().cpu()).abs().max())
    print(f"plugin API path vs fused step: max |a_hat diff| {err:.3e}")
    # same kernels, same tile configurations, integer GroupNorm accumulators: bit-identical
    assert torch.equal(a_api.float().cpu(), a_fused["a_hat"].float().cpu())


def test_policy_rejects_training_call(tiny):
    from genima_b200.agents import B200GenimaACTPolicy

    policy = B200GenimaACTPolicy.__new__(B200GenimaACTPolicy)
    policy.cfg, policy.impl = tiny["cfgs"][2], tiny["act"]
    with pytest.raises(NotImplementedError):
        policy.forward(torch.zeros(1, 8), torch.zeros(1, 4, 3, 64, 64), actions=torch.zeros(1, 20, 8),
                       task_emb=torch.zeros(1, 64))


def test_task_switch_a_b_a_replays_the_right_film(tiny):
    """a cached CUDA graph reads the FiLM affines of ITS task embedding by raw pointer.  Task sequence
    A -> B -> A (and enough other tasks in between to evict A from the FiLM cache) must replay A's graph on A's
    affines, equal to an eager run."""
    import gc

    acfg = tiny["cfgs"][2]
    act = tiny["act"]
    g = torch.Generator().manual_seed(5)
    S = acfg.image_size
    image = torch.randint(0, 256, (1, 4, 3, S, S), generator=g, dtype=torch.uint8).cuda()
    qpos = torch.randn(1, acfg.state_dim, generator=g).cuda()
    tasks = [torch.randn(1, acfg.task_emb_dim, generator=g).cuda() for _ in range(12)]
    eager = [act.forward(qpos, image, t)[0].clone() for t in tasks[:2]]
    assert not torch.equal(eager[0], eager[1])
    a0 = act.forward_graphed(qpos, image, tasks[0])[0].clone()
    b0 = act.forward_graphed(qpos, image, tasks[1])[0].clone()
    for t in tasks[2:]:                       # churn: evicts A / B from the FiLM cache and recycles allocator blocks
        act.film_affines(t)
        junk = [torch.randn(64, 2 * 64, device="cuda") for _ in range(8)]
        del junk
    gc.collect()
    a1 = act.forward_graphed(qpos, image, tasks[0])[0].clone()
    b1 = act.forward_graphed(qpos, image, tasks[1])[0].clone()
    assert torch.equal(a0, eager[0]) and torch.equal(b0, eager[1])
    assert torch.equal(a1, eager[0]) and torch.equal(b1, eager[1])


def test_many_prompts_do_not_invalidate_cached_graphs(tiny):
    """clearing the per-step time-embedding / K-V caches (many distinct prompts or step counts) must not free
    buffers a still-cached graph replays against."""
    pipe = tiny["pipe"]
    i = tiny["inputs"]
    S = tiny["cfgs"][2].image_size
    was = pipe.use_cuda_graph
    pipe.use_cuda_graph = True
    try:
        g = torch.Generator().manual_seed(9)
        tile = torch.randint(0, 256, (1, 2 * S, 2 * S, 3), generator=g, dtype=torch.uint8)
        kw = dict(image=tile, num_inference_steps=2, guidance_scale=0.0, latents=i["lat"], output_type="latent")
        ctx_a = i["ctx"]
        first = pipe(prompt_embeds=ctx_a, **kw).images.clone()
        pipe._temb_cache.clear()              # what 17 distinct (steps, batch, added) keys do
        pipe._kv_cache.clear()
        pipe._ctx_cache.clear()
        others = [torch.randn(1, 77, ctx_a.shape[-1], generator=g).half().float() for _ in range(3)]
        for c in others:
            pipe(prompt_embeds=c, **kw)
        junk = [torch.randn(1024, 1024, device="cuda") for _ in range(4)]
        del junk
        again = pipe(prompt_embeds=ctx_a, **kw).images
        assert torch.equal(first, again)
    finally:
        pipe.use_cuda_graph = was
