"""Plugin point 2: `B200GenimaACT` must be constructible and loadable exactly the way
`GenimaEvalWorkspace` treats `method.genima_act.GenimaACT` (controller/eval_genima.py:55-66, 91-103, 199-201) — replayed
here on the CPU with the stand-ins of tests/fake_workspace.py.  Nothing below needs a GPU: weights are bound to the
device lazily at the first `act`, which must then fail loudly without CUDA."""
import os

import pytest
import torch

import fake_workspace as fw
from genima_b200 import checkpoint as ckpt
from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, CLIPTextConfig


def _build(**over):
    fw.install_reference_stub_modules()
    node = dict(fw.GENIMA_ACT_YAML, _target_="genima_b200.agents.B200GenimaACT")
    obs_space, act_space = fw.make_spaces(fw.CAMERAS, 1, 256, 8, 8)
    kw = dict(device=torch.device("cpu"), observation_space=obs_space, action_space=act_space, num_train_envs=1,
              replay_alpha=0.6, replay_beta=0.4, frame_stack_on_channel=False)
    kw.update(over)
    return fw.instantiate(node, **kw)          # eval_genima.py:55-64


def test_hydra_construction_reads_the_reference_config():
    agent = _build()
    assert agent.train(False) is agent and agent.training is False            # eval_genima.py:66
    assert agent.cfg == ACTConfig()               # genima_act.yaml hyper-parameters + the spaces give the full-size policy
    assert agent.hparams["lr"] == 5.0e-05 and agent.hparams["replay_alpha"] == 0.6
    keys = list(agent.state_dict().keys())
    assert keys and all(k.startswith("actor.") for k in keys)
    assert "actor.encoder_model.backbone.conv1.weight" in keys and "actor.actor_model.action_head.weight" in keys
    assert not any("clip" in k for k in keys)     # train_act.py:262-279 strips clip_model.* before saving
    with pytest.raises(NotImplementedError):
        agent.update(None, 0)


def test_spaces_override_the_yaml():
    obs_space, act_space = fw.make_spaces(fw.CAMERAS[:2], 2, 128, 10, 7)
    agent = _build(observation_space=obs_space, action_space=act_space)
    assert (agent.cfg.state_dim, agent.cfg.num_views, agent.cfg.image_size, agent.cfg.action_dim) == (20, 4, 128, 7)
    node = dict(fw.GENIMA_ACT_YAML, _target_="genima_b200.agents.B200GenimaACT")
    node["actor_model"] = dict(node["actor_model"], pre_norm=True)
    with pytest.raises(NotImplementedError):
        fw.instantiate(node, device="cpu", observation_space=obs_space, action_space=act_space)


def test_load_controller_ckpt_replay(tmp_path):
    """eval_genima.py:91-103 against a RoboBase-format snapshot (train_act.py:262-279)."""
    acfg = ACTConfig()
    sd = W.synth_state_dict(W.act_shapes(acfg), salt=3)
    payload = {"cfg": {}, "_epoch": 3, "_num_iters": 7,
               "agent": {**{f"actor.{k}": v for k, v in sd.items()},
                         # RoboBase registers the encoder / actor_model modules on the agent too [U]: extra aliases
                         "encoder.backbone.conv1.weight": sd["encoder_model.backbone.conv1.weight"]}}
    path = str(tmp_path / "latest.pt")
    torch.save(payload, path)
    agent = _build()
    res = fw.load_controller_ckpt(agent, path)
    assert res.missing_keys == [] and res.unexpected_keys == ["encoder.backbone.conv1.weight"]
    got = agent.state_dict()
    assert set(got) == {f"actor.{k}" for k in sd} and all(torch.equal(got[f"actor.{k}"], v) for k, v in sd.items())
    with pytest.raises(RuntimeError):                 # strict=True behaves like torch: unexpected keys raise
        agent.load_state_dict(payload["agent"], strict=True)
    # a snapshot that lacks a policy tensor is refused by the workspace's own missing-key check ...
    broken = dict(payload, agent={k: v for k, v in payload["agent"].items() if k != "actor.actor_model.action_head.bias"})
    torch.save(broken, path)
    with pytest.raises(ValueError, match="Missing keys"):
        fw.load_controller_ckpt(_build(), path)
    # ... and by load_state_dict itself (nothing can run without it), even with strict=False
    with pytest.raises(RuntimeError, match="missing keys"):
        _build().load_state_dict(broken["agent"], strict=False)
    # wrong shape: the schema check names the tensor
    bad = dict(payload["agent"])
    bad["actor.actor_model.action_head.bias"] = torch.zeros(3)
    with pytest.raises(ValueError):
        _build().load_state_dict(bad, strict=False)


def test_eval_mode_and_loud_failure_without_cuda(tmp_path):
    agent = _build()
    agent.train(True)
    with fw.eval_mode(agent):                         # eval_genima.py:199-201
        assert agent.training is False
    assert agent.training is True
    with pytest.raises(RuntimeError, match="before load_state_dict"):
        agent.act({"low_dim_state": torch.zeros(1, 1, 8)}, step=0, eval_mode=True)
    if not torch.cuda.is_available():
        from genima_b200._cabi import GenimaB200Error

        sd = W.synth_state_dict(W.act_shapes(ACTConfig.tiny()), salt=3)
        tiny = _build(act_cfg=ACTConfig.tiny())
        tiny.load_state_dict({f"actor.{k}": v for k, v in sd.items()}, strict=False)
        obs = {f"{c}_rgb": torch.zeros(1, 1, 3, 64, 64, dtype=torch.uint8) for c in fw.CAMERAS}
        obs.update(low_dim_state=torch.zeros(1, 1, 8), lang_tokens=torch.zeros(1, 1, 77, dtype=torch.int32))
        with pytest.raises(GenimaB200Error):          # no CPU fallback: binding the policy needs the CUDA library
            tiny.act(obs, step=0, eval_mode=True)


def _openai_clip_text_sd(cfg: CLIPTextConfig, seed: int = 0):
    """Random text tower in OpenAI CLIP naming (what `clip.load` returns; genima_act.py:324-341 reads these names)."""
    g = torch.Generator().manual_seed(seed)
    d, ff = cfg.hidden_size, cfg.intermediate_size
    r = lambda *s: torch.randn(*s, generator=g) * 0.05  # noqa: E731
    sd = {"token_embedding.weight": r(cfg.vocab_size, d), "positional_embedding": r(cfg.max_positions, d),
          "ln_final.weight": 1 + r(d), "ln_final.bias": r(d), "text_projection": r(d, cfg.projection_dim),
          "visual.conv1.weight": r(4, 3, 2, 2), "logit_scale": torch.tensor(1.0)}
    for i in range(cfg.num_layers):
        p = f"transformer.resblocks.{i}"
        sd.update({f"{p}.attn.in_proj_weight": r(3 * d, d), f"{p}.attn.in_proj_bias": r(3 * d),
                   f"{p}.attn.out_proj.weight": r(d, d), f"{p}.attn.out_proj.bias": r(d),
                   f"{p}.ln_1.weight": 1 + r(d), f"{p}.ln_1.bias": r(d), f"{p}.ln_2.weight": 1 + r(d),
                   f"{p}.ln_2.bias": r(d), f"{p}.mlp.c_fc.weight": r(ff, d), f"{p}.mlp.c_fc.bias": r(ff),
                   f"{p}.mlp.c_proj.weight": r(d, ff), f"{p}.mlp.c_proj.bias": r(d)})
    return sd


def _encode_clip_text_openai(sd, cfg: CLIPTextConfig, tokens: torch.Tensor):
    """GenimaACT.encode_clip_text (controller/method/genima_act.py:314-346) on an OpenAI-named state dict, in fp32:
    token + positional embedding, pre-LN residual attention blocks with a causal mask and QuickGELU (OpenAI CLIP
    `ResidualAttentionBlock`), ln_final, EOT (argmax) row @ text_projection, first frame."""
    import torch.nn.functional as F

    shape = tokens.shape
    tks = tokens.reshape(-1, shape[-1]).long()
    x = sd["token_embedding.weight"][tks] + sd["positional_embedding"]
    T, d, h = x.shape[1], cfg.hidden_size, cfg.num_heads
    mask = torch.full((T, T), float("-inf")).triu_(1)
    for i in range(cfg.num_layers):
        p = f"transformer.resblocks.{i}"
        y = F.layer_norm(x, (d,), sd[f"{p}.ln_1.weight"], sd[f"{p}.ln_1.bias"])
        qkv = y @ sd[f"{p}.attn.in_proj_weight"].t() + sd[f"{p}.attn.in_proj_bias"]
        q, k, v = (t.reshape(-1, T, h, d // h).transpose(1, 2) for t in qkv.split(d, dim=-1))
        a = torch.softmax(q @ k.transpose(-1, -2) / (d // h) ** 0.5 + mask, -1) @ v
        a = a.transpose(1, 2).reshape(-1, T, d)
        x = x + a @ sd[f"{p}.attn.out_proj.weight"].t() + sd[f"{p}.attn.out_proj.bias"]
        y = F.layer_norm(x, (d,), sd[f"{p}.ln_2.weight"], sd[f"{p}.ln_2.bias"])
        y = y @ sd[f"{p}.mlp.c_fc.weight"].t() + sd[f"{p}.mlp.c_fc.bias"]
        y = y * torch.sigmoid(1.702 * y)
        x = x + y @ sd[f"{p}.mlp.c_proj.weight"].t() + sd[f"{p}.mlp.c_proj.bias"]
    x = F.layer_norm(x, (d,), sd["ln_final.weight"], sd["ln_final.bias"])
    emb = x.clone()
    x = x[torch.arange(x.shape[0]), tks.argmax(dim=-1)] @ sd["text_projection"]
    return x.reshape(shape[0], shape[1], -1)[:, 0], emb


def test_openai_clip_names_convert_to_the_transformers_graph(tmp_path):
    """The `clip.load` state dict the reference uses, converted by checkpoint.openai_clip_text_to_hf, must give the same
    task embedding through transformers' CLIPTextModelWithProjection (the graph DeviceCLIPText implements and
    tests/test_oracle_pins.py pins) as the reference's own encode_clip_text arithmetic on the original names."""
    from transformers import CLIPTextConfig as HFConfig
    from transformers import CLIPTextModelWithProjection

    cfg = CLIPTextConfig(vocab_size=200, hidden_size=64, intermediate_size=128, num_layers=2, num_heads=4,
                         act="quick_gelu", projection_dim=32)
    osd = _openai_clip_text_sd(cfg)
    path = str(tmp_path / "ViT-tiny.pt")
    torch.save(osd, path)
    hsd = ckpt.load_openai_clip_text(path, cfg)
    assert not any(k.startswith("visual") for k in hsd)
    hf = CLIPTextModelWithProjection(HFConfig(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, max_position_embeddings=cfg.max_positions,
        hidden_act="quick_gelu", projection_dim=cfg.projection_dim, eos_token_id=2)).eval()
    missing, unexpected = hf.load_state_dict(hsd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(1, 150, (2, 1, 77), generator=g)
    tokens[:, :, 0] = 198
    tokens[0, 0, 9:] = 0
    tokens[0, 0, 8] = 199                            # EOT has the largest id (argmax pooling)
    tokens[1, 0, 20:] = 0
    tokens[1, 0, 19] = 199
    want, want_emb = _encode_clip_text_openai(osd, cfg, tokens)
    with torch.no_grad():
        out = hf(input_ids=tokens.reshape(-1, 77))
    pooled_in = out.last_hidden_state[torch.arange(2), tokens.reshape(-1, 77).argmax(-1)]
    got = pooled_in @ hsd["text_projection.weight"].t()
    assert torch.allclose(out.last_hidden_state, want_emb, rtol=1e-4, atol=1e-5)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)
    with pytest.raises(FileNotFoundError):
        ckpt.load_openai_clip_text(os.path.join(str(tmp_path), "missing.pt"))
