"""Test stand-ins for what the reference's eval workspace does around the two plugin points, restated from
/root/reference/controller/eval_genima.py so that the tests read like the caller:

  * `instantiate(node, *args, **kw)`      hydra.utils.instantiate on a dict config: `_target_` import, `_partial_: true`
                                           nodes become functools.partial objects (what GenimaACT.build_actor receives as
                                           `self.actor_model`, genima_act.py:224-231), keyword overrides win;
  * `load_controller_ckpt(agent, path)`   eval_genima.py:91-103, line for line;
  * `eval_mode(*models)`                  robobase.utils.eval_mode as the loop uses it (eval_genima.py:199-201) [U];
  * `FakeSpace / make_spaces`             gymnasium-like observation / action spaces of the wrapped RLBench env
                                           (FrameStack puts the time axis first: rgb (T, 3, H, W), low_dim_state (T, S)).
hydra / omegaconf / gymnasium / robobase are not installed offline.
"""
import functools
import importlib
import sys
import types

import torch

# controller/cfgs/method/genima_act.yaml:3-39 (interpolations resolved the way hydra would: action_sequence = 20)
GENIMA_ACT_YAML = {
    "_target_": "method.genima_act.GenimaACT",
    "device": "cpu",
    "is_rl": False,
    "lr": 5.0e-05,
    "lr_backbone": 1.0e-05,
    "weight_decay": 0.0001,
    "num_train_steps": 200000,
    "adaptive_lr": False,
    "actor_grad_clip": None,
    "actor_model": {
        "_target_": "method.genima_act.GenimaMVTransformer", "_partial_": True, "input_shape": "???",
        "hidden_dim": 256, "enc_layers": 4, "dec_layers": 6, "dim_feedforward": 2048, "dropout": 0.1, "nheads": 8,
        "num_queries": 20, "pre_norm": False, "state_dim": 8, "action_dim": 8, "use_lang_cond": True,
        "data_augmentation": True,
    },
    "encoder_model": {
        "_target_": "robobase.method.act.ImageEncoderACT", "_partial_": True, "input_shape": "???", "hidden_dim": 256,
        "position_embedding": "sine", "lr_backbone": 1.0e-05, "masks": False, "backbone": "resnet18", "dilation": False,
        "use_lang_cond": True,
    },
}


def install_reference_stub_modules():
    """`method.genima_act` / `robobase.method.act` exist in the reference's environment; hydra resolves the `_partial_`
    nodes' `_target_` against them.  Empty stand-in classes are enough: the B200 agent only reads `.keywords`."""
    for mod, names in (("method", ()), ("method.genima_act", ("GenimaMVTransformer",)), ("robobase", ()),
                       ("robobase.method", ()), ("robobase.method.act", ("ImageEncoderACT",))):
        if mod not in sys.modules:
            sys.modules[mod] = types.ModuleType(mod)
        for n in names:
            if not hasattr(sys.modules[mod], n):
                setattr(sys.modules[mod], n, type(n, (), {"__init__": lambda self, **kw: None}))


def _resolve(target: str):
    mod, _, name = target.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(node, *args, **overrides):
    node = dict(node)
    target = _resolve(node.pop("_target_"))
    partial = bool(node.pop("_partial_", False))
    kwargs = {}
    for k, v in node.items():
        if isinstance(v, dict) and "_target_" in v:
            v = instantiate(v)
        kwargs[k] = v
    kwargs.update(overrides)
    if partial:
        kwargs = {k: v for k, v in kwargs.items() if v != "???"}
        return functools.partial(target, *args, **kwargs)
    return target(*args, **kwargs)


class FakeSpace:
    def __init__(self, shape):
        self.shape = tuple(shape)


class FakeDictSpace(dict):
    @property
    def spaces(self):
        return self


def make_spaces(cameras, frame_stack: int, image_size: int, state_dim: int, action_dim: int):
    obs = FakeDictSpace()
    for c in cameras:
        obs[f"{c}_rgb"] = FakeSpace((frame_stack, 3, image_size, image_size))
    obs["low_dim_state"] = FakeSpace((frame_stack, state_dim))
    obs["lang_tokens"] = FakeSpace((frame_stack, 77))
    return obs, FakeSpace((action_dim,))


def load_controller_ckpt(controller_agent, checkpoint_path, device="cpu"):
    """controller/eval_genima.py:91-103."""
    checkpoint = torch.load(checkpoint_path, map_location=device, weights_only=False)
    missing_keys = [
        k for k in controller_agent.state_dict().keys() if k not in checkpoint["agent"].keys() and "clip" not in k
    ]
    if len(missing_keys) > 0:
        raise ValueError(f"Missing keys in controller checkpoint: {missing_keys}")
    return controller_agent.load_state_dict(checkpoint["agent"], strict=False)


class eval_mode:
    """robobase.utils.eval_mode [U]: remember `.training`, switch to eval, restore on exit."""

    def __init__(self, *models):
        self.models = models

    def __enter__(self):
        self.prev_states = []
        for m in self.models:
            self.prev_states.append(m.training)
            m.train(False)

    def __exit__(self, *args):
        for m, s in zip(self.models, self.prev_states):
            m.train(s)
        return False


CAMERAS = ("wrist", "front", "right_shoulder", "left_shoulder")     # controller/eval_genima.py:229-232


def write_tiny_clip_tokenizer(path):
    """A syntactically complete CLIP BPE vocabulary (byte alphabet + a few merges): no real vocabulary exists offline."""
    import json
    import os

    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    chars = [chr(c) for c in cs]
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    merges = ["t h", "th e</w>", "o p", "op e", "ope n</w>", "b o", "bo x</w>"]
    for m in merges:
        a, b = m.split()
        vocab.setdefault(a + b, len(vocab))
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "vocab.json"), "w") as f:
        json.dump(vocab, f)
    with open(os.path.join(path, "merges.txt"), "w") as f:
        f.write("#version: 0.2\n" + "\n".join(merges) + "\n")
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"model_max_length": 77, "pad_token": "<|endoftext|>", "bos_token": "<|startoftext|>",
                   "eos_token": "<|endoftext|>", "unk_token": "<|endoftext|>", "tokenizer_class": "CLIPTokenizer"}, f)
    return vocab
