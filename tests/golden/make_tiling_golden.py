"""Generates tests/golden/tiling.json by running the REFERENCE's own tile_images / untile_images
(/root/reference/controller/utils/misc.py:6-47) and its half-resolution transform
(/root/reference/controller/agent/diffusion_agent.py:55-62) on seeded inputs, in the build container.
The GPU box has no /root/reference; only the committed .json travels: inputs are regenerated from numpy
RandomState(0) (bit-stable across numpy versions) and the outputs are pinned by SHA-256 plus a strided sample.

    python tests/golden/make_tiling_golden.py
"""
import hashlib
import importlib.util
import json
import os

import numpy as np
from PIL import Image

REF = "/root/reference/controller/utils/misc.py"
spec = importlib.util.spec_from_file_location("ref_misc", REF)
ref_misc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_misc)


def half_resolution_transform(resolution=512):
    from torchvision import transforms

    return transforms.Compose([
        transforms.Resize(resolution // 2, interpolation=transforms.InterpolationMode.BILINEAR),
        transforms.CenterCrop(resolution // 2),
    ])


def main():
    rng = np.random.RandomState(0)
    cameras = ["wrist", "front", "right_shoulder", "left_shoulder"]
    num_frames = 2
    views = rng.randint(0, 256, size=(4, num_frames, 256, 256, 3), dtype=np.uint8)   # [camera, t, H, W, 3]
    rgbs = [Image.fromarray(views[c, t]) for c in range(4) for t in range(num_frames)]  # camera-major, eval_genima.py:167-173
    tiles = ref_misc.tile_images(rgbs, num_frames)
    tile_arr = np.stack([np.asarray(t) for t in tiles])                               # [T, 512, 512, 3]
    gen = rng.randint(0, 256, size=(num_frames, 512, 512, 3), dtype=np.uint8)
    un = ref_misc.untile_images([Image.fromarray(g) for g in gen], cameras, half_resolution_transform())
    def pin(a):
        a = np.ascontiguousarray(a)
        return {"shape": list(a.shape), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
                "sample_stride_4099": a.reshape(-1)[::4099][:64].tolist()}

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiling.json")
    with open(out, "w") as f:
        json.dump({"generator": "numpy RandomState(0): views randint(0,256,(4,2,256,256,3)) then gen randint(0,256,(2,512,512,3))",
                   "cameras": cameras, "num_frames": num_frames, "views": pin(views), "tiles": pin(tile_arr),
                   "gen": pin(gen), "untiled": {c: pin(un[c]) for c in cameras}}, f, indent=1)
    print("wrote", out, {c: un[c].shape for c in cameras})


if __name__ == "__main__":
    main()
