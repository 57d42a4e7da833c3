"""Host-side glue with the reference's own call signatures (controller/utils/misc.py:6-47), for callers that keep the
PIL-based loop of controller/eval_genima.py:163-234 unchanged.  Pure data movement on uint8 — no arithmetic."""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np
from PIL import Image


_ORIGINS = ((0, 0), (256, 0), (0, 256), (256, 256))   # paste position (x, y) of cameras 0..3 (misc.py:13-18)


def tile_images(rgbs: Sequence[Image.Image], num_frames: int) -> List[Image.Image]:
    """rgbs: camera-major list (index = camera * num_frames + t) of four cameras' 256x256 PIL images -> per frame one
    512x512 tile, cameras 0..3 at (x, y) = (0, 0), (256, 0), (0, 256), (256, 256)."""
    if not isinstance(rgbs[0], Image.Image):
        raise AssertionError("Images must be PIL Images")
    if rgbs[0].size != (256, 256):
        raise AssertionError("For tiling, image sizes must be 256x256")
    tiles = []
    for t in range(num_frames):
        tile = np.empty((512, 512, 3), dtype=np.uint8)
        for k, (x, y) in enumerate(_ORIGINS):
            im = rgbs[k * num_frames + t]
            tile[y:y + 256, x:x + 256] = np.asarray(im if im.mode == "RGB" else im.convert("RGB"))
        tiles.append(Image.fromarray(tile))
    return tiles


def untile_images(gen_images: Sequence[Image.Image], cameras: Sequence[str],
                  resize_transform: Callable[[Image.Image], Image.Image]) -> Dict[str, np.ndarray]:
    """Per camera (quadrant order as above) a uint8 array [T, 3, 256, 256] of the resized quadrant crops."""
    if gen_images[0].size != (512, 512):
        raise AssertionError("For untiling, image sizes must be 512x512")
    boxes = [(0, 0, 256, 256), (256, 0, 512, 256), (0, 256, 256, 512), (256, 256, 512, 512)]
    if getattr(resize_transform, "is_identity_for", lambda size: False)((256, 256)):
        # Resize(256) + CenterCrop(256) of a 256 x 256 crop is the identity: one array view of the tile,
        # four strided copies straight into contiguous [T, 3, 256, 256] outputs
        T = len(gen_images)
        res = {c: np.empty((T, 3, 256, 256), dtype=np.uint8) for c in cameras}
        for t, tile in enumerate(gen_images):
            a = np.asarray(tile if tile.mode == "RGB" else tile.convert("RGB"))
            for (x0, y0, x1, y1), cam in zip(boxes, cameras):
                res[cam][t] = a[y0:y1, x0:x1].transpose(2, 0, 1)
        return res
    out: Dict[str, list] = {c: [] for c in cameras}
    for tile in gen_images:
        for box, cam in zip(boxes, cameras):
            crop = np.asarray(resize_transform(tile.crop(box)))
            out[cam].append(np.transpose(crop, (2, 0, 1))[None])
    return {c: np.concatenate(v, axis=0) for c, v in out.items()}
