"""numpy restatement of controller/utils/misc.py:6-47 (tile_images / untile_images) on uint8 arrays
(TEST INFRASTRUCTURE — see oracle/__init__).  PINNED: tests/golden/tiling.npz was produced by importing the reference's
own functions in this container (tests/golden/make_tiling_golden.py)."""
from __future__ import annotations

import numpy as np


def tile_views(views: np.ndarray) -> np.ndarray:
    """views [4, 256, 256, 3] u8 (camera order = cfg.env.cameras) -> [512, 512, 3]; paste positions (x, y):
    (0, 0), (256, 0), (0, 256), (256, 256) (misc.py:13-16)."""
    s = views.shape[1]  # 256 in the reference (it asserts so); kept symbolic for the small test configurations
    tile = np.zeros((2 * s, 2 * s, 3), dtype=np.uint8)
    tile[:s, :s] = views[0]
    tile[:s, s:] = views[1]
    tile[s:, :s] = views[2]
    tile[s:, s:] = views[3]
    return tile


def untile_views(tile: np.ndarray) -> np.ndarray:
    """[512, 512, 3] -> [4, 256, 256, 3]; crop boxes (left, upper, right, lower) of misc.py:25-30.  The reference's
    Resize(256, bilinear) + CenterCrop(256) on a 256x256 crop is an identity."""
    s = tile.shape[0] // 2
    return np.stack([tile[:s, :s], tile[:s, s:], tile[s:, :s], tile[s:, s:]], axis=0)
