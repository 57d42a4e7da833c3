"""numpy restatement of controller/utils/misc.py:6-47 (tile_images / untile_images) on uint8 arrays
(TEST INFRASTRUCTURE — see oracle/__init__).  PINNED: tests/golden/tiling.npz was produced by importing the reference's
own functions in this container (tests/golden/make_tiling_golden.py)."""
from __future__ import annotations

import numpy as np


def tile_views(views: np.ndarray) -> np.ndarray:
    """views [4, 256, 256, 3] u8 (camera order = cfg.env.cameras) -> [512, 512, 3]; paste positions (x, y):
    (0, 0), (256, 0), (0, 256), (256, 256) (misc.py:13-16)."""
    tile = np.zeros((512, 512, 3), dtype=np.uint8)
    tile[:256, :256] = views[0]
    tile[:256, 256:] = views[1]
    tile[256:, :256] = views[2]
    tile[256:, 256:] = views[3]
    return tile


def untile_views(tile: np.ndarray) -> np.ndarray:
    """[512, 512, 3] -> [4, 256, 256, 3]; crop boxes (left, upper, right, lower) of misc.py:25-30.  The reference's
    Resize(256, bilinear) + CenterCrop(256) on a 256x256 crop is an identity."""
    return np.stack([tile[:256, :256], tile[:256, 256:], tile[256:, :256], tile[256:, 256:]], axis=0)
