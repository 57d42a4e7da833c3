"""Regenerates tests/golden/pipelines_tiny.npz from the fp32 CPU oracle (oracle/pipeline.py) on the seeded cases of
pipeline_cases.py.  Run from the repository root:  python tests/golden/make_pipeline_golden.py
(The reference's own stack — diffusers 0.29.0, RoboBase — is not installable offline, so the vectors come from the
oracle restatement)"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import pipeline_cases as pc  # noqa: E402

if __name__ == "__main__":
    torch.manual_seed(0)
    out = {}
    for name in pc.CASES:
        out.update(pc.run_oracle(name))
    np.savez_compressed(os.path.join(HERE, "pipelines_tiny.npz"), **out)
    for k, v in out.items():
        print(f"{k:28s} {v.dtype} {tuple(v.shape)} absmax {np.abs(v.astype(np.float64)).max():.4f}")
