import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run on the GPU box with `-m gpu`")


@pytest.fixture(scope="session")
def ops():
    """Session-wide genima_b200.ops.Ops on cuda:0 (fails loudly if the CUDA library cannot be used)."""
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test collected on a box without CUDA (run with -m 'not gpu' here)")
    from genima_b200.ops import Ops

    return Ops(0)


def report_close(name, out, ref, rtol=1e-3, atol=1e-4):
    """assert |out - ref| <= atol + rtol * |ref| element-wise (fp32 compare) with a readable failure message."""
    import torch

    o = out.detach().to("cpu", torch.float32).reshape(-1)
    r = ref.detach().to("cpu", torch.float32).reshape(-1)
    assert o.shape == r.shape, f"{name}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(o).all(), f"{name}: non-finite values in kernel output"
    err = (o - r).abs()
    tol = atol + rtol * r.abs()
    bad = err > tol
    nbad = int(bad.sum())
    worst = int(torch.argmax(err - tol))
    msg = (f"{name}: {nbad}/{o.numel()} outside rtol={rtol} atol={atol}; max abs err {float(err.max()):.3e}; "
           f"worst idx {worst}: out {float(o[worst]):.6f} ref {float(r[worst]):.6f}; ref absmax {float(r.abs().max()):.3f}")
    print(("FAIL " if nbad else "ok   ") + msg)
    assert nbad == 0, msg
