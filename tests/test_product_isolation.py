"""The product path never touches the CPU oracle and has no CPU fallback."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "genima_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            with open(os.path.join(pkg, name)) as f:
                src = f.read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{name} imports the oracle"
            assert "torch.nn.functional" not in src and "import torch.nn" not in src, \
                f"{name} uses torch.nn: arithmetic must go through the C ABI"


def test_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from genima_b200._cabi import GenimaB200Error
    from genima_b200.ops import Ops

    with pytest.raises(GenimaB200Error):
        Ops(0)


def test_smoke_entry_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import __graft_entry__ as ge

    with pytest.raises(RuntimeError):
        ge.smoke()
