"""Host-side logic of the fused GroupNorm statistics (no GPU): bucket choice and the layouts the statistics pass accepts."""
import pytest

from genima_b200.ops import Ops, gn_bucket_for


def test_bucket_divides_every_group_of_every_concat():
    # SD-2.1 U-Net: channels 320 / 640 / 1280, 32 groups; decoder concats 2560, 1920, 1280, 960, 640
    b = gn_bucket_for((320, 640, 1280, 1280), 32)
    assert b == 10
    for c0, c1 in [(1280, 1280), (1280, 640), (640, 640), (640, 320), (320, 320), (320, 0), (1280, 0)]:
        cpg = (c0 + c1) // 32
        assert cpg % b == 0 and c0 % b == 0 and c1 % b == 0
    # KL-VAE decoder: 128 / 256 / 512 channels
    assert gn_bucket_for((128, 256, 512, 512), 32) == 4


@pytest.mark.parametrize("channels,groups", [((48, 96), 32), ((32, 64), 32), ((40,), 8)])
def test_bucket_is_disabled_when_not_an_even_integer(channels, groups):
    b = gn_bucket_for(channels, groups)
    assert b == 0 or (b % 2 == 0 and all((c // groups) % b == 0 for c in channels))


def test_rows_per_image_rule_matches_the_kernel():
    # gemm.cu fill_out_geom: multiples of 128, or powers of two in [16, 64]
    ok = [16, 32, 64, 128, 256, 1024, 4096, 16384]
    bad = [1, 8, 20, 48, 100, 192 + 1, 258]
    assert all(Ops._gn_rows_ok(r) for r in ok)
    assert not any(Ops._gn_rows_ok(r) for r in bad)
