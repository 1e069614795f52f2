"""K5 kernel-level parity: the tcgen05 implicit-GEMM convolution against the scalar validation
kernel on the shapes YOLOv8s uses (both read the same bf16 data; differences are accumulation order)."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # batch, h, w, cin, cout, k, stride, act, residual, f32-out
    (2, 16, 16, 64, 64, 1, 1, 0, 0, 0),
    (2, 16, 16, 64, 64, 3, 1, 1, 1, 0),
    (2, 16, 16, 64, 64, 1, 1, 0, 0, 1),
    (2, 32, 32, 64, 128, 3, 2, 1, 0, 0),
    (2, 16, 16, 32, 32, 3, 1, 1, 1, 0),
    (2, 32, 32, 32, 64, 3, 2, 1, 0, 0),
    (2, 16, 16, 96, 64, 1, 1, 1, 0, 0),
    (3, 20, 20, 256, 512, 3, 2, 1, 0, 0),
    (4, 20, 20, 256, 256, 3, 1, 1, 1, 0),
    (2, 12, 12, 1024, 512, 1, 1, 1, 0, 0),
    (1, 24, 24, 384, 256, 1, 1, 1, 0, 0),
    (5, 48, 48, 128, 128, 3, 1, 1, 0, 0),
    (8, 96, 96, 32, 64, 3, 2, 1, 0, 0),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_tcgen05_conv_matches_scalar_conv(case):
    from wtracker_b200._lib import lib

    d = ctypes.c_double(-1.0)
    rc = lib().wt_selftest_conv(*case, 0, ctypes.byref(d))
    assert rc == 0, lib().wt_last_error().decode()
    # outputs are O(1); bf16 has 8 mantissa bits -> one ulp at 2.0 is 0.0156
    assert 0.0 <= d.value <= 0.04, d.value
