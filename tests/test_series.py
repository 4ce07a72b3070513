"""The three-term economised series of the normal weight (svgf_device.cuh economised_series3, used by the packed a-trous
kernel for phiN >= 100) against pow(d, phiN) itself (reference src/Filter.cuh:407-427), in float64 on the CPU."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
from fit_series import max_weight_error  # noqa: E402


@pytest.mark.parametrize("phiN", [100.0, 128.0, 160.0, 256.0, 1024.0])
def test_weight_error_bound(phiN):
    c = phiN * 1.4426950408889634
    assert max_weight_error(phiN) <= 0.60 / c ** 3      # 1.9e-7 at phiN = 100, 8.9e-8 at the default 128
    assert max_weight_error(phiN) <= 2e-7
