"""The rotation kernels' once-per-control-step sin / cos (cps_device.cuh: sincos_folded, resync_angle2) against the math
library on the device: every float of [-pi, pi], bit for bit (cps_selftest_sincos)."""
import pytest

pytestmark = pytest.mark.gpu


def test_sincos_folded_is_sincosf_on_every_float_of_the_folded_range():
    from cartpolesimulation_b200.core import Engine
    eng = Engine(32, 5, integrator="ODE_v0", cost=None, device=0)
    try:
        assert eng.selftest_sincos() == 0
    finally:
        eng.close()
