"""Drop-in check against the UNMODIFIED reference controller (build container only; skipped where /root/reference
is absent, e.g. on the GPU box).  See tests/_dropin_probe.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/Control_Toolkit"), reason="reference tree not present")


def test_reference_controller_mpc_drives_the_plugin():
    out = subprocess.run([sys.executable, os.path.join(REPO, "tests", "_dropin_probe.py")], capture_output=True,
                         text=True, timeout=300)
    line = [l for l in out.stdout.splitlines() if l.startswith("PROBE_JSON ")]
    assert line, out.stdout[-2000:] + out.stderr[-3000:]
    r = json.loads(line[0][len("PROBE_JSON "):])
    # discovered by the reference's glob + import_module, constructed and configured by the reference's controller_mpc
    assert r["optimizer_class"] == "optimizer_mppi_b200" and r["optimizer_name"] == "mppi-b200"
    assert r["predictor_class"].startswith("SI_Toolkit.") and r["cost_class"].startswith("Control_Toolkit.")
    assert (r["num_rollouts"], r["mpc_horizon"]) == (2000, 50)
    calls = r["calls"]
    create = [c for c in calls if c[0] == "create"][0][1]
    assert create["integrator"] == "ODE" and create["substeps"] == 10 and create["cost"] == "quadratic_boundary_grad_minimal"
    assert create["num_rollouts"] == 2000 and create["horizon"] == 50 and create["interp_period"] == 10
    assert abs(create["dt"] - 0.02) < 1e-12
    # weights were read from the reference plugin's own YAML block (edited ep_weight_up = 41.5)
    cost = [c for c in calls if c[0] == "cost_params"][0][1]
    np.testing.assert_allclose(cost, [10.0, 10000.0, 41.5, 1.0, 5.0, 1.0, 0.85], rtol=1e-6)
    mp = [c for c in calls if c[0] == "mppi_params"][0][1]
    np.testing.assert_allclose(mp, [1.0, 1.0, 100.0, 1000.0, 0.03, -1.0, 1.0], rtol=1e-6)
    # variable parameters: re-read every step from the controller's VariableParameters, uploaded only on change
    var = [c[1] for c in calls if c[0] == "variable"]
    assert len(var) == 1
    np.testing.assert_allclose(var[0], [0.1, -1.0, 0.3, 0.1], rtol=1e-6)
    steps = [c[1] for c in calls if c[0] == "step"]
    assert len(steps) == 2
    assert steps[0]["u_prev"] == 0.0 and steps[1]["u_prev"] == 0.25      # last RETURNED control
    assert steps[0]["noise_shape"] == [6, 2000] and steps[0]["layout"] == 1
    np.testing.assert_allclose(steps[0]["s"][:2], [3.1, 0.1], rtol=1e-6)
    assert r["u1"] == 0.25 and r["u_type"] == "ndarray"
    resets = [c for c in calls if c[0] == "reset"]
    assert len(resets) == 2 and resets[-1][1] == 0.0                   # configure + controller_reset
