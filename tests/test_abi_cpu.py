"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/cps.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side packers agree with the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(REPO, "include", "cps.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cps_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from cartpolesimulation_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    L = C.CDLL(_lib.library_path())
    for n in names:
        assert hasattr(L, n), n
    assert _lib.lib().cps_abi_version() == 1


def test_config_struct_matches_header():
    from cartpolesimulation_b200 import _lib
    assert C.sizeof(_lib.cps_config) == 11 * 4


def test_num_inducing_points_matches_reference():
    from cartpolesimulation_b200 import _lib
    from tests.parity import load_golden
    z, meta = load_golden("interp")
    for (T, p) in meta["combos"]:
        assert _lib.lib().cps_num_inducing_points(T, p) == z[f"T{T}_p{p}__W"].shape[0]
    assert _lib.lib().cps_num_inducing_points(0, 10) == -1


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cartpolesimulation_b200 import _lib
    L = _lib.lib()
    cfg = _lib.cps_config(C.sizeof(_lib.cps_config), 0, 2000, 50, 10, 0.02, 1, 2, 0, 10, 0)
    h = C.c_void_p()
    rc = L.cps_create(C.byref(cfg), C.byref(h))
    assert rc == 2 and not h.value
    assert b"no CPU fallback" in L.cps_last_error(None)
    with pytest.raises(RuntimeError):
        from cartpolesimulation_b200.core import Engine
        Engine(16, 10)


def test_create_rejects_bad_config_before_touching_cuda():
    from cartpolesimulation_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    bad = [dict(struct_size=8), dict(num_rollouts=0), dict(horizon=0), dict(substeps=0), dict(dt=0.0),
           dict(interp_period=0), dict(integrator=7), dict(cost_id=9), dict(noise_mode=3)]
    for kw in bad:
        base = dict(struct_size=C.sizeof(_lib.cps_config), device=0, num_rollouts=8, horizon=5, substeps=10, dt=0.02,
                    integrator=1, cost_id=2, noise_mode=0, interp_period=10, flags=0)
        base.update(kw)
        cfg = _lib.cps_config(*[base[f[0]] for f in _lib.cps_config._fields_])
        rc = L.cps_create(C.byref(cfg), C.byref(h))
        assert rc in (1, 3), (kw, rc)
        assert L.cps_last_error(None)


def test_cost_packers_agree_with_oracle_and_reference():
    from cartpolesimulation_b200 import config as cfg
    from oracle import oracle as O
    from tests.parity import load_golden
    z, _ = load_golden("costs")
    for name in ("default", "quadratic_boundary"):
        assert float(cfg.max_cost(name, cfg.DEFAULT_COST_CONFIG[name])) == float(z[f"{name}__max_cost"])
        np.testing.assert_array_equal(cfg.cost_vector(name, cfg.DEFAULT_COST_CONFIG[name]), O.cost_params(name))
    np.testing.assert_array_equal(cfg.cost_vector("quadratic_boundary_grad_minimal"),
                                  O.cost_params("quadratic_boundary_grad_minimal"))
    v = cfg.cost_vector("quadratic_boundary_grad")
    up, down = O.cost_params("quadratic_boundary_grad", None, 1.0), O.cost_params("quadratic_boundary_grad", None, -1.0)
    assert len(v) == 19
    np.testing.assert_array_equal(np.r_[v[0:7], v[16:18], v[7], v[18]], up)
    np.testing.assert_array_equal(np.r_[v[8:15], v[16:18], v[15], v[18]], down)
    np.testing.assert_array_equal(cfg.physics_vector(), O.physics_vector())
    with pytest.raises(ValueError):
        cfg.physics_vector(mass=1.0)
    with pytest.raises(ValueError):
        cfg.cost_vector("quadratic_boundary_nonconvex")
