"""The numpy/C restatement of the legacy controller_mppi_cartpole iteration (oracle/legacy.py) against recordings of the
unmodified reference controller (tests/golden/legacy_*.npz, made by oracle/gen_golden_legacy.py)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import legacy as OL

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "legacy_*.npz")))


def load(path):
    g = np.load(path)
    return g, json.loads(str(g["meta"]))


def test_fixtures_present():
    assert len(FIXTURES) == 5


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_iteration_matches_reference(path):
    g, meta = load(path)
    cfg = dict(meta["weights"], R=meta["R"], LBD=meta["LBD"], NU=meta["NU"])
    for it in range(g["s"].shape[0]):
        if it % meta["update_every"] != 0:
            # no solve on this iteration (:481): the returned control is u[0] of the shifted sequence
            assert np.array_equal(g["u_updated"][it], g["u_in"][it])
            continue
        out = OL.iteration(meta["predictor"], g["s"][it], g["u_in"][it], g["u_prev_in"][it], g["delta_u"][it],
                           target_position=meta["target_position"], cfg=cfg)
        S_ref = g["S"][it]
        # tolerance: 1e-5 of the cost range (the rollouts differ from the reference's by libm ulps)
        assert np.max(np.abs(out["S"] - S_ref) / np.maximum(np.abs(S_ref), 1.0)) < 1e-5
        assert np.max(np.abs(out["u_updated"] - g["u_updated"][it])) < 1e-4  # selected control: 1e-4 (north_star)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_sampler_stream_is_bit_exact(path):
    """Generator(SFC64(seed)): five uniforms in configure() (:353-358), one perturbation array per update iteration, one
    uniform per step for the actuator noise (:526)."""
    from numpy.random import SFC64, Generator
    g, meta = load(path)
    rng = Generator(SFC64(meta["seed"]))
    for _ in range(5):
        rng.uniform(-1.0, 1.0)
    for it in range(g["s"].shape[0]):
        if it % meta["update_every"] == 0:
            # SQRTRHODTINV is a numpy float64 scalar in the reference (:91): under NEP 50 "stdev * float32 array" is
            # float64, so iid / repeated perturbations are float64 arrays; the fixture stores them rounded to float32
            du = OL.initialize_perturbations(rng, meta["K"], meta["T"], np.float64(meta["SQRTRHODTINV"]), meta["sampling"])
            assert np.array_equal(du.astype(np.float32), g["delta_u"][it])
        Q = np.float32(g["u_updated"][it][0] * (1 + meta["p_Q"] * rng.uniform(-1.0, 1.0)))
        Q = np.clip(Q, -1.0, 1.0, dtype=np.float32)
        assert Q == g["Q"][it]


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_product_sampler_is_bit_exact(path):
    """The host-side sampler of the product class (pure numpy; no GPU needed) reproduces the reference's stream."""
    from numpy.random import SFC64, Generator
    from cartpolesimulation_b200.controller_mppi_cartpole_b200 import controller_mppi_cartpole_b200 as C
    g, meta = load(path)
    c = object.__new__(C)  # sampler only: no engine, no device
    c.num_rollouts, c.mpc_horizon = meta["K"], meta["T"]
    c.rng_mppi = Generator(SFC64(meta["seed"]))
    for _ in range(5):
        c.rng_mppi.uniform(-1.0, 1.0)
    sigma = 0.02 * (1 / np.sqrt(0.02))
    for it in range(g["s"].shape[0]):
        if it % meta["update_every"] == 0:
            du = c.initialize_perturbations(stdev=sigma, sampling_type=meta["sampling"])
            assert np.array_equal(np.asarray(du, np.float32), g["delta_u"][it])
        c.rng_mppi.uniform(-1.0, 1.0)
