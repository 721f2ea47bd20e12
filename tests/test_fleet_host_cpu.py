"""CPU tests of the fleet's host-side table preparation (cartpolesimulation_b200/fleet.py) against recordings of
the reference's data-generator pieces (tests/golden/datagen_host.npz, plant_flip.npz; oracle/gen_golden_plant.py)."""
import numpy as np

from cartpolesimulation_b200 import fleet as F
from oracle import oracle as O
from tests.parity import load_golden


def test_initial_state_and_trace_match_reference():
    z, meta = load_golden("datagen_host")
    for seed in range(meta["n"]):
        rng = np.random.default_rng([5, seed])
        length = 3.6 if seed < 4 else 0.9
        cfg = F.DataGenConfig(length_of_experiment=length, track_relative_complexity=2.0,
                              turning_points_period="regular" if seed % 2 == 0 else "random")
        s0 = F.random_initial_state(rng, cfg)
        np.testing.assert_array_equal(s0, z[f"s0_{seed}"])
        end_at = 1.0 * 0.198 * rng.uniform(-1.0, 1.0)
        f = F.random_trace_function(rng, cfg, ("previous", "0-derivative-smooth", "linear")[seed % 3], float(s0[4]), end_at)
        np.testing.assert_array_equal(np.asarray(f(np.minimum(z["times"], length))), z[f"tp_{seed}"])


def test_control_times_accumulate_like_the_plant():
    z, _ = load_golden("plant_hanging")
    cfg = F.DataGenConfig()
    np.testing.assert_array_equal(F.control_times(61, cfg), z["ctrl_time"])


def test_target_equilibrium_schedule_matches_plant():
    z, _ = load_golden("plant_flip")
    cfg = F.DataGenConfig(keep_target_equilibrium_x_seconds_up=0.1, keep_target_equilibrium_x_seconds_down=0.05)
    np.testing.assert_array_equal(F.target_equilibrium_schedule(31, cfg), z["ctrl_te"])
    # and agrees with the oracle's tick-level trace
    np.testing.assert_array_equal(F.target_equilibrium_schedule(31, cfg), O.target_equilibrium_trace(300, 1.0, 0.1, 0.05)[::10])
    # shipped schedule: 10 s up, 2.5 s down
    te = F.target_equilibrium_schedule(1500, F.DataGenConfig())
    flips = np.nonzero(np.diff(te))[0]
    assert te[0] == 1 and len(flips) == 4 and abs(flips[0] * 0.02 - 10.0) < 0.05 and abs((flips[1] - flips[0]) * 0.02 - 2.5) < 0.05


def test_make_experiments_is_sharding_invariant():
    cfg = F.DataGenConfig(length_of_experiment=4.0)
    s0, tp, te = F.make_experiments(6, 50, cfg, seed=3)
    s0b, tpb, teb = F.make_experiments(3, 50, cfg, seed=3, experiment_offset=3)
    np.testing.assert_array_equal(s0[3:], s0b)
    np.testing.assert_array_equal(tp[:, 3:], tpb)
    assert s0.shape == (6, 6) and tp.shape == (50, 6) and te.shape == (50, 6)
    assert np.abs(tp).max() <= 0.198 + 1e-6 and (np.abs(s0[:, 4]) <= 0.8 * 0.198 + 1e-6).all()
    np.testing.assert_allclose(tp[0], s0[:, 4], atol=1e-7)   # start_at_target
