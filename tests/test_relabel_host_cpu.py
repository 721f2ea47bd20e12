"""CPU tests of the host side of offline relabelling (cartpolesimulation_b200/relabel.py): the reference's
attribute-name conventions and the row/evaluation expansion handed to the device, with the device stage recorded."""
import numpy as np
import pandas as pd
import pytest

from cartpolesimulation_b200 import relabel as RL
from tests.parity import load_golden

STATE = ["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"]


class Recorder:
    def __init__(self, E):
        self.E, self.calls, self.resets = E, [], 0

    def reset(self, period=0):
        self.resets += 1

    def relabel(self, states, tp, te, L, mp, noise=None):
        self.calls.append(dict(states=states, tp=tp, te=te, L=L, mp=mp))
        return np.tile(np.arange(states.shape[0], dtype=np.float32)[:, None], (1, self.E))


def test_expansion_matches_what_the_reference_feeds_its_controller():
    z, m = load_golden("relabel_plain_ode")
    E = m["files"]
    dfs = [pd.DataFrame(z[f"f{f}__table"], columns=m["columns"]) for f in range(E)]
    rec = Recorder(E)
    cfg = dict(state_components=STATE, environment_attributes_dict=m["environment_attributes_dict"])
    out = RL.add_control_along_trajectories(dfs, cfg, "Q_calculated_offline", relabeller=rec)
    c = rec.calls[0]
    assert rec.resets == 1 and c["mp"] is None
    for f in range(E):   # exactly the recorded controller.step inputs of the reference run
        np.testing.assert_array_equal(c["states"][:, f], z[f"f{f}__s"])
        np.testing.assert_array_equal(c["tp"][:, f], z[f"f{f}__tp"].astype(np.float32))
        np.testing.assert_array_equal(c["te"][:, f], z[f"f{f}__te"].astype(np.float32))
        np.testing.assert_array_equal(c["L"][:, f], z[f"f{f}__L"].astype(np.float32))
        np.testing.assert_array_equal(out[f]["Q_calculated_offline"].to_numpy(), np.arange(m["rows"]))


def test_integration_expands_rows_and_averages():
    z, m = load_golden("relabel_integrate_ode")
    dfs = [pd.DataFrame(z[f"f{f}__table"], columns=m["columns"]) for f in range(2)]
    dfs[1] = dfs[1].iloc[:3].copy()
    rec = Recorder(2)
    cfg = dict(state_components=STATE, environment_attributes_dict=m["environment_attributes_dict"])
    out = RL.add_control_along_trajectories(dfs, cfg, "Q", integration_num_evals=6, relabeller=rec, seed=1,
                                            save_output_only=True)
    c = rec.calls[0]
    ev = 8   # 6 is rounded up to the next power of two, as the reference does for Sobol sequences (:326-331)
    assert c["states"].shape == (5 * ev, 2, 6)
    s = c["states"].reshape(5, ev, 2, 6)
    assert (s == s[:, :1]).all()                      # the same recorded state for every evaluation of a row
    L = c["L"].reshape(5, ev, 2)
    assert (L >= 0.25).all() and (L <= 0.55).all() and np.unique(L[:, :, 0]).size == 5 * ev
    assert abs(L[:, :, 0].mean() - 0.4) < 0.02        # low-discrepancy sample of [0.25, 0.55]
    np.testing.assert_array_equal(s[3:, :, 1], np.broadcast_to(s[2:3, :, 1], s[3:, :, 1].shape))  # short file idles on its last row
    assert [len(o) for o in out] == [5, 3]
    np.testing.assert_allclose(out[0]["Q"].to_numpy(), np.arange(5) * ev + (ev - 1) / 2)   # mean over the evaluations


def test_attribute_name_conventions():
    df = pd.DataFrame({"L": np.linspace(0.3, 0.5, 50)})
    d2, env = RL.process_random_sampling(df.copy(), {"L": "L_random_uniform_0.25_0.55", "x": "L"}, np.random.default_rng(0))
    assert env == {"L": "L_random_uniform", "x": "L"} and d2["L_random_uniform"].between(0.25, 0.55).all()
    d3, env = RL.process_random_sampling(df.copy(), {"L": "L_random_uniform_0.2_0.5_0.1_"}, np.random.default_rng(0))
    assert set(np.round(d3["L_random_uniform"], 6)) <= {0.2, 0.3, 0.4, 0.5}
    f, r, env = RL.get_integration_features({"L": "L_integrate_0.25_0.55_", "target_position": "target_position"})
    assert f == ["L"] and r == {"L": (0.25, 0.55)} and env["L"] == "L"
    with pytest.raises(ValueError):
        RL.add_control_along_trajectories([df.assign(**{c: 0.0 for c in STATE})],
                                          dict(environment_attributes_dict={"L": "L_integrate_0.25_0.55_"}),
                                          integration_method="simpson", relabeller=Recorder(1))


class LockstepFake:
    """Stands in for Relabeller in nquad mode: the 'controller' of file e answers u = a_e + b_e * L,
    so the quadrature answer is known and the per-file step counters show who was masked out of which launch."""

    def __init__(self, E):
        import torch
        self.E, self.K, self.n_ind, self.device = E, 4, 2, torch.device("cpu")
        self.steps = np.zeros(E, dtype=np.int64)
        self.launches, self.resets = [], 0
        self.a, self.b = np.linspace(-0.2, 0.3, E), np.linspace(1.0, -2.0, E)

    def reset(self, period=0):
        self.resets += 1

    def close(self):
        pass

    def relabel_device(self, states, target_position=None, target_equilibrium=None, pole_length=None, m_pole=None,
                       noise=None, Q_out=None, J_out=None, active=None):
        act = active.numpy()[0].astype(bool)
        self.launches.append(act.copy())
        L = pole_length.numpy()[0]
        for e in np.nonzero(act)[0]:
            Q_out[0, e] = float(self.a[e] + self.b[e] * L[e])
            self.steps[e] += 1
        return Q_out


def test_nquad_runs_the_files_in_lockstep_rounds():
    """integration_method='nquad': scipy's adaptive quadrature per file and row, evaluations of all files gathered into
    masked launches; files that have no rows left (or are done with the row) sit launches out."""
    E = 3
    rows = [4, 2, 3]
    dfs = [pd.DataFrame({**{c: np.full(n, 0.1 * (e + 1)) for c in STATE}, "time": np.arange(n) * 0.02, "L": np.full(n, 0.4),
                         "target_position": np.zeros(n)}) for e, n in enumerate(rows)]
    fake = LockstepFake(E)
    cfg = dict(state_components=STATE, environment_attributes_dict={"L": "L_integrate_0.25_0.55_",
                                                                   "target_position": "target_position"})
    out = RL.add_control_along_trajectories(dfs, cfg, "Q", integration_method="nquad", integration_num_evals=64,
                                            relabeller=fake, save_output_only=True)
    assert fake.resets == 1 and [len(o) for o in out] == rows
    for e in range(E):   # a linear integrand: the 21-point Gauss-Kronrod rule is exact, one interval per row
        np.testing.assert_allclose(out[e]["Q"].to_numpy(), fake.a[e] + fake.b[e] * 0.4, rtol=0, atol=1e-6)
        assert fake.steps[e] == 21 * rows[e]
    assert len(fake.launches) == 21 * max(rows)                      # lockstep: one launch per round, not per file
    assert [int(l.sum()) for l in fake.launches[::21]] == [3, 3, 2, 1]   # the short files drop out row by row


def test_differentiation_expansion_and_labels_match_the_reference():
    """`<col>_differentiate_`: five controller steps per row at value + {-2..2} * 0.5e-3 (what the reference fed its
    controller), and from the controls the reference got back, the same derivative / central-output labels."""
    z, m = load_golden("relabel_differentiate_ode")
    E, R = m["files"], m["rows"]
    dfs = [pd.DataFrame(z[f"f{f}__table"], columns=m["columns"]) for f in range(E)]

    class Replay(Recorder):   # answers with the controls the reference's controller returned
        def relabel(self, states, tp, te, L, mp, noise=None):
            super().relabel(states, tp, te, L, mp)
            return np.stack([z[f"f{f}__u"] for f in range(E)], axis=1).astype(np.float32)

    rec = Replay(E)
    cfg = dict(state_components=STATE, environment_attributes_dict=m["environment_attributes_dict"])
    out = RL.add_control_along_trajectories(dfs, cfg, "Q_calculated_offline", relabeller=rec)
    c = rec.calls[0]
    for f in range(E):
        np.testing.assert_array_equal(c["states"][:, f], z[f"f{f}__s"])
        np.testing.assert_array_equal(c["L"][:, f], z[f"f{f}__L"].astype(np.float32))
        np.testing.assert_array_equal(c["tp"][:, f], z[f"f{f}__tp"].astype(np.float32))
        assert list(out[f].columns[-2:]) == ["Q_calculated_offline_dL", "_calculated_offline_dL"]   # the reference's names
        np.testing.assert_allclose(out[f]["Q_calculated_offline_dL"].to_numpy(), z[f"f{f}__Q_calculated_offline_dL"],
                                   rtol=1e-5, atol=1e-4)   # float32 controls vs the reference's float64 bookkeeping
        np.testing.assert_allclose(out[f]["_calculated_offline_dL"].to_numpy(), z[f"f{f}___calculated_offline_dL"],
                                   rtol=0, atol=1e-7)
    with pytest.raises(ValueError):
        RL.add_control_along_trajectories(dfs, dict(state_components=STATE, environment_attributes_dict={
            "L": "L_differentiate_", "target_position": "target_position_integrate_-0.1_0.1_"}), relabeller=Recorder(E))
