"""Generate tests/golden/recording_reference.csv with the reference's own csv_logger (CartPole/csv_logger.py).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/gen_golden_csv.py
A small seeded history dictionary (float32 state columns, python-float time, as CartPole.dict_history holds them) is
written by the UNMODIFIED create_csv_file + save_data_to_csv_file; the input record rows are stored next to it
(recording_reference.npz).  GitPython's Repo is pointed at a throw-away object (the reference tree is not a git
checkout here); the revision line is normalised to 'n/a' afterwards.
"""
import os
import sys
import tempfile
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
from oracle import ref_loader as R  # noqa: E402

if __name__ == "__main__":
    R.load()
    import CartPole.csv_logger as CL
    CL.Repo = lambda **k: types.SimpleNamespace(head=types.SimpleNamespace(object=types.SimpleNamespace(hexsha="n/a")))
    rng = np.random.default_rng(3)
    P = 25
    rec = np.zeros((P, 16), np.float32)
    rec[:, 0] = np.cumsum(np.full(P, 0.02)) - 0.02
    rec[:, 1:14] = rng.standard_normal((P, 13)).astype(np.float32)
    rec[:, 13] = 1.0
    from cartpolesimulation_b200.recording import CSV_COLUMNS, csv_header, experiment_history
    times, tt = [], 0.0
    for i in range(P * 10):
        if i % 10 == 0:
            times.append(tt)
        tt = tt + 0.002  # CartPole.step_time (CartPole/__init__.py:326-327)
    hist = experiment_history(rec, 0.395, 0.087, 1.5e-4, times)
    assert tuple(hist.keys()) == CSV_COLUMNS
    header = csv_header(0.5, 0.002, 0.02, 0.02, "mpc", "mppi", {"L": 0.395, "m_pole": 0.087})
    with tempfile.TemporaryDirectory() as d:
        path = CL.create_csv_file("Experiment", hist.keys(), path_to_experiment_recordings=d, title="golden", header=header)
        CL.save_data_to_csv_file(path, hist, np.inf, mode="save offline")
        path2 = CL.create_csv_file("Experiment", hist.keys(), path_to_experiment_recordings=d, title="golden", header=header)
        assert os.path.basename(path2) == "Experiment-1.csv", path2
        CL.save_data_to_csv_file(path2, hist, 4, mode="save offline")
        open(os.path.join(GOLDEN, "recording_reference.csv"), "w", newline="").write(open(path, newline="").read())
        open(os.path.join(GOLDEN, "recording_reference_round4.csv"), "w", newline="").write(open(path2, newline="").read())
    np.savez_compressed(os.path.join(GOLDEN, "recording_reference.npz"), record=rec, times=np.array(times))
    print("ok")
