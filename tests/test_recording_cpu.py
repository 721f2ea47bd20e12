"""CSV recordings: cartpolesimulation_b200.recording against files written by the reference's own
csv_logger (tests/golden/recording_reference*.csv, made by oracle/gen_golden_csv.py), and the readers the SI_Toolkit
pipeline uses (pandas.read_csv(comment='#'))."""
import os

import numpy as np
import pandas as pd

from cartpolesimulation_b200 import recording as REC

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _inputs():
    z = np.load(os.path.join(GOLDEN, "recording_reference.npz"))
    header = REC.csv_header(0.5, 0.002, 0.02, 0.02, "mpc", "mppi", {"L": 0.395, "m_pole": 0.087})
    return z["record"], z["times"], header


def test_byte_identical_to_reference_logger(tmp_path):
    rec, times, header = _inputs()
    hist = REC.experiment_history(rec, 0.395, 0.087, 1.5e-4, times)
    assert tuple(hist.keys()) == REC.CSV_COLUMNS
    p0 = REC.write_experiment_csv(str(tmp_path), "Experiment", hist, "golden", header)
    p1 = REC.write_experiment_csv(str(tmp_path), "Experiment", hist, "golden", header, rounding_decimals=4)
    assert os.path.basename(p0) == "Experiment.csv" and os.path.basename(p1) == "Experiment-1.csv"  # never overwrite
    for mine, ref in ((p0, "recording_reference.csv"), (p1, "recording_reference_round4.csv")):
        assert open(mine, newline="").read() == open(os.path.join(GOLDEN, ref), newline="").read()


def test_pipeline_reader_round_trip(tmp_path):
    rec, times, header = _inputs()
    p = REC.write_experiment_csv(str(tmp_path), "Experiment", REC.experiment_history(rec, times=times), "t", header)
    df = pd.read_csv(p, comment="#")  # SI_Toolkit/load_and_normalize.py load_data
    assert list(df.columns) == list(REC.CSV_COLUMNS) and len(df) == rec.shape[0]
    np.testing.assert_array_equal(df["angle"].to_numpy(np.float32), rec[:, 1])
    np.testing.assert_array_equal(df["Q_ccrc"].to_numpy(np.float32)[1:], rec[:-1, 9])
    assert df["Q_ccrc"][0] == 0.0
    np.testing.assert_allclose(df["time"].to_numpy(), times, rtol=0, atol=1e-15)  # pandas' fast float parser


def test_fleet_layout_and_split(tmp_path):
    rng = np.random.default_rng(0)
    P, E = 6, 10
    r = rng.standard_normal((P, E, 16)).astype(np.float32)
    r[:, :, 0] = (np.arange(P, dtype=np.float32) * np.float32(0.02))[:, None]
    paths = REC.save_fleet_recordings(r, str(tmp_path), frac_train=0.8, frac_val=0.1)
    folders = [os.path.basename(os.path.dirname(p)) for p in paths]
    assert folders == ["Train"] * 8 + ["Validate"] + ["Test"]  # data_generator.py:301-307
    assert sorted(os.listdir(tmp_path / "Train"))[:3] == ["Experiment-1.csv", "Experiment-2.csv", "Experiment-3.csv"]
    # a second rank of a sharded fleet continues the global split
    paths2 = REC.save_fleet_recordings(r[:, :2], str(tmp_path / "b"), experiment_offset=8, number_of_experiments=10,
                                       frac_train=0.8, frac_val=0.1)
    assert [os.path.basename(os.path.dirname(p)) for p in paths2] == ["Validate", "Test"]
    df = pd.read_csv(paths[3], comment="#")
    np.testing.assert_array_equal(df["position"].to_numpy(np.float32), r[:, 3, 6])
    assert abs(df["time"][5] - 0.1) < 1e-12
