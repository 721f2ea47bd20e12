"""Parity metrics shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle (GPU) tests.

Metric: element-wise relative error is meaningless at zero crossings
(angleD, positionD, angle_sin pass through 0), so trajectories are compared per state channel with the
range-relative error  max|a-b| / range  over all rollouts and time steps.  range = the larger of max|b| and the
channel's range in the reference's own normalisation table (GymlikeCartPole/Dense-7IN-32H1-32H2-1OUT-0/
NI_2024-08-17_22-23-01.csv: angleD +-18.38, position +-0.198, positionD +-1.125, angle_cos/sin +-1; angle +-pi) --
the scale SURVEY.md section 8c used when it derived the 1e-5 figure ("<= 2.6e-5 abs on angleD (<= 1.3e-6 of
range)").  Without it a batch that hangs near +-pi with |angleD| < 2 would be judged against those small numbers,
where plain fp32 rounding of the reference itself already exceeds 1e-5 (tests/test_gpu_parity.py::
test_fp32_noise_floor quantifies that).  The angle channel is compared modulo 2*pi (an fp32-vs-fp64 wrap
decision at +-pi legitimately flips it by 2*pi).
"""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHANNELS = ["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"]
RANGES = {"angle": np.pi, "angleD": 18.38, "angle_cos": 1.0, "angle_sin": 1.0, "position": 0.198, "positionD": 1.125}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def traj_err(a, b):
    """per-channel norm-wise relative error of trajectories a vs reference b ([..., 6])."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = a - b
    d[..., 0] = (d[..., 0] + np.pi) % (2 * np.pi) - np.pi
    out = {}
    for c, n in enumerate(CHANNELS):
        scale = max(RANGES[n], np.abs(b[..., c]).max())
        out[n] = np.abs(d[..., c]).max() / scale
    return out


def max_traj_err(a, b):
    return max(traj_err(a, b).values())


def vec_err(a, b):
    """norm-wise relative error of a vector quantity (costs, controls)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
