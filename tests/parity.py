"""Parity metrics shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle (GPU) tests.

Metric: element-wise relative error is meaningless at zero crossings
(angleD, positionD, angle_sin pass through 0), so trajectories are compared per state channel with the
range-relative error  max|a-b| / range  over all rollouts and time steps, where range = max|b| for angleD,
position and positionD and the natural range for the bounded channels (pi for angle, 1 for angle_cos /
angle_sin -- otherwise a batch that stays near +-pi, where |sin| is tiny, would be judged against that tiny
number); the angle channel is compared modulo 2*pi (an fp32-vs-fp64 wrap decision at +-pi legitimately flips
it by 2*pi).
"""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHANNELS = ["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"]


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
    natural = {"angle": np.pi, "angle_cos": 1.0, "angle_sin": 1.0}
    for c, n in enumerate(CHANNELS):
        scale = natural.get(n, max(np.abs(b[..., c]).max(), 1e-30))
        out[n] = np.abs(d[..., c]).max() / scale
    return out


def max_traj_err(a, b):
    return max(traj_err(a, b).values())


def vec_err(a, b):
    """norm-wise relative error of a vector quantity (costs, controls)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
