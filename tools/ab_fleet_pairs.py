#!/usr/bin/env python
"""A/B of the packed (two rollouts per thread) and the one-per-thread fleet kernel over the fleet size."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _event_times  # noqa: E402
from cartpolesimulation_b200.fleet import Fleet, make_experiments  # noqa: E402

for E in (4, 16, 64, 256, 1024):
    row = []
    for no_pairs in (True, False):
        fl = Fleet(E, 2000, 50, noise="philox", seed=1, device=0, no_pairs=no_pairs)
        s0, tp, te = make_experiments(E, 4)
        fl.reset(s0)
        row.append(float(np.median(_event_times(lambda: fl.run(1), 20))) * 1e3)
        fl.close()
    print(f"E={E}: fleet period one/two per thread {row[0]:.1f} / {row[1]:.1f} us")
