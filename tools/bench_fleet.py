#!/usr/bin/env python
"""Timing of the closed-loop fleet (cps_fleet_run: E experiments x (MPPI solve + plant period) per launch) -- the ncu
target for fleet_kernel (BASELINE.json configs[4])."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--E", type=int, default=1024)
    ap.add_argument("--K", type=int, default=2000)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--periods", type=int, default=10)
    args = ap.parse_args()
    import bench
    print(json.dumps(bench.fleet_bench(0, E_total=args.E, periods=args.periods, K=args.K, T=args.T)))


if __name__ == "__main__":
    main()
