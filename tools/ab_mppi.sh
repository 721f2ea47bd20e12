#!/bin/bash
# A/B of the MPPI solve kernels: previous build (tools/ab/libcps_b200_prev.so) against the in-tree library.
set -u
for so in tools/ab/libcps_b200_prev.so ""; do
  echo "== ${so:-in-tree}"
  for cfg in "--K 2000 --T 50 --integrator ODE_v0" "--K 2000 --T 50 --integrator ODE" "--K 65536 --T 100 --integrator ODE --cost quadratic_boundary" "--K 65536 --T 100 --integrator ODE" "--K 8192 --T 100 --integrator ODE --cost quadratic_boundary"; do
    CPS_B200_LIB=${so:+$PWD/$so} timeout 300 python tools/bench_mppi.py --iters 100 $cfg 2>&1 | tail -1
  done
done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
