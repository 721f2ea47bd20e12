#!/bin/bash
# bisect of the last MPPI loop changes: every tools/ab/*.so against the in-tree library, three repeats of the latency cases
set -u
for so in tools/ab/libcps_b200_*.so ""; do
  echo "== ${so:-in-tree}"
  for rep in 1 2; do
  for cfg in "--K 2000 --T 50 --integrator ODE_v0" "--K 2000 --T 50 --integrator ODE" "--K 65536 --T 100 --integrator ODE --cost quadratic_boundary"; do
    CPS_B200_LIB=${so:+$PWD/$so} timeout 300 python tools/bench_mppi.py --iters 200 $cfg 2>&1 | tail -1 | sed -E 's/single launch ([0-9.]+) us.*\), ([0-9.]+) us per solve.*/single \1 stream \2/'
  done; done
done
