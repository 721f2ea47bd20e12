#!/bin/bash
set -u
for so in tools/ab/libcps_b200_prev.so ""; do
  echo "== ${so:-in-tree}"
  for cfg in "--K 2000 --T 50 --integrator ODE --cost quadratic_boundary" "--K 65536 --T 100 --integrator ODE --cost quadratic_boundary" "--K 8192 --T 100 --integrator ODE --cost quadratic_boundary" "--K 2000 --T 50 --integrator ODE_v0 --cost default"; do
    CPS_B200_LIB=${so:+$PWD/$so} timeout 300 python tools/bench_mppi.py --iters 200 $cfg 2>&1 | tail -1 | sed -E 's/single launch ([0-9.]+) us.*\), ([0-9.]+) us per solve.*/single \1 stream \2/'
  done
done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
