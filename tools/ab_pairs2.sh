#!/bin/bash
set -u
for K in 32768 65536 131072 262144; do for cost in quadratic_boundary quadratic_boundary_grad_minimal; do for np in "" "--no-pairs"; do
  timeout 300 python tools/bench_mppi.py --iters 100 --K $K --T 100 --integrator ODE --cost $cost $np 2>&1 | tail -1 | sed -E 's/single launch ([0-9.]+) us.*\), ([0-9.]+) us per solve.*/single \1 stream \2/'
done; done; done
