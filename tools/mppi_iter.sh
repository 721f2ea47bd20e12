#!/bin/bash
# One iteration on the ODE MPPI solve kernels (one GPU): parity tests, then timings at configs[0] and configs[3].
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py tests/test_gpu_fleet.py tests/test_gpu_net.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python tools/bench_mppi.py --K 2000 --T 50 --integrator ODE_v0
python tools/bench_mppi.py --K 2000 --T 50 --integrator ODE
python tools/bench_mppi.py --K 65536 --T 100 --cost quadratic_boundary
python tools/bench_mppi.py --K 65536 --T 100 --cost quadratic_boundary_grad_minimal
