#!/bin/bash
# Round-2 pass D (one GPU): A/B of the rollout_pair_kernel rework (old build / 8 blocks per SM / 9 blocks per SM = in-tree),
# the full GPU suite, ncu --set full of the new kernel.
set -u
mkdir -p gpurun_out
for v in old mb8; do
  echo "== $v"; CPS_B200_LIB=$PWD/tools/ab/libcps_b200_$v.so timeout 300 python tools/bench_rollout.py --iters 30 2>&1 | tail -2
done
echo "== in-tree"; timeout 300 python tools/bench_rollout.py --iters 30 2>&1 | tail -2
timeout 300 python tools/bench_rollout.py --iters 30 --integrator ODE 2>&1 | tail -2
timeout 300 python tools/bench_rollout.py --iters 30 --no-traj 2>&1 | tail -2
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:rollout_pair_kernel -s 3 -c 1 -o gpurun_out/r02_rollout_pair python tools/bench_rollout.py --iters 3 --no-pairs-skip > gpurun_out/ncu_rollout_pair.log 2>&1
tail -2 gpurun_out/ncu_rollout_pair.log
