#!/bin/bash
# Round-1 final validation pass (one GPU): all GPU parity tests, smoke, default bench + reference arm, ncu --set full of the
# headline kernel (rollout_pair_kernel) and of fleet_kernel, launch list of a short bench run.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:rollout_pair_kernel -s 2 -c 1 -o gpurun_out/pair_v0 python tools/bench_rollout.py --iters 3 > gpurun_out/ncu_pair.log 2>&1
timeout 600 $NCU -k regex:fleet_kernel -s 3 -c 1 -o gpurun_out/fleet_E1024 python tools/bench_fleet.py --E 1024 --periods 4 > gpurun_out/ncu_fleet.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --mppi-calls 10 > gpurun_out/bench_under_ncu.log 2>&1
head -c 600 gpurun_out/bench.json; echo
ls -la gpurun_out | head -30
