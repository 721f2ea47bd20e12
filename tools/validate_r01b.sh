#!/bin/bash
# Round-1 second validation pass (one GPU): all GPU parity tests, default bench + reference arm, ncu --set full of
# plan_kernel (CEM, K = 2000 and K = 65536) and of fleet_kernel in relabel mode is covered by the fleet capture; launch list.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:plan_kernel -s 4 -c 1 -o gpurun_out/plan_K2000 python tools/bench_plan.py --K 2000 --iters 3 > gpurun_out/ncu_plan.log 2>&1
timeout 600 $NCU -k regex:plan_kernel -s 4 -c 1 -o gpurun_out/plan_K65536 python tools/bench_plan.py --K 65536 --T 100 --iters 3 > gpurun_out/ncu_plan_big.log 2>&1
python tools/bench_plan.py --K 200 --T 35 --best_k 40 > gpurun_out/plan_timing.txt 2>&1
python tools/bench_plan.py --K 2000 --T 50 >> gpurun_out/plan_timing.txt 2>&1
python tools/bench_plan.py --K 65536 --T 100 >> gpurun_out/plan_timing.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --mppi-calls 10 > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/bench.json
cat gpurun_out/plan_timing.txt
ls -la gpurun_out
