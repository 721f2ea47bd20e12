#!/bin/bash
# Planner validation (one GPU): parity tests, timing, ncu --set full of plan_kernel (CEM, K = 2000) and of the
# multi-block selection (cem_select_kernel, K = 65536).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plan.py -x -q 2>&1 | tail -5
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:plan_kernel -s 4 -c 1 -o gpurun_out/plan_K2000 python tools/bench_plan.py --K 2000 --iters 3 > gpurun_out/ncu_plan.log 2>&1
timeout 600 $NCU -k regex:cem_select_kernel -s 4 -c 1 -o gpurun_out/select_K65536 python tools/bench_plan.py --K 65536 --T 100 --iters 3 > gpurun_out/ncu_select.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/plan_launches.csv python tools/bench_plan.py --K 65536 --T 100 --iters 3 > /dev/null 2>&1
python tools/bench_plan.py --K 200 --T 35 --best_k 40 > gpurun_out/plan_timing.txt 2>&1
python tools/bench_plan.py --K 2000 --T 50 >> gpurun_out/plan_timing.txt 2>&1
python tools/bench_plan.py --K 16384 --T 50 >> gpurun_out/plan_timing.txt 2>&1
python tools/bench_plan.py --K 65536 --T 100 >> gpurun_out/plan_timing.txt 2>&1
cat gpurun_out/plan_timing.txt
