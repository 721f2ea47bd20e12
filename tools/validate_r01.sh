#!/bin/bash
# Round-1 validation pass (one GPU): GPU parity tests, default bench, reference arm, launch list, fleet_kernel ncu.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:fleet_kernel -s 3 -c 1 -o gpurun_out/fleet_E1024 python tools/bench_fleet.py --E 1024 --periods 4 > gpurun_out/ncu_fleet.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --mppi-calls 10 > gpurun_out/bench_under_ncu.log 2>&1
python tools/bench_fleet.py --E 1024 --periods 20 > gpurun_out/fleet_timing.txt 2>&1
cat gpurun_out/bench.json
ls -la gpurun_out
