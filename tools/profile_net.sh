#!/bin/bash
# Round-1 profiling pass (run under gpurun, ONE GPU): FP32 form probe, ncu --set full of the neural kernels and of the
# large-K MPPI kernel.  Outputs land in gpurun_out/ (scratch); summaries are produced here with tools/ncu_summary.py
# and committed under profiles/.
set -u
mkdir -p gpurun_out
./tools/peak_probe > gpurun_out/peak_probe.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:net_tc_kernel -s 3 -c 1 -o gpurun_out/net_tc_K65536 python tools/bench_net.py --K 65536 --kernel tensor --iters 3 > gpurun_out/ncu_net_tc.log 2>&1
timeout 600 $NCU -k regex:net_kernel -s 3 -c 1 -o gpurun_out/net_fp32_K2000 python tools/bench_net.py --K 2000 --kernel fp32 --iters 3 > gpurun_out/ncu_net_fp32.log 2>&1
timeout 600 $NCU -k regex:mppi_kernel -s 3 -c 1 -o gpurun_out/mppi_K65536_T100 python tools/bench_mppi.py --K 65536 --T 100 --cost quadratic_boundary --iters 3 > gpurun_out/ncu_mppi_big.log 2>&1
python tools/bench_net.py --K 65536 --kernel tensor --iters 20 > gpurun_out/net_tc_timing.txt 2>&1
python tools/bench_net.py --K 2000 --iters 50 >> gpurun_out/net_tc_timing.txt 2>&1
ls -la gpurun_out
