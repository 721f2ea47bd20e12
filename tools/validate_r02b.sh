#!/bin/bash
# Round-2 validation pass B (one GPU): all GPU parity tests, then timings of the network kernels (FP32 / tensor-core).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
{
for kern in fp32 tensor; do for K in 2000 8192 65536; do
  echo "== --kernel $kern --K $K"; timeout 300 python tools/bench_net.py --kernel $kern --K $K --iters 30
done; done
} > gpurun_out/bench_net.txt 2>&1; cat gpurun_out/bench_net.txt
timeout 300 python tools/bench_mppi.py > gpurun_out/bench_mppi.txt 2>&1; tail -20 gpurun_out/bench_mppi.txt
