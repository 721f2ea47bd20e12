#!/bin/bash
# One iteration on the tensor-core network kernel (one GPU): its parity tests, timings, then the pipeline trace.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_plugin.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for K in 2000 8192 65536; do timeout 300 python tools/bench_net.py --kernel tensor --K $K --iters 30 | head -1; done
bash tools/tc/trace_build.sh > /dev/null 2>&1
timeout 120 python tools/bench_net.py --kernel tensor --K 2000 --iters 1 2>&1 | grep -A7 TRACE | head -8 | cut -c1-250
timeout 120 python tools/bench_net.py --kernel tensor --K 65536 --iters 1 2>&1 | grep -A7 TRACE | head -8 | cut -c1-250
