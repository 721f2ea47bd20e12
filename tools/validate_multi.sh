#!/bin/bash
# Multi-GPU validation (gpurun --gpus N): the sharded-solve parity tests, then the N-GPU bench line.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x 2>&1 | tee gpurun_out/pytest_multi_${N}gpu.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --mppi-calls 200 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}, l["config"].get("timed_region"))
print("e2e", l["e2e"])
print(json.dumps(l["mppi_sharded"], indent=1))
PY
