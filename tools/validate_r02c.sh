#!/bin/bash
# Round-2 validation pass C (one GPU): every GPU parity test with the measured-error record, smoke(), the bench line and the
# reference arm.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","gpu_launches")}, l["roofline"]["kernel"], l["roofline"]["frac"])
m=l["mppi_solve"]
print("neural", {k:m["neural_GRU_2x64"][k] for k in ("latency_ms_median","kernel_ms_median","kernel_ms_in_stream","api")})
print("ODE_v0", m["ODE_v0"]); print("fwd", json.dumps(m["forward_optimizers"])[:1500])
PY
