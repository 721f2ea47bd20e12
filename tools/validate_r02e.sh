#!/bin/bash
# Round-2 pass E (one GPU): packed-instruction form probe, full GPU suite, bench line, ncu of the reworked rollout_pair_kernel.
set -u
mkdir -p gpurun_out
./tools/ffma2_forms > gpurun_out/ffma2_forms.txt 2>&1; cat gpurun_out/ffma2_forms.txt
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_1gpu.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","gpu_launches")}, l["roofline"]["kernel"], l["roofline"]["frac"], l["roofline"]["peak"], l["roofline"]["frac_of_theoretical"])
print("e2e", l["e2e"]["value"], l["e2e"]["final_states_only"]["value"])
m=l["mppi_solve"]
print("neural", {k:m["neural_GRU_2x64"][k] for k in ("latency_ms_median","kernel_ms_median","kernel_ms_in_stream","api")})
print("ODE_v0", m["ODE_v0"]); print("keys", list(m.keys()))
for k in m:
    if k not in ("neural_GRU_2x64","ODE_v0","forward_optimizers"): print(k, json.dumps(m[k])[:400])
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:rollout_pair_kernel -s 3 -c 1 -o gpurun_out/r02_rollout_pair python tools/bench_rollout.py --iters 3 --no-pairs-skip > gpurun_out/ncu_rollout_pair.log 2>&1
tail -1 gpurun_out/ncu_rollout_pair.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-mppi > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
