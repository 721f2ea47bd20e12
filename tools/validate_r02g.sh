#!/bin/bash
# Round-2 pass G (one GPU, final): full GPU suite, smoke, bench line, reference arm, rollout kernel timings, launch list.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for extra in "" "--no-traj" "--integrator ODE"; do timeout 300 python tools/bench_rollout.py --iters 30 --no-pairs-skip $extra 2>&1 | tail -1; done
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","gpu_launches")}, l["roofline"]["kernel"], l["roofline"]["frac"], l["roofline"]["peak"])
print("e2e", l["e2e"]["value"], l["e2e"]["final_states_only"]["value"])
m=l["mppi_solve"]
print("neural", {k:m["neural_GRU_2x64"][k] for k in ("latency_ms_median","kernel_ms_median","kernel_ms_in_stream")})
print("ODE_v0", m["ODE_v0"]); print("ODE", m["ODE"]); print("K65536", m["ODE_K65536_T100"])
print("fleet", {k:m["fleet_1024x2000x50"][k] for k in ("ms_per_period","state_steps_per_s_per_gpu")})
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-mppi > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
