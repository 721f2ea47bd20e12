#!/bin/bash
# Round-2 pass F (one GPU): forward-planner A/B (previous build vs in-tree), full GPU suite, smoke, bench line, reference arm,
# compute-sanitizer over the reworked rollout / MPPI kernels.
set -u
mkdir -p gpurun_out
for so in tools/ab/libcps_b200_prev.so ""; do
  echo "== ${so:-in-tree}"
  CPS_B200_LIB=${so:+$PWD/$so} timeout 300 python tools/bench_plan.py --K 2000 --iters 100 2>&1 | tail -1
done
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
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
print("fwd", json.dumps(m["forward_optimizers"])[:1200])
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in mppi rollout fleet; do
    log=gpurun_out/sanitizer_${tool}_${part}.log
    timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > $log 2>&1
    echo "$tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver finished' $log | tr '\n' ' ')"
  done
done
