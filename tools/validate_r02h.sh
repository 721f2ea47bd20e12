#!/bin/bash
# Round-2 pass H (one GPU, final): sanitizer over the reworked tensor-core network kernel, full GPU suite, smoke, bench line,
# reference arm, ncu capture of net_tc_kernel at config 3 (K = 2000), launch list of the bench.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in net grad; do
    log=gpurun_out/sanitizer_${tool}_${part}.log
    timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > $log 2>&1
    echo "$tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver finished' $log | tr '\n' ' ')"
  done
done
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
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
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
for K in 2000 4096 6000 8192 65536; do timeout 200 python tools/bench_net.py --K $K --kernel tensor 2>&1 | tail -2; done > gpurun_out/net_tc_timing.txt 2>&1
for hk in "32 tensor" "32 fp32" "48 tensor"; do set -- $hk; timeout 200 python tools/bench_net.py --K 2000 --hidden $1 --kernel $2 2>&1 | tail -2; done >> gpurun_out/net_tc_timing.txt 2>&1
timeout 300 python tools/bench_rpgd.py > gpurun_out/rpgd_timing.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:net_tc_kernel -s 3 -c 1 -o gpurun_out/r02_net_tc_K2000_v2 python tools/bench_net.py --K 2000 --kernel tensor --iters 3 > gpurun_out/ncu_net_tc.log 2>&1; echo "ncu net rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-mppi > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
