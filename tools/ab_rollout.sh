#!/bin/bash
# A/B of rollout_pair_kernel builds (tools/ab/libcps_b200_*.so, built by hand with -DCPS_PAIR_MIN_BLOCKS=...) against the
# in-tree library, plus the operand-form probe of the packed FP32 instructions.
set -u
mkdir -p gpurun_out
./tools/ffma2_forms 2>&1 | tee gpurun_out/ffma2_forms.txt
for so in tools/ab/libcps_b200_*.so; do
  echo "== $so"; CPS_B200_LIB=$PWD/$so timeout 300 python tools/bench_rollout.py --iters 30 --no-pairs-skip 2>&1 | tail -1
done
echo "== in-tree"; timeout 300 python tools/bench_rollout.py --iters 30 --no-pairs-skip 2>&1 | tail -1
timeout 300 python tools/bench_rollout.py --iters 30 --no-pairs-skip --integrator ODE 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sincos.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
