#!/bin/bash
# Builds libcps_b200.so with the tensor-core kernel's pipeline trace enabled (-DCPS_TC_TRACE); rebuild with `make` afterwards.
set -e
cd "$(dirname "$0")/../../cartpolesimulation_b200/csrc"
make -j8 > /dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DCPS_TC_TRACE -c -o cps_net_tc.o cps_net_tc.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcps_b200.so *.o
touch cps_net_tc.cu
