#!/bin/bash
set -u
for so in tools/ab/libcps_b200_prev.so ""; do
  echo "== ${so:-in-tree}"
  CPS_B200_LIB=${so:+$PWD/$so} timeout 300 python tools/bench_fleet.py 2>&1 | tail -3
done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
