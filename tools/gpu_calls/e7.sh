#!/bin/bash
mkdir -p gpurun_out
for v in "" tools/ab/libhh_b200_chains1.so tools/ab/libhh_b200_chains4.so; do
  echo "== ${v:-default (2 chains)}"
  HH_B200_LIB=$v timeout 300 python tools/prof_c4.py 64 2>&1 | grep -E "^parts|lin3"
done
