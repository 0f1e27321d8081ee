#!/bin/bash
# round 2, call A: inverse-BWT group-size / sampling sweep on the unchanged round-1 kernels
mkdir -p gpurun_out
for gs in 4194304 8388608 16777216 25165824 33554432 67108864 1073741824; do
  echo "group_syms=$gs"
  RCZ_IBWT_GROUP_SYMS=$gs timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep bwt_decode
done 2>&1 | tee gpurun_out/r2_ibwt_group_sweep.txt
for slog in 5 7 8; do
  echo "slog=$slog group=16Mi"
  RCZ_IBWT_SLOG=$slog RCZ_IBWT_GROUP_SYMS=16777216 timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep bwt_decode
done 2>&1 | tee -a gpurun_out/r2_ibwt_group_sweep.txt
