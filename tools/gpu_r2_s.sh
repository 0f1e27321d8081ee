#!/bin/bash
# round 2 (late): the kernels changed after the sanitizer passes — fused gather on bulk stores (one-GPU form), DC decode — under
# memcheck / racecheck / synccheck, and a fresh ncu capture of dc_decode_kernel.  Logs -> gpurun_out/r2_sanitize3_*.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 python -m pytest tests/test_lz4_kernel.py tests/test_dc_kernels.py -m gpu -x -q 2>&1 | tail -3
run() { local name=$1 tool=$2 to=$3; shift 3
  echo "== $name ($tool)"
  timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest "$@" -m gpu -x -q > gpurun_out/r2_sanitize3_${name}_${tool}.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/r2_sanitize3_${name}_${tool}.txt | tail -4
}
run gather memcheck 600 tests/test_lz4_kernel.py -k "gather_bulk_stores"
run gather racecheck 900 tests/test_lz4_kernel.py -k "gather_bulk_stores"
run gather synccheck 600 tests/test_lz4_kernel.py -k "gather_bulk_stores"
run dc memcheck 900 tests/test_dc_kernels.py -k "test_dc_gpu"
run dc racecheck 900 tests/test_dc_kernels.py -k "test_dc_gpu_alphabet"
N="ncu --set full --import-source on --clock-control none -f"
timeout 900 $N -k regex:dc_decode_kernel -c 1 -o gpurun_out/r2_prof_dc_decode3 python tools/opbench.py ibwt dc --blocks 2 --reps 1 > gpurun_out/r2_prof_dc.log 2>&1
ls -la gpurun_out/r2_prof_dc_decode3.ncu-rep
