#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (small cases): memcheck on every op, racecheck + synccheck on the kernels with
# shared-memory protocols (lz4 parse / materialise, inverse BWT partition, inflate).  Logs -> gpurun_out/sanitize_*.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name tool timeout pytest-args...
  local name=$1 tool=$2 to=$3; shift 3
  echo "== $name ($tool)"
  timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest "$@" -m gpu -x -q > gpurun_out/sanitize_${name}_${tool}.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_${name}_${tool}.txt | tail -4
}
run lz4 memcheck 600 tests/test_lz4_kernel.py -k "gpu_cases or gpu_window_cases"
run lz4enc memcheck 400 tests/test_lz4_encode_kernel.py -k "test_lz4_encode_gpu and True"
run ibwt memcheck 400 tests/test_bwt_decode_kernel.py -k "gpu_cases"
run bwt memcheck 600 tests/test_bwt_encode_kernel.py -k "gpu_cases"
run flate memcheck 600 tests/test_flate_kernel.py tests/test_zlib_kernel.py -k "test_inflate_gpu and True or test_zlib_gpu and True"
run misc memcheck 600 tests/test_ari_rle_kernels.py tests/test_dc_kernels.py tests/test_mtf_kernel.py -k "test_ari_gpu and True or test_rle_gpu and True or test_dc_gpu and True or test_mtf_gpu and True"
run pipeline memcheck 600 tests/test_pipeline.py -k "test_pipeline_gpu and True and 65536"
run lz4 racecheck 900 tests/test_lz4_kernel.py -k "gpu_window_cases and True"
run ibwt racecheck 600 tests/test_bwt_decode_kernel.py -k "gpu_cases and True"
run flate racecheck 600 tests/test_flate_kernel.py -k "test_inflate_gpu and True"
run lz4 synccheck 600 tests/test_lz4_kernel.py -k "gpu_window_cases and True"
run ibwt synccheck 400 tests/test_bwt_decode_kernel.py -k "gpu_cases and True"
