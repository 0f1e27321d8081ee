#!/bin/bash
# round 2 (last): memcheck over the kernels changed after the earlier sanitizer passes — ari (prefetching byte reader), inflate (deferred
# copy stores), the pipeline, the chunked host paths.  Logs -> gpurun_out/r2_sanitize4_*.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { local name=$1 tool=$2 to=$3; shift 3
  echo "== $name ($tool)"
  timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest "$@" -m gpu -x -q > gpurun_out/r2_sanitize4_${name}_${tool}.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/r2_sanitize4_${name}_${tool}.txt | tail -4
}
run ari memcheck 600 tests/test_ari_rle_kernels.py -k "test_ari_gpu"
run flate memcheck 900 tests/test_flate_kernel.py tests/test_zlib_kernel.py -k "test_inflate_gpu and True or test_zlib_gpu and True"
run flate racecheck 900 tests/test_flate_kernel.py -k "test_inflate_gpu and True"
run chunked memcheck 900 tests/test_host_chunked.py -k "20000 and True"
run pipeline memcheck 900 tests/test_pipeline.py -k "test_pipeline_gpu and True and 65536"
