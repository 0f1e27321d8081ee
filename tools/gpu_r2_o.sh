#!/bin/bash
# round 2, call O: forward BWT with discarding — parity (incl. 256 text blocks of 4 MiB against the oracle) and timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_bwt_encode_kernel.py tests/test_pipeline.py tests/test_host_mirrors.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2o_gputests.txt
timeout 600 python tools/opbench.py bwt --blocks 64 --reps 3 2>&1 | grep bwt_encode | tee gpurun_out/r2o_opbench.txt
timeout 600 python tools/opbench.py bwt --blocks 256 --reps 3 2>&1 | grep bwt_encode | tee -a gpurun_out/r2o_opbench.txt
timeout 900 python bench.py --steps 5 --warmup 3 --codecs lz4,bwt,pipeline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -c 300 gpurun_out/r2o_bench.json; tail -3 gpurun_out/r2o_bench.err
