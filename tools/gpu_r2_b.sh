#!/bin/bash
# round 2, call B: full GPU parity suite (incl. the C5 pipeline at size) + the judged bench line with per_codec legs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/r2b_gputests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 6000 gpurun_out/r2b_bench.json; tail -20 gpurun_out/r2b_bench.err
