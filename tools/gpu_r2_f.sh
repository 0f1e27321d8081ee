#!/bin/bash
# round 2, call F: full GPU parity suite (new inflate, lz4 encode, CLI), op timings, the judged bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2f_gputests.txt
timeout 600 python tools/opbench.py flate lz4enc --blocks 256 --reps 3 2>&1 | tee gpurun_out/r2f_opbench.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:inflate_kernel -c 1 -o gpurun_out/r2f_prof_inflate -f python tools/opbench.py flate --blocks 64 --reps 1 > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
