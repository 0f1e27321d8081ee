#!/bin/bash
# round 2, call N: the judged line and the reference arm of the (near-)final commit, and the op timings table
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; tail -c 300 gpurun_out/r2n_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 300 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
timeout 900 python tools/opbench.py lz4 lz4enc ibwt bwt dc flate zlib mtf ari rle --blocks 64 --reps 3 > gpurun_out/r2n_opbench.txt 2>&1; tail -3 gpurun_out/r2n_opbench.txt
