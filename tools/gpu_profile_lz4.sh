#!/bin/bash
# round-end evidence for the lz4 path: bench line, launch list, full ncu captures of the two big kernels
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_lz4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_lz4.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_parse -s 3 -c 1 -f -o gpurun_out/prof_lz4_parse python bench.py --steps 1 --warmup 3 > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_mat -s 3 -c 1 -f -o gpurun_out/prof_lz4_mat python bench.py --steps 1 --warmup 3 > gpurun_out/b3.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_lz4_reference.json
cat gpurun_out/bench_lz4.json
