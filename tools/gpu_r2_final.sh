#!/bin/bash
# round 2, final evidence at the last commit: the whole GPU parity suite, smoke(), the judged bench line (N = 1) and the reference arm,
# the op-level table, and the launch list of the bench command.  Outputs -> gpurun_out/r2_final_*
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2_final_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_final_gputests.txt
timeout 900 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; tail -c 400 gpurun_out/r2_final_bench_reference.json
timeout 1200 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.json; tail -3 gpurun_out/r2_final_bench.err
timeout 900 python tools/opbench.py --blocks 64 --reps 3 > gpurun_out/r2_final_opbench.txt 2>&1; tail -3 gpurun_out/r2_final_opbench.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_final_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_final_launches_bench.log 2>&1; wc -l gpurun_out/r2_final_launches_bench.csv
