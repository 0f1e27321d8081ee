#!/bin/bash
# round 2, call E: inverse BWT with grid-wide head ranking / placement; lz4 host pipeline without host syncs; dc decode diet
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dc_kernels.py tests/test_bwt_decode_kernel.py tests/test_lz4_kernel.py tests/test_pipeline.py -m gpu -x -q -k "not 4mib_text" 2>&1 | tail -4
for slog in 4 5 6; do
  echo "slog=$slog"
  RCZ_IBWT_SLOG=$slog timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_"
done 2>&1 | tee gpurun_out/r2e_ibwt_slog.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ibwt_ -c 16 --csv --log-file gpurun_out/r2e_ibwt_launches.csv python tools/opbench.py ibwt --blocks 64 --reps 1 > /dev/null 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --codecs lz4,bwt,pipeline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 600 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
