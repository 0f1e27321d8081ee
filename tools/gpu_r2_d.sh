#!/bin/bash
# round 2, call D: inverse BWT with dense sampling / thread-per-chain compaction / single-match scatter; dc decode v3; C5 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dc_kernels.py tests/test_bwt_decode_kernel.py tests/test_pipeline.py -m gpu -x -q -k "not 4mib_text" 2>&1 | tail -4
for slog in 3 4 5 6; do for pf in 0 2; do
  echo "slog=$slog prefetch=$pf"
  RCZ_IBWT_SLOG=$slog RCZ_IBWT_PREFETCH=$pf timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_random"
done; done 2>&1 | tee gpurun_out/r2d_ibwt_slog.txt
for ctas in 2 6 8; do echo "slog=4 walk_ctas=$ctas"; RCZ_IBWT_WALK_CTAS=$ctas timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_random"; done 2>&1 | tee -a gpurun_out/r2d_ibwt_slog.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ibwt_ -c 14 --csv --log-file gpurun_out/r2d_ibwt_launches.csv python tools/opbench.py ibwt --blocks 64 --reps 1 > /dev/null 2>&1
timeout 300 python tools/opbench.py dc --blocks 64 --reps 3 2>&1 | grep "dc_" | tee gpurun_out/r2d_dc.txt
timeout 600 python bench.py --steps 5 --warmup 3 --codecs lz4,bwt,pipeline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 600 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
