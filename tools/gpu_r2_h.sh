#!/bin/bash
# round 2, call H: inflate with literal batches; per-kernel times of the inverse BWT with chain records
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flate_kernel.py tests/test_zlib_kernel.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/opbench.py flate zlib --blocks 256 --reps 3 2>&1 | tee gpurun_out/r2h_opbench.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ibwt_ -c 8 --csv --log-file gpurun_out/r2h_ibwt_launches.csv python tools/opbench.py ibwt --blocks 64 --reps 1 > /dev/null 2>&1
