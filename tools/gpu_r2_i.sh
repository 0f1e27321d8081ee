#!/bin/bash
# round 2, call I: placement kernel with limited blocks in flight (sweep), lz4 parity after the peer-store template
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bwt_decode_kernel.py tests/test_lz4_kernel.py -m gpu -x -q 2>&1 | tail -3
for pc in 1 2 3 4 8; do echo "place_ctas=$pc"; RCZ_IBWT_PLACE_CTAS=$pc timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_random"; done 2>&1 | tee gpurun_out/r2i_place.txt
RCZ_IBWT_PLACE_CTAS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ibwt_ -c 8 --csv --log-file gpurun_out/r2i_ibwt_launches.csv python tools/opbench.py ibwt --blocks 64 --reps 1 > /dev/null 2>&1
