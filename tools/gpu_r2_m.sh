#!/bin/bash
# round 2, call M: forward BWT with the single-match partition, dc decode straight-line loop; parity suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2m_gputests.txt
timeout 600 python tools/opbench.py bwt dc --blocks 64 --reps 3 2>&1 | tee gpurun_out/r2m_opbench.txt
timeout 900 ncu --set full --import-source on --clock-control none -f -k regex:^scatter_kernel -s 8 -c 1 -o gpurun_out/r2_prof_bwt_encode_scatter python tools/opbench.py bwt --blocks 16 --reps 1 > /dev/null 2>&1
ls -la gpurun_out/r2_prof_bwt_encode_scatter.ncu-rep
