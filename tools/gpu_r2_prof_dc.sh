#!/bin/bash
# round 2: ncu --set full capture of dc_decode_kernel (2 hexdump-text blocks of 4 MiB)
mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none -f"
timeout 900 $N -k regex:dc_decode_kernel -c 1 -o gpurun_out/r2_prof_dc_decode2 python tools/opbench.py ibwt dc --blocks 2 --reps 1 > gpurun_out/r2_prof_dc.log 2>&1
tail -3 gpurun_out/r2_prof_dc.log; ls -la gpurun_out/r2_prof_dc_decode2.ncu-rep
