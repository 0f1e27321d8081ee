#!/bin/bash
# round 2, call C: new ari / dc-decode kernels (parity + timings), inverse-BWT walk variants (L2 prefetch of the next tables), honest gather ceiling
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ari_rle_kernels.py tests/test_dc_kernels.py tests/test_bwt_decode_kernel.py tests/test_pipeline.py -m gpu -x -q -k "not 4mib_text" 2>&1 | tail -4
( cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_bench2 gather_bench2.cu && /tmp/gather_bench2 ) 2>&1 | tee gpurun_out/r2c_gather_bench2.txt
timeout 600 python tools/opbench.py ari dc ibwt --blocks 64 --reps 5 2>&1 | grep -v "bwt_encode" | tee gpurun_out/r2c_opbench.txt
for pf in 0 1 2 4; do for ctas in 2 3 4; do
  echo "prefetch=$pf walk_ctas=$ctas"
  RCZ_IBWT_PREFETCH=$pf RCZ_IBWT_WALK_CTAS=$ctas timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_random"
done; done 2>&1 | tee gpurun_out/r2c_ibwt_prefetch.txt
echo "l2fetch=32 prefetch=2"; RCZ_L2_FETCH=32 timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_random" | tee -a gpurun_out/r2c_ibwt_prefetch.txt
for pf in 0 2; do
  RCZ_IBWT_PREFETCH=$pf timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum --clock-control none -k regex:ibwt_walk -c 2 --csv --log-file gpurun_out/r2c_walk_pf$pf.csv python tools/opbench.py ibwt --blocks 64 --reps 1 > /dev/null 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 --codecs lz4,pipeline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
