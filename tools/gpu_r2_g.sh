#!/bin/bash
# round 2, call G (2 GPUs): the judged bench under torchrun at N = 2 (per_codec legs sharded, gather legs), after a quick iBWT parity check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bwt_decode_kernel.py tests/test_pipeline.py -m gpu -x -q -k "not 4mib_text" 2>&1 | tail -3
timeout 300 python tools/opbench.py ibwt --blocks 64 --reps 5 2>&1 | grep "bwt_decode_" | tee gpurun_out/r2g_ibwt.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -c 1500 gpurun_out/r2g_bench_n2.json; tail -15 gpurun_out/r2g_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --codecs lz4 2>&1 | tail -c 600
