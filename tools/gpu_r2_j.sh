#!/bin/bash
# round 2, call J (2 GPUs): fused decode + gather over peer memory against ncclAllGather; the whole bench line at N = 2
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --codecs lz4 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
tail -c 3000 gpurun_out/r2j_bench_n2.json; grep -v "^\*\|OMP_NUM" gpurun_out/r2j_bench_n2.err | tail -25
