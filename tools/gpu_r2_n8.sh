#!/bin/bash
# round 2 (8 GPUs): the judged bench line under torchrun at N = 8 — sharded per_codec legs, ncclAllGather vs the fused gather
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 2500 gpurun_out/r2_bench_n8.json; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n8.err | tail -15
