#!/bin/bash
# round 2 (N GPUs, default 8): the LZ4 leg alone under torchrun — ncclAllGather after the decode vs the fused (TMA bulk store) gather
N=${1:-8}
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --codecs lz4 > gpurun_out/r2_bench_lz4_n$N.json 2> gpurun_out/r2_bench_lz4_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_lz4_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"]); print(json.dumps(d.get("gather"), indent=1))
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_lz4_n$N.err | tail -15
