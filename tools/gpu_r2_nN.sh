#!/bin/bash
# the judged bench line under torchrun at N GPUs (default 2), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("lz4", d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "gather", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["gather"].items() if k not in ("fused", "collective")}, {k: v for k, v in d["gather"]["fused"].items() if k != "what"})
for k, v in d["per_codec"].items():
    print(k, v.get("error") or (round(v["value"], 2), round(v["ms_per_step"], 2), "e2e", round(v["e2e"]["value"], 2), v["scaling"]))
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n$N.err | tail -5
