#!/bin/bash
# round 2: C5 pipeline leg against the size of the ari streams (RCZ_BENCH_ARI_CHUNK): stage times and container size
mkdir -p gpurun_out
for ch in 65536 32768 16384 8192; do
RCZ_BENCH_ARI_CHUNK=$ch timeout 900 python bench.py --codecs pipeline --steps 5 > gpurun_out/r2_bench_ac.json 2> gpurun_out/r2_bench_ac.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_ac.json").read().strip().splitlines()[-1])
for k, v in d["per_codec"].items():
    print($ch, k, round(v["value"], 3), round(v["ms_per_step"], 1), [round(x, 1) for x in v["roofline"].get("stage_ms")], "container", v["config"]["container_bytes_per_gpu"], "e2e", round(v["e2e"]["value"], 2))
PY
tail -2 gpurun_out/r2_bench_ac.err
done
