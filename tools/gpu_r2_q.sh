#!/bin/bash
# round 2: DC decode with the register-resident list — parity tests, then the C5 pipeline leg (stage times)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dc_kernels.py tests/test_pipeline.py tests/test_ari_rle_kernels.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --codecs pipeline > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_q.json").read().strip().splitlines()[-1])
for k, v in d["per_codec"].items():
    print(k, round(v["value"], 3), round(v["ms_per_step"], 1), v["roofline"].get("stage_ms"), "e2e", round(v["e2e"]["value"], 2))
PY
tail -3 gpurun_out/r2_bench_q.err
