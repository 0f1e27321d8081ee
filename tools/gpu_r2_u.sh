#!/bin/bash
# round 2 (late): inflate with deferred match-copy stores — parity tests, then the flate leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flate_kernel.py tests/test_zlib_kernel.py tests/test_host_chunked.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --codecs flate > gpurun_out/r2_bench_u.json 2> gpurun_out/r2_bench_u.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_u.json").read().strip().splitlines()[-1])
for k, v in d["per_codec"].items():
    print(k, "value", round(v["value"], 2), "ms", round(v["ms_per_step"], 1), "e2e", round(v["e2e"]["value"], 2), round(v["e2e"]["ms_per_step"], 1))
PY
tail -2 gpurun_out/r2_bench_u.err
