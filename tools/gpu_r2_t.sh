#!/bin/bash
# round 2 (late): HOST batches in pipelined chunks (host_chunked) — parity tests, then the e2e of the bwt and flate legs with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_chunked.py tests/test_host_mirrors.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -3
for hc in default 0; do
if [ $hc = default ]; then unset RCZ_HOST_CHUNK_BYTES; else export RCZ_HOST_CHUNK_BYTES=$hc; fi
timeout 1200 python bench.py --codecs bwt,flate --steps 5 > gpurun_out/r2_bench_t.json 2> gpurun_out/r2_bench_t.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_t.json").read().strip().splitlines()[-1])
for k, v in d["per_codec"].items():
    print("chunking $hc", k, "value", round(v["value"], 2), "e2e", round(v["e2e"]["value"], 2), round(v["e2e"]["ms_per_step"], 1))
PY
tail -2 gpurun_out/r2_bench_t.err
done
