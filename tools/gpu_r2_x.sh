#!/bin/bash
# host_chunked chunk size against the e2e of the bwt legs (RCZ_HOST_CHUNK_BYTES overrides every op's default)
mkdir -p gpurun_out
for hc in 50331648 67108864 100663296; do
RCZ_HOST_CHUNK_BYTES=$hc timeout 600 python bench.py --codecs bwt --steps 5 > gpurun_out/r2_bench_x.json 2> gpurun_out/r2_bench_x.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_x.json").read().strip().splitlines()[-1])
print("chunk $hc", " | ".join("%s e2e %.2f GB/s %.1f ms" % (k, v["e2e"]["value"], v["e2e"]["ms_per_step"]) for k, v in d["per_codec"].items()))
PY
done
