#!/bin/bash
# quick GPU loop for the lz4 kernels: parity tests, bench line, per-kernel launch list
set -o pipefail
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_lz4_kernel.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l); continue
    print('value', d['value'], 'ms', d['ms_per_step'], 'kern_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bq.log 2>&1
grep -E "lz4_" gpurun_out/launches_q.csv | awk -F'","' '{split($5,a,"("); print a[1], $NF}' | tail -6
