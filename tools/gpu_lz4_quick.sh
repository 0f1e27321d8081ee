#!/bin/bash
# quick GPU loop for the lz4 kernels: parity tests, the LZ4 leg of the bench line (stage times), other inputs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_lz4_kernel.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --codecs lz4 2>gpurun_out/lz4q.err | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l); continue
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'stages', [round(x, 3) for x in d['roofline']['stage_ms']], 'frac', round(d['roofline']['frac'], 4), 'e2e', round(d['e2e']['value'], 1))
"
tail -2 gpurun_out/lz4q.err
timeout 300 python tools/opbench.py lz4 --blocks 256 --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['op'], round(d['GBps'], 1), d.get('stage_ms'))
"
