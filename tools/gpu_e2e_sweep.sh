#!/bin/bash
for cb in 33554432 50331648 67108864 100663296; do
echo "chunk=$cb"
RCZ_LZ4_CHUNK_BYTES=$cb timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l); continue
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
"
done
