#!/bin/bash
for pad in 0 8000 20000 36000; do
echo "pad=$pad"
SKG_FILL_SMEM_PAD=$pad python bench.py --steps 1 --warmup 1 --packages 20000 --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['traversal_roofline']; print(t['ms'], t['ms_count_pass'], t['frac'])"
done
