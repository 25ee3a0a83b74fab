#!/bin/bash
for cv in -1 100 80 70; do
echo "carveout=$cv"
SKG_FILL_CARVEOUT=$cv python bench.py --steps 1 --warmup 1 --packages 20000 --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['traversal_roofline']; print(t['ms'], t['ms_count_pass'], t['frac'])"
done
