#!/bin/bash
for v in 6 8 10 12 14; do
  echo "refill=$v"; SKG_REFILL=$v python bench.py --steps 1 --warmup 1 --skip-cpu --skip-traversal 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['stage_ms_per_step'])"
done
