#!/bin/bash
# Cartesian path-record kernel: refill threshold sweep (bench.py traversal leg only)
for r in ${SWEEP:-12 14 16 18 20 24}; do
  SKG_PATH_REFILL=$r python bench.py --skip-cpu --steps 1 --warmup 0 --packages 20000 --rays 16777216 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['traversal_roofline']; print('refill $r fill %.3f ms count %.3f ms frac %.3f' % (t['ms'], t['ms_count_pass'], t['frac']))"
done
