#!/bin/bash
# octree traversal throughput against the two tunables
for lat in 32 64 128; do for rf in 4 8 16 24; do
  SKG_TREE_LATTICE=$lat SKG_PATH_REFILL=$rf python tools/gpu_other_grids.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('lattice $lat refill $rf', d['grid'], 'fill_ms %.3f' % d['fill_ms'], 'steps/s %.3e' % d['steps_per_s'], 'pk/s %.3e' % d['packets_per_s'])
"
done; done
