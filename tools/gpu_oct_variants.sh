#!/bin/bash
# tree-grid throughput for every library variant on the command line ("default" = the product library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset SKG_LIBRARY; else export SKG_LIBRARY=$PWD/skirt_b200/variants/libskirtgpu_$v.so; fi
  python tools/gpu_other_grids.py 2>&1 | tee gpurun_out/og_$v.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('$v', d['grid'], 'fill_ms %.3f' % d['fill_ms'], 'steps/s %.3e' % d['steps_per_s'], 'GB/s %.0f' % d['gbs'], 'pk/s %.3e' % d['packets_per_s'], d['stage_ms'])
"
done
