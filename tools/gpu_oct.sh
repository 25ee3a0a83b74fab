#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rf 2>&1 | tail -6
python tools/gpu_other_grids.py 2>&1 | tee gpurun_out/og_plain.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d['grid'], 'cells', d['cells'], 'fill_ms %.3f' % d['fill_ms'], 'steps/s %.3e' % d['steps_per_s'], 'GB/s %.0f' % d['gbs'], 'pk/s %.3e' % d['packets_per_s'], d['stage_ms'])
"
