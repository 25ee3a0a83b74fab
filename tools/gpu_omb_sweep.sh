#!/bin/bash
# resident CTAs per SM the stage kernels of the tree / adaptive-mesh / Voronoi grids are compiled for (register cap): 4 (product), 5, 6
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" --skip-cpu --skip-traversal --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/omb_$name.json 2> gpurun_out/omb_$name.err
python - gpurun_out/omb_$name.json $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s = d["stage_ms_per_step"]
    print(f"{sys.argv[2]:12s} cells {d['config'].get('cells')} pk/s {d['value']:.4e} stages {[round(v, 1) for v in s.values()]}")
except Exception as ex:
    print(sys.argv[2], "no line:", ex)
PY
}
for V in "" omb5 omb6; do
  if [ -n "$V" ]; then export SKG_LIBRARY=$PWD/skirt_b200/variants/libskirtgpu_$V.so; fi
  run C3_$V --config C3 --packages 1e8
  run C5_$V --config C5 --packages 2e7
  run C4_$V --config C4 --particles 200000 --packages 1e6
done
