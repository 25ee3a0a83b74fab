#!/bin/bash
# refill thresholds on the Voronoi grid (C4 at a reduced size)
for envs in "A=1" "SKG_REFILL=8 SKG_PEEL_REFILL=8" "SKG_REFILL=20 SKG_PEEL_REFILL=16" "SKG_REFILL=24 SKG_PEEL_REFILL=24" "SKG_REFILL=28 SKG_PEEL_REFILL=28"; do
  env $envs python bench.py --config C4 --particles 50000 --nlambda 10 --packages 1e5 --steps 1 --warmup 1 --skip-traversal --skip-cpu --skip-atomics --e2e-steps 1 > gpurun_out/c4s.json 2> gpurun_out/c4s.err || { tail -3 gpurun_out/c4s.err; continue; }
  python - "$envs" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/c4s.json").read().strip().splitlines()[-1]); s = d["stage_ms_per_step"]
print(f"{sys.argv[1]:40s} pk/s {d['value']:.3e} stages {[round(v,1) for v in s.values()]}")
PY
done
