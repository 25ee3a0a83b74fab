#!/bin/bash
# bench (stellar phase only) under a list of environment settings: tools/gpu_env_sweep.sh "SKG_REFILL=16" "SKG_REFILL=20 SKG_PEEL_REFILL=24" ...
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --skip-cpu --skip-traversal --steps 2 --warmup 1 $BENCH_ARGS > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err || { echo "$envs failed"; tail -3 gpurun_out/sweep_$i.err; continue; }
  python - "$envs" gpurun_out/sweep_$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
s = d["stage_ms_per_step"]
print(f"{sys.argv[1]:40s} pk/s {d['value']:.3e} e2e {d['e2e']['value']:.3e} launch {s['launch_ms']:.1f} peel {s['peel_ms']:.1f} absorb {s['absorb_ms']:.1f} prop {s['propagate_ms']:.1f} kernel {s['kernel_ms']:.1f}")
PY
done
