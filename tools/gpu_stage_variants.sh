#!/bin/bash
# shooting-stage A/B: bench (stellar phase only, no traversal / CPU legs) for every library variant named on the command line
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset SKG_LIBRARY; else export SKG_LIBRARY=$PWD/skirt_b200/variants/libskirtgpu_$v.so; fi
  python bench.py --skip-cpu --skip-traversal --steps 2 --warmup 1 $BENCH_ARGS > gpurun_out/var_${v}.json 2> gpurun_out/var_${v}.err || { echo "$v failed"; tail -3 gpurun_out/var_${v}.err; continue; }
  python - "$v" gpurun_out/var_${v}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
s = d["stage_ms_per_step"]
print(f"{sys.argv[1]:10s} pk/s {d['value']:.3e} e2e {d['e2e']['value']:.3e} launch {s['launch_ms']:.1f} peel {s['peel_ms']:.1f} absorb {s['absorb_ms']:.1f} prop {s['propagate_ms']:.1f} kernel {s['kernel_ms']:.1f}")
PY
done
