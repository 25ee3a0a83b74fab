#!/bin/bash
# bench every library variant given on the command line (names under skirt_b200/variants, or "default"), twice each
mkdir -p gpurun_out
for v in "$@"; do
  for rep in 1 2; do
    if [ "$v" = default ]; then unset SKG_LIBRARY; else export SKG_LIBRARY=$PWD/skirt_b200/variants/libskirtgpu_$v.so; fi
    python bench.py --skip-cpu --steps 2 --warmup 2 > gpurun_out/var_${v}_$rep.json 2> gpurun_out/var_${v}_$rep.err || { echo "$v failed"; tail -3 gpurun_out/var_${v}_$rep.err; }
    python - "$v" gpurun_out/var_${v}_$rep.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
s = d["stage_ms_per_step"]; t = d["traversal_roofline"]
print(f"{sys.argv[1]:10s} pk/s {d['value']:.3e}  launch {s['launch_ms']:.1f} peel {s['peel_ms']:.1f} absorb {s['absorb_ms']:.1f} prop {s['propagate_ms']:.1f} | count {t['ms_count_pass']:.3f} fill {t['ms']:.3f} ms frac {t['frac']:.3f}")
PY
  done
done
