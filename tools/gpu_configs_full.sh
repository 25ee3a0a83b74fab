#!/bin/bash
# the BASELINE configurations at full size through bench.py (engine arm with the bounded CPU reference beside it)
mkdir -p gpurun_out
tag=${TAG:-full}
run() { name=$1; shift; python bench.py "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "== $name rc=$?"; tail -c 400 gpurun_out/${tag}_$name.err | tail -n 3
python - gpurun_out/${tag}_$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d["stage_ms_per_step"]; c = d.get("cpu_baseline") or {}
    print(f"   cells {d['config'].get('cells')} pk/s {d['value']:.3e} e2e {d['e2e']['value']:.3e} cpu {c.get('value')} setup {d['setup_s']:.0f}s stages {[round(v,1) for v in s.values()]} roofline {d['roofline']['kernel'][:24]} {d['roofline']['frac']:.3f}",
          "trav", round(d.get('traversal_roofline', {}).get('frac', 0), 3), d.get("selfabs_cycles_per_step"))
except Exception as ex:
    print("   no line:", ex)
PY
}
for c in "$@"; do
  case $c in
    C1) run C1 --config C1 --steps 3 --warmup 3 ;;
    C3) run C3 --config C3 --steps 1 --warmup 1 --e2e-steps 1 ;;
    C4) run C4 --config C4 --steps 1 --warmup 1 --e2e-steps 1 ;;
    C5) run C5 --config C5 --steps 1 --warmup 1 --e2e-steps 1 ;;
  esac
done
