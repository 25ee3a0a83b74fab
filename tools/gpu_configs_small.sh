#!/bin/bash
# every BASELINE configuration through bench.py at a reduced size (debug pass): engine arm with the CPU reference beside it
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" > gpurun_out/small_$name.json 2> gpurun_out/small_$name.err; echo "== $name rc=$?"; tail -c 600 gpurun_out/small_$name.err | tail -n 4
python - gpurun_out/small_$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d["stage_ms_per_step"]; c = d.get("cpu_baseline") or {}
    print(f"   cells {d['config'].get('cells')} pk/s {d['value']:.3e} e2e {d['e2e']['value']:.3e} cpu {c.get('value')} stages {[round(v,1) for v in s.values()]} roofline {d['roofline']['kernel'][:24]} {d['roofline']['frac']:.3f}",
          "trav", round(d.get('traversal_roofline', {}).get('frac', 0), 3), d.get("selfabs_cycles_per_step"))
except Exception as ex:
    print("   no line:", ex)
PY
}
run C1 --config C1 --steps 2 --warmup 1 --rays 1048576
run C3 --config C3 --maxlevel 6 --packages 2e6 --steps 1 --warmup 1 --rays 1048576 --ref-packages 2e4
run C4 --config C4 --particles 50000 --nlambda 10 --packages 1e5 --steps 1 --warmup 1 --rays 262144 --ref-packages 2e3
run C5 --config C5 --depth 3 --nlambda 20 --packages 1e5 --steps 1 --warmup 1 --rays 262144 --ref-packages 2e3
