#!/bin/bash
# what ray coherence is worth to the walkers: bench.py's traversal leg with the rays in random order and in Morton order
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" --skip-cpu --steps 1 --warmup 1 > gpurun_out/sort_$name.json 2> gpurun_out/sort_$name.err
python - gpurun_out/sort_$name.json $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); t = d["traversal_roofline"]
    print(f"{sys.argv[2]:16s} cells {d['config'].get('cells')} fill {t['ms']:.3f} ms frac {t['frac']:.3f} one-pass {t['through_api']['one_pass_ms']:.3f} ms  steps/ray {t['packet_steps']/t['rays']:.1f}")
except Exception as ex:
    print(sys.argv[2], "no line:", ex)
PY
}
for S in "" 1; do
export SKG_BENCH_SORT_RAYS=$S
run C2_sort$S --config C2 --packages 2e5 --rays 2097152
run C3_sort$S --config C3 --maxlevel 8 --packages 1e6 --rays 2097152
run C4_sort$S --config C4 --particles 200000 --nlambda 10 --packages 1e5 --rays 1048576
run C5_sort$S --config C5 --depth 5 --nlambda 10 --packages 1e5 --rays 1048576
done
