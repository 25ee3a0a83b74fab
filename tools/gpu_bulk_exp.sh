#!/bin/bash
# record kernel: five 256-bit stores per group of four records (product) against one bulk copy of the async proxy per group
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" --skip-cpu --steps 1 --warmup 1 --packages 2e5 > gpurun_out/bulk_$name.json 2> gpurun_out/bulk_$name.err
python - gpurun_out/bulk_$name.json $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); t = d["traversal_roofline"]
    print(f"{sys.argv[2]:16s} rays {t['rays']} fill {t['ms']:.3f} ms frac {t['frac']:.3f} one-pass {t['through_api']['one_pass_ms']:.3f} ms frac {t['through_api_frac']:.3f}")
except Exception as ex:
    print(sys.argv[2], "no line:", ex)
PY
}
for V in "" bulk; do
  if [ -n "$V" ]; then export SKG_LIBRARY=$PWD/skirt_b200/variants/libskirtgpu_$V.so; fi
  run C2_16M_$V --config C2 --rays 16777216
  run C2_4M_$V --config C2 --rays 4194304
done
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
