#!/bin/bash
# Voronoi parity / Monte Carlo tests, then the C4 line at a reduced size (200k cells, 1e8 packets)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_mc_gpu.py tests/test_dust_gpu.py -m gpu -x -q -k "voro or C4" 2>&1 | tail -4
python bench.py --config C4 --particles ${1:-200000} --packages ${2:-1e6} --skip-cpu --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/vq_C4.json 2> gpurun_out/vq_C4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/vq_C4.json").read().strip().splitlines()[-1]); s = d["stage_ms_per_step"]; t = d["traversal_roofline"]
print(f"C4 cells {d['config'].get('cells')} pk/s {d['value']:.4e} stages {[round(v, 1) for v in s.values()]}")
print(f"traversal rays {t['rays']} steps {t['packet_steps']} fill {t['ms']:.3f} ms count {t['ms_count_pass']:.3f} ms frac {t['frac']:.4f} through_api {t['through_api_frac']:.4f}")
PY
