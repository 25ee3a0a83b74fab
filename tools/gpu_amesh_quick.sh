#!/bin/bash
# adaptive-mesh parity / Monte Carlo tests, then the C5 line at a reduced packet budget (full-size mesh) with the traversal leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_mc_gpu.py tests/test_dust_gpu.py -m gpu -x -q -k "amesh or adaptive or C5" 2>&1 | tail -4
python bench.py --config C5 --packages ${1:-5e6} --skip-cpu --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/aq_C5.json 2> gpurun_out/aq_C5.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/aq_C5.json").read().strip().splitlines()[-1]); s = d["stage_ms_per_step"]; t = d["traversal_roofline"]
print(f"C5 cells {d['config'].get('cells')} pk/s {d['value']:.4e} stages {[round(v, 1) for v in s.values()]}")
print(f"traversal rays {t['rays']} steps {t['packet_steps']} fill {t['ms']:.3f} ms count {t['ms_count_pass']:.3f} ms frac {t['frac']:.4f} through_api {t['through_api_frac']:.4f}")
PY
