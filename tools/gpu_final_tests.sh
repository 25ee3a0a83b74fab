#!/bin/bash
# final checks of the round on one GPU: the GPU test suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_final_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench.json")); r = json.load(open("gpurun_out/r02_final_bench_reference.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["s_per_step"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("stages", d["stage_ms_per_step"]); print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "atomics", d["atomics"]["frac"])
t = d["traversal_roofline"]; print("traversal", t["frac"], t["through_api_frac"], t["traffic"])
print("cpu", d["cpu_baseline"]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], "ratio e2e", d["e2e"]["value"] / r["value"])
PY
