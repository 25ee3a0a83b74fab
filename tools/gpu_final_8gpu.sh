#!/bin/bash
# eight GPUs: C2 weak and strong, and C5 (pan flow with self-absorption cycles, the BASELINE 8-GPU configuration) with its budget split over the GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
$TR bench.py --gpus 8 --steps 5 --warmup 3 --skip-traversal > gpurun_out/r02_final_bench_8gpu_weak.json 2> gpurun_out/8w.err; echo "weak rc=$?"
$TR bench.py --gpus 8 --steps 5 --warmup 3 --skip-traversal --scaling strong > gpurun_out/r02_final_bench_8gpu_strong.json 2> gpurun_out/8s.err; echo "strong rc=$?"
$TR bench.py --gpus 8 --config C5 --steps 1 --warmup 1 --e2e-steps 1 --skip-traversal --scaling strong > gpurun_out/r02_final_bench_8gpu_C5_strong.json 2> gpurun_out/8c5.err; echo "C5 rc=$?"
python - <<'PY'
import json
for n in ("weak", "strong", "C5_strong"):
    try:
        d = json.load(open(f"gpurun_out/r02_final_bench_8gpu_{n}.json")); print(n, d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["of_device_value"], d.get("allreduce_ms"), d.get("selfabs_cycles_per_step"), d["config"].get("packets_per_step_all_phases"))
    except Exception as ex:
        print(n, "no line", ex)
PY
