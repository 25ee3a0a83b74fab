#!/bin/bash
# two GPUs: the 2-rank parity test (stellar -> self-absorption cycles -> emission against the 1-rank run) and the weak / strong bench lines
mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r02_final_multigpu_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_final_multigpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-traversal > gpurun_out/r02_final_bench_2gpu_weak.json 2> gpurun_out/2w.err; echo "weak rc=$?"
$TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-traversal --scaling strong > gpurun_out/r02_final_bench_2gpu_strong.json 2> gpurun_out/2s.err; echo "strong rc=$?"
python - <<'PY'
import json
for n in ("weak", "strong"):
    try:
        d = json.load(open(f"gpurun_out/r02_final_bench_2gpu_{n}.json")); print(n, d["value"], "e2e", d["e2e"]["value"], d["e2e"]["of_device_value"], d.get("allreduce_ms"), d["config"].get("packets_per_step_all_phases"))
    except Exception as ex:
        print(n, "no line", ex)
PY
