#!/bin/bash
# quick GPU pass: parity tests + the bench without the CPU legs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --skip-cpu "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
