#!/bin/bash
# first GPU pass of a change: parity tests, the bench (both arms), the ncu launch list and one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
SMALL="python bench.py --steps 1 --warmup 1 --packages 20000 --rays 262144 --skip-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'stellarKernel|pathFillKernel' -c 4 -o gpurun_out/prof $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -c 1500 gpurun_out/bench.json
