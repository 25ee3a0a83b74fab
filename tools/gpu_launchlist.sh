#!/bin/bash
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --packages 200000 --rays 1048576 --skip-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"; tail -c 600 gpurun_out/plain.log
