#!/bin/bash
# usage: gpu_ncu_full.sh <kernel regex> <skip> <count>   (one full capture of the named kernels)
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --packages 200000 --rays 1048576 --skip-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-4} -c ${3:-2} -f -o gpurun_out/prof $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
