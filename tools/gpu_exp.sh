#!/bin/bash
python tools/gpu_other_grids.py > gpurun_out/og_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"pathFillKernel" -s 2 -c 1 -f -o gpurun_out/prof_oct python tools/gpu_other_grids.py > gpurun_out/ncu_oct.log 2>&1
echo rc=$?
