#!/bin/bash
# full ncu captures: path kernels at the bench's ray count, and one launch each of the stage kernels mid-phase
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --packages 200000 --rays 4194304 --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"pathFillKernel|pathCountKernel" -s 2 -c 2 -f -o gpurun_out/prof_path $CMD > gpurun_out/ncu_full2.log 2>&1; echo "ncu path rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"absorbStage|peelStage" -s 8 -c 2 -f -o gpurun_out/prof_stages $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu stages rc=$?"
tail -2 gpurun_out/ncu_full.log
