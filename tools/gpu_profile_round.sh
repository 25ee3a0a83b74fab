#!/bin/bash
# evidence pass for profiles/: launch list of the small bench + one full capture of the four hot kernels
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --packages 200000 --rays 1048576 --skip-cpu"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"absorbStage|peelStage|propagateStage|pathFillKernel" -s 8 -c 4 -f -o gpurun_out/prof_stages $SMALL > gpurun_out/ncu_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"pathFillKernel|pathCountKernel" -s 0 -c 2 -f -o gpurun_out/prof_path $SMALL > gpurun_out/ncu_full2.log 2>&1
echo "ncu rc=$?"
