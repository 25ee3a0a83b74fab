#!/bin/bash
# DRAM traffic of every absorbStage launch of ONE full-size phase (1e8 packets), for roofline.traffic
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --skip-cpu --skip-traversal"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:absorbStage -c 48 --csv --log-file gpurun_out/absorb_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "ncu rc=$?"
