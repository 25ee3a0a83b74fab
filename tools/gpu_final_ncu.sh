#!/bin/bash
# one ncu pass per call.  usage: gpu_final_ncu.sh launches | stagetraffic | pathtraffic | full <kernel regex> <skip> <count> <tag>
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --packages 200000 --rays 1048576 --skip-cpu"
case $1 in
launches)
    $SMALL > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_final_launches.csv $SMALL > gpurun_out/ncu.log 2>&1
    echo "ncu rc=$?"; python tools/ncu_traffic.py gpurun_out/r02_final_launches.csv > gpurun_out/r02_final_launches_summary.json; head -c 1500 gpurun_out/r02_final_launches_summary.json ;;
stagetraffic)
    # every stage kernel launch of ONE full-size C2 step (1e8 packets): DRAM bytes and durations, caches left as the run leaves them
    CMD="python bench.py --steps 1 --warmup 0 --skip-cpu --skip-traversal --e2e-steps 1"
    $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
        -k regex:"absorbStage|peelStage|propagateStage|launchStage" -c 192 --csv --log-file gpurun_out/r02_final_stage_traffic.csv $CMD > gpurun_out/ncu.log 2>&1
    echo "ncu rc=$?"; python tools/ncu_traffic.py gpurun_out/r02_final_stage_traffic.csv | tee gpurun_out/r02_final_stage_traffic_summary.json ;;
pathtraffic)
    CMD="python bench.py --steps 1 --warmup 1 --packages 200000 --rays 4194304 --skip-cpu"
    $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
        -k regex:"pathFillKernel|pathCountKernel|pathCapacityKernel" -c 64 --csv --log-file gpurun_out/r02_final_path_traffic.csv $CMD > gpurun_out/ncu.log 2>&1
    echo "ncu rc=$?"; python tools/ncu_traffic.py gpurun_out/r02_final_path_traffic.csv | tee gpurun_out/r02_final_path_traffic_summary.json
    cat gpurun_out/plain.log | python -c "
import json,sys
for line in sys.stdin.read().splitlines():
    if line.startswith('{'):
        t = json.loads(line)['traversal_roofline']; print('rays', t['rays'], 'packet_steps', t['packet_steps'], 'bytes', t['bytes'])" ;;
full)
    $SMALL > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-4} -c ${4:-2} -f -o gpurun_out/r02_final_$5 $SMALL > gpurun_out/ncu.log 2>&1
    echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log; python tools/ncu_summary.py gpurun_out/r02_final_$5.ncu-rep > gpurun_out/r02_final_$5_ncu.txt 2>&1; head -60 gpurun_out/r02_final_$5_ncu.txt ;;
fullcfg)
    # full <kernel regex> <skip> <count> <tag> <bench.py arguments...>: one full capture on another configuration
    re=$2; sk=$3; ct=$4; tag=$5; shift 5
    CMD="python bench.py --steps 1 --warmup 1 --skip-cpu $*"
    $CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$re" -s $sk -c $ct -f -o gpurun_out/r02_final_$tag $CMD > gpurun_out/ncu.log 2>&1
    echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log; python tools/ncu_summary.py gpurun_out/r02_final_$tag.ncu-rep > gpurun_out/r02_final_${tag}_ncu.txt 2>&1; head -80 gpurun_out/r02_final_${tag}_ncu.txt ;;
esac
