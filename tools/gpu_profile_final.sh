#!/bin/bash
# launch list of two staggered-regime steps at the bench's batch + one full-set capture of the large-batch collide instance: tools/gpu_profile_final.sh <tag>
tag=$1
CMD="python tools/gpu_regime.py ScratchItchJaco-v0 393216 stagger 2"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
tail -1 gpurun_out/plain_$tag.log
ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_${tag}_1.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"avg_collide" -s 2 -c 1 -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_${tag}_2.log 2>&1
ls -la gpurun_out/ | grep $tag
