#!/bin/bash
# launch list (+ optional full capture of one kernel) of two steps at a small batch: tools/gpu_profile_small.sh <tag> <env_id> <n_env> <regime> [kernel regex for --set full]
tag=$1; env_id=$2; n=$3; regime=$4; kre=$5
CMD="python tools/gpu_regime.py $env_id $n $regime 2"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
tail -1 gpurun_out/plain_$tag.log
ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_${tag}_1.log 2>&1
if [ -n "$kre" ]; then
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"$kre" -s 6 -c 2 -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_${tag}_2.log 2>&1
fi
