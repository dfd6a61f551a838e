#!/bin/bash
# ncu evidence for a named regime of the episode (B200_PROFILING.md recipe): plain run first, then the launch list of the
# same command, then one --set full capture of the sub-step kernels.
# usage: tools/gpu_profile_regime.sh <tag> <env_id> <n_env> <regime> [full: 0|1]
tag=$1; env_id=${2:-ScratchItchJaco-v0}; n=${3:-196608}; regime=${4:-stagger}; full=${5:-1}
CMD="python tools/gpu_regime.py $env_id $n $regime 2"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
cat gpurun_out/plain_$tag.log | tail -1
ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_${tag}_1.log 2>&1
if [ "$full" = "1" ]; then
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"avg_(solve|dynamics|collide|narrow)" -s 8 -c 4 -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_${tag}_2.log 2>&1
fi
ls -la gpurun_out/ | grep $tag
