#!/bin/bash
# Round-end ncu evidence (B200_PROFILING.md recipe): plain run first, then the launch list of the same command, then one
# --set full capture of each sub-step kernel.  usage: tools/gpu_profile.sh <tag>
tag=${1:-r1n}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-episode"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
# launches before the timed region: 1 reset-observation kernel + 3 warm-up steps x 44 launches
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -s 133 -c 88 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_${tag}_1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"avg_(solve|dynamics|collide|narrow)" -s 12 -c 4 -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_${tag}_2.log 2>&1
ls -la gpurun_out/ | tail -5
