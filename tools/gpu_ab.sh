#!/bin/bash
# A/B timing of library variants built by tools/ab_build.sh: tools/gpu_ab.sh "<regimes>" <n_env> <name>...   (env id: $AVG_AB_ENV)
regimes=$1; n=$2; shift 2
for name in "$@"; do
  for r in $regimes; do
    echo -n "[$name] "
    AVG_B200_LIB=build_ab/libavg_$name.so python tools/gpu_regime.py ${AVG_AB_ENV:-ScratchItchJaco-v0} $n $r 10 2>&1 | tail -1
  done
done
