#!/bin/bash
for d in "$@"; do echo "AVG_DBG=$d"; AVG_DBG=$d AVG_KERNEL_TIMES=1 python bench.py --steps 13 --warmup 3 --no-cpu-baseline --no-episode 2>&1 >/dev/null | grep "avg kernel times" | tail -1; done
