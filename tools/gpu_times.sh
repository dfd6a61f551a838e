#!/bin/bash
# per-kernel times of the step at the headline size (AVG_KERNEL_TIMES serialises the host: not a bench number)
AVG_KERNEL_TIMES=1 python bench.py --steps 13 --warmup 3 --no-cpu-baseline --no-episode 2>&1 >/dev/null | grep "avg kernel times" | tail -1
