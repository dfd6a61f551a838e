#!/bin/bash
AVG_DBG=${1:-32} python bench.py --steps 10 --warmup 3 --envs-per-gpu 32768 --no-cpu-baseline --no-episode 2>&1 >/dev/null | grep -A26 "narrowphase histogram"
