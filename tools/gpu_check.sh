#!/bin/bash
# usage (on the GPU box): tools/gpu_check.sh <tag>  -> GPU parity suite, then the bench at the headline size
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-episode 2>gpurun_out/bench_$tag.err | tee gpurun_out/bench_$tag.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('VALUE', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])"
