#!/bin/bash
# A/B several builds of the library: tools/gpu_ab.sh libA.so libB.so ...  (paths relative to assistive_vr_gym_b200/)
for l in "$@"; do
  echo "== $l"
  AVG_B200_LIB=$PWD/assistive_vr_gym_b200/$l AVG_KERNEL_TIMES=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-episode 2>&1 >/dev/null | grep "avg kernel times" | head -1
  AVG_B200_LIB=$PWD/assistive_vr_gym_b200/$l python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-episode 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('VALUE', d['value'], 'e2e', d['e2e']['value'])"
done
