#!/bin/bash
# A/B helper for gpurun: ab.sh "<label>=<ENV=VAL ...>" ... ; runs each variant twice, interleaved
for rep in 1 2; do
  for v in "$@"; do
    label="${v%%=*}"; envs="${v#*=}"
    env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --train-steps 0 --cpt-steps 0 --sweep-tokens 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), round(d['roofline']['achieved'],1), d['gpu_launches'])"
  done
done
