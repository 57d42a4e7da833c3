#!/bin/bash
# A/B the step time under environment switches: bash tools/ab_bench.sh "VAR=val VAR2=val" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "=== $cfg"
  env $cfg python bench.py --steps 20 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-batched 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step', round(d['ms_per_step'],2), 'unet_ms', round(d['unet_ms_per_step'],2), 'steps/s', round(d['value'],2))
"
done
