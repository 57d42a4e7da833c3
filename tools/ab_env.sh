#!/bin/bash
# A/B of one environment switch on the device-resident agent step.  Usage: bash tools/ab_env.sh VAR=val [VAR=val ...]
for kv in "baseline=1" "$@"; do
  env "$kv" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-batched --no-roofline \
      --gpu-baseline none --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$kv', round(d['ms_per_step'],3), 'ms/step; unet+controlnet', round(d['unet_ms_per_step'],3), 'ms; launches', d['launches_per_step'])"
done
