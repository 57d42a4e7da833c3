#!/bin/bash
# Timing ablation: how much of the agent step sits behind one kernel class?  Each line re-runs the device-resident step
# with that class's launches dropped (outputs uninitialised: timing only).  Usage (GPU box): bash tools/ablate.sh
for skip in 0 1 2 3; do
  GENIMA_B200_SKIP=$skip timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-batched --no-roofline \
      --gpu-baseline none --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('skip=$skip (1: gn_apply, 2: attention)', round(d['ms_per_step'],3), 'ms/step; unet+controlnet', round(d['unet_ms_per_step'],3), 'ms; launches', d['launches_per_step'])"
done
