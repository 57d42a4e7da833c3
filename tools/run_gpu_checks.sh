#!/bin/bash
# Run each GPU test file in its own process with a timeout so one trapped kernel cannot take the rest down.
# Usage (on the GPU box): bash tools/run_gpu_checks.sh [files...]
mkdir -p gpurun_out
FILES=${@:-tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_attention.py tests/test_gpu_norm_elementwise.py tests/test_gpu_networks.py tests/test_gpu_step.py}
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in $FILES; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout ${GPU_TEST_TIMEOUT:-300} python -m pytest $f -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -5 gpurun_out/$name.log
done
