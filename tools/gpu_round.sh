#!/bin/bash
# One GPU visit: per-file pytest processes (a faulting kernel cannot hide the other files) + logs.
# Usage (under gpurun):  bash tools/gpu_round.sh [extra pytest args]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe conv bf16" | tee gpurun_out/probe.log
timeout 300 python tests/probe_conv.py bf16 all >> gpurun_out/probe.log 2>&1; echo "exit $?" >> gpurun_out/probe.log
echo "== probe conv fp32" >> gpurun_out/probe.log
timeout 300 python tests/probe_conv.py fp32 all >> gpurun_out/probe.log 2>&1; echo "exit $?" >> gpurun_out/probe.log
tail -40 gpurun_out/probe.log | cut -c1-200
for f in tests/test_gpu_ops.py tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_e2e.py; do
  name=$(basename $f .py)
  echo "== $f"
  timeout 900 python -m pytest $f -m gpu -q -s -p no:cacheprovider --timeout 600 "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -12 gpurun_out/$name.log | cut -c1-200
done
