#!/bin/bash
# Run every GPU test module in its own process (a trapped kernel poisons the CUDA context) and
# keep logs under gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
status=0
for f in ${@:-tests/test_gpu_*.py}; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$name.log
  [ $rc -ne 0 ] && grep -E "^E  " gpurun_out/$name.log | cut -c1-400 | head -20
  [ $rc -ne 0 ] && status=1
done
exit $status
