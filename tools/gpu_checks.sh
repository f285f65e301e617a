#!/bin/bash
# Run each GPU test file in its own process (a device trap poisons the CUDA context of that process only),
# with a timeout, and keep the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
status=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -q -m gpu -x --tb=short > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "== $f rc=$rc"; tail -n 25 "gpurun_out/$name.log"
  [ $rc -ne 0 ] && status=1
done
exit $status
