#!/bin/bash
# tools/gpu_calls/retry.sh <timeout> <script>: gpurun with retries while the pod answers "busy" (exit code 3)
t=$1; s=$2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $t -- "bash $s" > gpurun_out/retry_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 200
done
tail -40 gpurun_out/retry_last.log
echo "retry.sh: rc=$rc after $i attempt(s)"
