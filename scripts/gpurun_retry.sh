#!/bin/bash
# usage: gpurun_retry.sh <timeout> <gpus> <command...> ; retries while the pod answers busy (exit 3)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1; fi
  rc=$?
  if grep -q "status=transient" gpurun_out/.retry.log; then sleep 120; continue; fi
  break
done
cat gpurun_out/.retry.log | tail -120
exit $rc
