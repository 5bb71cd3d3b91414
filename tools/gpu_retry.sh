#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> [gpus] -- runs tools/_call.sh on a GPU box, retrying while the pod is busy
T=${1:-900}; G=${2:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- 'bash tools/_call.sh'; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- 'bash tools/_call.sh'; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
