#!/bin/bash
# usage: tools/gpu_call.sh TAG TIMEOUT [--gpus N] -- 'command'   (retries while the pod answers busy; log in gpurun_out/call_TAG.log)
TAG=$1; shift; TO=$1; shift
EXTRA=()
while [ "$1" != "--" ]; do EXTRA+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO "${EXTRA[@]}" -- "$@" > gpurun_out/call_$TAG.log 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" gpurun_out/call_$TAG.log && [ $rc -ne 0 ]; then sleep 90; continue; fi
  break
done
tail -60 gpurun_out/call_$TAG.log
exit $rc
