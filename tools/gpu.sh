#!/bin/bash
# usage: tools/gpu.sh <logfile> <timeout_s> <command...>   -- retries gpurun while the pod answers "transient"/busy
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "__DONE__ rc=$rc attempt=$attempt" >> "$log"
