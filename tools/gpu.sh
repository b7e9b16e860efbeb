#!/bin/bash
# usage: tools/gpu.sh <logfile> <timeout_s> <command...>   -- retries gpurun while the pod answers "transient"/busy
log=$1; shift; to=$1; shift
make -C imp_release_b200/csrc -j8 > /tmp/imp_make.log 2>&1 || { echo 'BUILD FAILED'; grep -E 'error' /tmp/imp_make.log | head; exit 1; }
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "__DONE__ rc=$rc attempt=$attempt" >> "$log"
