#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy): tools/gpurun_retry.sh <log> [gpurun args...] -- '<command>'
LOG=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy\|answers busy" "$LOG"; then sleep 45; continue; fi
  break
done
echo "attempts: $attempt" >> "$LOG"
