#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
# usage: [LEAN=1] [KEEP_VARIANTS=1] [GPUS=n] tools/gpurun_retry.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.."
for attempt in 1 2 3 4 5 6 7 8; do
  if [ -n "$LEAN" ]; then tools/gpurun_lean.sh "$1" "$2"; rc=$?
  else /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2"; rc=$?; fi
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: busy, retrying in 90 s"
  sleep 90
done
exit 3
