#!/bin/bash
# local helper: retry a gpurun call while the pod answers "busy / draining" (nothing is charged for those)
# usage: tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] pod busy"; sleep 150; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
