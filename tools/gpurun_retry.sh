#!/bin/bash
# usage: tools/gpurun_retry.sh <log> [--gpus N] <timeout> '<command>'   -- retries while the pod answers busy (exit 3)
log=$1; shift
gp=""; if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun $gp --timeout "$to" -- "$1" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "rc=3 (gave up)" >> "$log"
