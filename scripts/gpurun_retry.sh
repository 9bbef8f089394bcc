#!/bin/bash
# usage: scripts/gpurun_retry.sh <tag> <timeout_s> <command...>   -- retries while the pod answers "busy" (exit 3)
tag=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/$tag.stdout 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
