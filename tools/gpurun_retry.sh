#!/bin/bash
# gpurun_retry.sh <timeout_s> <command...> — retry a gpurun call while the pod answers "transient" (nothing charged).
# Gives up after 25 attempts. The log of the accepted call is printed at the end.
T=$1; shift
for i in $(seq 1 25); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gpurun_retry: still transient after 25 attempts"
exit 3
