#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 40); do
	/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
	rc=$?
	if [ $rc -ne 3 ]; then cat /tmp/gpurun_last.log | tail -80; exit $rc; fi
	sleep 120
done
echo "gave up"; exit 3
