#!/bin/bash
# usage: chain_after_pid.sh <pid> <logfile> <gpurun args...>  -- start a gpurun call once process <pid> has exited
pid=$1; shift
while kill -0 $pid 2>/dev/null; do sleep 15; done
sleep 5
exec scripts/gpu_retry.sh "$@"
