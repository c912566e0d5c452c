#!/bin/bash
# wait for the running gpurun call to finish, then run the 2-GPU checks
while /usr/local/graft/bin/gpurun --status | grep -q '"in_flight": 1'; do sleep 20; done
sleep 5
scripts/gpu_retry.sh gpurun_out/call_r02_multi2.txt --gpus 2 --timeout 2400 -- 'bash scripts/r02_multi2.sh'
