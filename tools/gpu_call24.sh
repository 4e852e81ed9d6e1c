#!/bin/bash
mkdir -p gpurun_out
T=${1:-r3g}
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; grep -v "^$" gpurun_out/${T}_gputest.log | grep -i "chain\|passed\|failed\|error" | tail -8
timeout 120 python tools/prof_chain.py > gpurun_out/${T}_chain.log 2>&1; tail -n 2 gpurun_out/${T}_chain.log
timeout 300 python tools/time_train_step.py > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err; cat gpurun_out/${T}_train.json; tail -n 3 gpurun_out/${T}_train.err
