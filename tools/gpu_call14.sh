#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2o}
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_zz_gpu_scripts.py -m gpu -q -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 400 python tools/time_train_step.py > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err
tail -4 gpurun_out/${T}_gputest.log; cat gpurun_out/${T}_train.json; tail -3 gpurun_out/${T}_train.err
