#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2l}
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 400 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_train_launches.csv python tools/prof_train.py learned > gpurun_out/${T}_train_launches.log 2>&1
grep -v "^$" gpurun_out/${T}_gputest.log | tail -5; cat gpurun_out/${T}_lvis_variants.json; tail -2 gpurun_out/${T}_train_launches.log
