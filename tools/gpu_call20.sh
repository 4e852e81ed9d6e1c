#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2y}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; tail -n 4 gpurun_out/${T}_gputest.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
head -c 400 gpurun_out/${T}_bench.json; tail -n 3 gpurun_out/${T}_bench.err
