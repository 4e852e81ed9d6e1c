#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2p}
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep -v "^$" gpurun_out/${T}_gputest.log | grep -i "bf16 train\|passed\|failed" | tail -6; head -c 300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
