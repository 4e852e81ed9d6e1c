#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2d}
timeout 700 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_lvis_variants.json; tail -3 gpurun_out/${T}_lvis_variants.err; grep -v "^$" gpurun_out/${T}_gputest.log | tail -8; head -c 3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
