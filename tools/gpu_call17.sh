#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2s}
timeout 300 python tools/time_sigma_variants.py > gpurun_out/${T}_sigma_variants.json 2> gpurun_out/${T}_sigma_variants.err
timeout 500 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err
timeout 200 python tools/sigma_timeline.py f16 > gpurun_out/${T}_timeline_f16.txt 2> gpurun_out/${T}_timeline.err
cat gpurun_out/${T}_sigma_variants.json gpurun_out/${T}_lvis_variants.json; head -18 gpurun_out/${T}_timeline_f16.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; tail -n 4 gpurun_out/${T}_gputest.log
