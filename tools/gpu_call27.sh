#!/bin/bash
mkdir -p gpurun_out
T=${1:-r3j}
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log; tail -n 3 gpurun_out/${T}_gputest.log
timeout 200 python tools/time_stage_b.py > gpurun_out/${T}_stage_b.json 2> gpurun_out/${T}_stage_b.err; cat gpurun_out/${T}_stage_b.json; tail -n 2 gpurun_out/${T}_stage_b.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -n 1 gpurun_out/${T}_smoke.log
