#!/bin/bash
# One bounded GPU session (gpurun): every step under its own timeout, outputs under gpurun_out/.
mkdir -p gpurun_out
T=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 600 python tools/check_lvis_variants.py > gpurun_out/${T}_ew8.json 2> gpurun_out/${T}_ew8.err
timeout 300 python tools/time_sigma_variants.py > gpurun_out/${T}_sigma_variants.json 2> gpurun_out/${T}_sigma_variants.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
tail -5 gpurun_out/${T}_gputest.log; cat gpurun_out/${T}_ew8.json; cat gpurun_out/${T}_sigma_variants.json; tail -2 gpurun_out/${T}_smoke.log
