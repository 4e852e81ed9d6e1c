#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2u}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "stage_b or fused or lvis or brdf or chunk" > gpurun_out/${T}_gputest.log 2>&1; tail -n 6 gpurun_out/${T}_gputest.log
timeout 300 python tools/time_stage_b.py > gpurun_out/${T}_stage_b.json 2> gpurun_out/${T}_stage_b.err; cat gpurun_out/${T}_stage_b.json; tail -n 3 gpurun_out/${T}_stage_b.err
timeout 500 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err; cat gpurun_out/${T}_lvis_variants.json
