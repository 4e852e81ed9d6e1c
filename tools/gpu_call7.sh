#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2g}
timeout 500 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
NF_STAGEB_SINGLE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_stage_b" > gpurun_out/${T}_gputest_single.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest_single.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_lvis_variants.json; tail -3 gpurun_out/${T}_lvis_variants.err; for f in gputest gputest_single; do grep -v "^$" gpurun_out/${T}_$f.log | tail -3; done; head -c 400 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
