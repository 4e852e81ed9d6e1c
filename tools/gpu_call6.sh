#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2f}
timeout 700 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
NF_STAGEB_SINGLE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_stage_b" > gpurun_out/${T}_gputest_single.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest_single.log
NF_LVIS_SELF=0 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_reference_code.py -m gpu -q -p no:cacheprovider -k "stage_b or model_call or lvis or config3" > gpurun_out/${T}_gputest_issuer.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest_issuer.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_lvis_variants.json; tail -3 gpurun_out/${T}_lvis_variants.err; for f in gputest gputest_single gputest_issuer; do grep -v "^$" gpurun_out/${T}_$f.log | tail -4; done; head -c 600 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
