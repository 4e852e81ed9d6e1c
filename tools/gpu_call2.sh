#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2b}
timeout 120 python tools/tmem_bw.py > gpurun_out/${T}_tmem_bw.json 2> gpurun_out/${T}_tmem_bw.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_reference_code.py -m gpu -q -s -p no:cacheprovider -k "integrate or microfacet or network_call or trainer" > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 200 python tools/time_integrate.py > gpurun_out/${T}_time_integrate.json 2> gpurun_out/${T}_time_integrate.err
cat gpurun_out/${T}_tmem_bw.json; tail -3 gpurun_out/${T}_tmem_bw.err; grep -v "^$" gpurun_out/${T}_gputest.log | tail -15; cat gpurun_out/${T}_time_integrate.json
