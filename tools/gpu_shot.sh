#!/bin/bash
# One bounded GPU call: the scripts end-to-end test on the real kernels + the integrate timing.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k 'model_call_matches_golden or ragged or process_view or config5' > gpurun_out/parity_subset.log 2>&1
echo "parity subset exit $?" >> gpurun_out/parity_subset.log
tail -3 gpurun_out/parity_subset.log
timeout 200 python -m pytest tests/test_zz_gpu_scripts.py -x -q > gpurun_out/e2e_scripts.log 2>&1
echo "e2e exit $?" >> gpurun_out/e2e_scripts.log
timeout 60 python tools/time_integrate.py > gpurun_out/time_integrate.json 2> gpurun_out/time_integrate.err
echo "integrate exit $?" >> gpurun_out/e2e_scripts.log
tail -5 gpurun_out/e2e_scripts.log
