#!/bin/bash
# One bounded GPU call: parity tests that go through nf_integrate_fwd + the integrate timing.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k 'model_call_matches_golden or l512 or config3 or config5 or ragged or full_size or empty_and or loss_matches or grazing' > gpurun_out/parity_integrate.log 2>&1
echo "parity exit $?" >> gpurun_out/parity_integrate.log
tail -15 gpurun_out/parity_integrate.log
timeout 60 python tools/time_integrate.py > gpurun_out/time_integrate_v3.json 2> gpurun_out/time_integrate.err
echo "integrate exit $?"
cat gpurun_out/time_integrate_v3.json
timeout 60 python tools/diag_integrate.py > gpurun_out/diag_integrate_v3.json 2>> gpurun_out/time_integrate.err
cat gpurun_out/diag_integrate_v3.json
