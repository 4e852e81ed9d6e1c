#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/time_integrate.py > gpurun_out/time_integrate_lb3.json 2> gpurun_out/time_integrate.err
echo "integrate exit $?"
cat gpurun_out/time_integrate_lb3.json
timeout 100 python -m pytest tests/test_gpu_parity.py -q -k 'model_call_matches_golden or config3 or config5 or ragged or grazing' 2>&1 | tail -3
