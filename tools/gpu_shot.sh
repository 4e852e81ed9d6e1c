#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -q -k 'sphere_renderer or grazing' > gpurun_out/parity_sphere.log 2>&1
echo "exit $?" >> gpurun_out/parity_sphere.log
tail -12 gpurun_out/parity_sphere.log
