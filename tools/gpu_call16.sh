#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2q}
timeout 300 python tools/time_sigma_variants.py > gpurun_out/${T}_sigma_variants.json 2> gpurun_out/${T}_sigma_variants.err
timeout 500 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err
cat gpurun_out/${T}_sigma_variants.json gpurun_out/${T}_lvis_variants.json; tail -3 gpurun_out/${T}_sigma_variants.err gpurun_out/${T}_lvis_variants.err
