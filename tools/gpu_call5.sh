#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2e}
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
# launch list of two views (fused + unfused Stage B)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_round2.py > gpurun_out/${T}_launches.log 2>&1
# full capture of the tensor kernels + rendering-equation kernel of the same command
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc2_kernel|sigma_tc_kernel|integrate_kernel|point_tc_kernel" -c 9 -o gpurun_out/${T}_prof python tools/prof_round2.py > gpurun_out/${T}_prof.log 2>&1
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv > gpurun_out/${T}_prof_raw.csv 2>/dev/null
ls -la gpurun_out/${T}_prof.ncu-rep
grep -v "^$" gpurun_out/${T}_gputest.log | tail -6; head -c 1500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_prof.log
