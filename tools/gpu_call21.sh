#!/bin/bash
mkdir -p gpurun_out
T=${1:-r3a}
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_bench_step.csv python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/${T}_launches_bench_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lvis_tc3_kernel|sigma_tc_kernel|integrate_kernel|point_tc_kernel" -c 9 -o gpurun_out/${T}_prof python tools/prof_round2.py > gpurun_out/${T}_prof.log 2>&1
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv > gpurun_out/${T}_prof_raw.csv 2>/dev/null
rm -f gpurun_out/${T}_prof.ncu-rep
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
grep -v "^$" gpurun_out/${T}_gputest.log | tail -4; head -c 400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_prof.log; tail -1 gpurun_out/${T}_smoke.log
