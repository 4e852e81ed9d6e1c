#!/bin/bash
mkdir -p gpurun_out
T=${1:-r3h}
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gputest.log
timeout 120 python tools/prof_chain.py > gpurun_out/${T}_chain.log 2>&1
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
grep -v "^$" gpurun_out/${T}_gputest.log | grep -i "passed\|failed\|error\|pytest rc" | tail -5; tail -n 2 gpurun_out/${T}_chain.log; head -c 330 gpurun_out/${T}_bench.json; echo; tail -n 2 gpurun_out/${T}_bench.err; tail -n 1 gpurun_out/${T}_smoke.log
