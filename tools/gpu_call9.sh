#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2i}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 tools/scale_probe.py > gpurun_out/${T}_probe2.json 2> gpurun_out/${T}_probe2.err
cat gpurun_out/${T}_probe2.json; tail -5 gpurun_out/${T}_probe2.err
