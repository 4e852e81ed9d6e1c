#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2h}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/scale_probe.py > gpurun_out/${T}_probe2.json 2> gpurun_out/${T}_probe2.err
NCCL_MAX_CTAS=2 NCCL_MIN_CTAS=1 timeout 400 $TR --master-port 29512 tools/scale_probe.py > gpurun_out/${T}_probe2_ctas2.json 2> gpurun_out/${T}_probe2_ctas2.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
cat gpurun_out/${T}_probe2.json gpurun_out/${T}_probe2_ctas2.json; tail -3 gpurun_out/${T}_probe2.err; head -c 500 gpurun_out/${T}_bench2.json; tail -3 gpurun_out/${T}_bench2.err
