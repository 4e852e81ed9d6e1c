#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2j}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "jitter or fused" > gpurun_out/${T}_gputest.log 2>&1
head -c 2500 gpurun_out/${T}_bench2.json; tail -4 gpurun_out/${T}_bench2.err; tail -3 gpurun_out/${T}_gputest.log
