#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2n}
N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench${N}.json 2> gpurun_out/${T}_bench${N}.err
head -c 3000 gpurun_out/${T}_bench${N}.json; tail -5 gpurun_out/${T}_bench${N}.err
