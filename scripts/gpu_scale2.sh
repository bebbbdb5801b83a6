#!/bin/bash
# multi-GPU bench at N = $1 (torchrun, one rank per GPU): headline C3 + C5 in the configs block
N=${1:-2}; T=${2:-r2s}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/${T}_n${N}.json 2> gpurun_out/${T}_n${N}.err
tail -3 gpurun_out/${T}_n${N}.err; cut -c1-400 gpurun_out/${T}_n${N}.json
