#!/bin/bash
# usage: bash scripts/gpu_scale.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n${N}.json 2> gpurun_out/scale_n${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --config C5 > gpurun_out/scale_c5_n${N}.json 2>> gpurun_out/scale_n${N}.err
cut -c1-600 gpurun_out/scale_n${N}.json; cut -c1-600 gpurun_out/scale_c5_n${N}.json; tail -3 gpurun_out/scale_n${N}.err
