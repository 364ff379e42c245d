#!/bin/bash
# round-2 session Y (N GPUs): fused all-gather over peer memory against NCCL
N=${1:-2}
mkdir -p gpurun_out
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR scripts/check_peer_gather.py $((10000 * N + 3)) 10 > gpurun_out/r2y_peer_n$N.txt 2>&1; tail -4 gpurun_out/r2y_peer_n$N.txt
