#!/bin/bash
# round-2 session Z (8 GPUs): fused all-gather over peer memory against NCCL, then the bench lines of configs 2 and 4
N=${1:-8}
mkdir -p gpurun_out
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR scripts/check_peer_gather.py $((10000 * N)) 20 > gpurun_out/r2z_peer_n$N.txt 2>&1; tail -1 gpurun_out/r2z_peer_n$N.txt
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2z_c2_n$N.json 2> gpurun_out/r2z_c2_n$N.err; tail -c 300 gpurun_out/r2z_c2_n$N.err; head -c 330 gpurun_out/r2z_c2_n$N.json; echo
SIMPLYP_GATHER=nccl $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2z_c2_n${N}_nccl.json 2> gpurun_out/r2z_c2_n${N}_nccl.err; head -c 330 gpurun_out/r2z_c2_n${N}_nccl.json; echo
$TR bench.py --gpus $N --config 4 --steps 5 --warmup 3 > gpurun_out/r2z_c4_n$N.json 2> gpurun_out/r2z_c4_n$N.err; head -c 330 gpurun_out/r2z_c4_n$N.json; echo
