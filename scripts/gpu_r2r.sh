#!/bin/bash
# round-2 GPU session R (final build) (N GPUs of one box): sharding invariance, bench config 2 (weak) and config 4 (strong, 10^6 members)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR scripts/check_sharding.py 100003 > gpurun_out/r2r_shard_n$N.txt 2>&1; tail -2 gpurun_out/r2r_shard_n$N.txt
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2r_bench_c2_n$N.json 2> gpurun_out/r2r_bench_c2_n$N.err; tail -c 400 gpurun_out/r2r_bench_c2_n$N.err; head -c 330 gpurun_out/r2r_bench_c2_n$N.json; echo
$TR bench.py --gpus $N --config 4 --steps 5 --warmup 3 > gpurun_out/r2r_bench_c4_n$N.json 2> gpurun_out/r2r_bench_c4_n$N.err; tail -c 400 gpurun_out/r2r_bench_c4_n$N.err; head -c 330 gpurun_out/r2r_bench_c4_n$N.json; echo
