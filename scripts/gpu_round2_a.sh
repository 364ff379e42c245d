#!/bin/bash
# round-2 GPU session A: parity suite, then one bench line per BASELINE config (1 GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; tail -c 600 gpurun_out/r2a_bench_c2.json
python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; tail -c 300 gpurun_out/r2a_bench_c3.err; head -c 400 gpurun_out/r2a_bench_c3.json
python bench.py --config 5 --steps 2 --warmup 3 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; tail -c 300 gpurun_out/r2a_bench_c5.err; head -c 400 gpurun_out/r2a_bench_c5.json
python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; tail -c 300 gpurun_out/r2a_bench_c4.err; head -c 400 gpurun_out/r2a_bench_c4.json
