#!/bin/bash
# round-2 final check on a fresh box: what the driver runs at round end (GPU tests, smoke, both bench arms)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2z_pytest.log; tail -3 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; head -c 300 gpurun_out/r2z_ref.json; echo
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench.err; head -c 400 gpurun_out/r2z_bench.json; echo
python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; head -c 300 gpurun_out/r2z_bench_default.json; echo
