#!/bin/bash
# round-2 session H: GPU tests + smoke on the current build, then baseline timings of the network configs (before day epochs)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_pytest.log; tail -3 gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_c3.json 2> gpurun_out/r2h_c3.err; head -c 300 gpurun_out/r2h_c3.json; echo
python bench.py --config 5 --members 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_c5m8.json 2> gpurun_out/r2h_c5m8.err; head -c 300 gpurun_out/r2h_c5m8.json; echo
