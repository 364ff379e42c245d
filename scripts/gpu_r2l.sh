#!/bin/bash
# round-2 session L: rate-scaled tolerances (explicit and Rosenbrock), 3-resident plan — GPU tests, smoke, bench lines of every config
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2l_pytest.log; tail -3 gpurun_out/r2l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; tail -2 gpurun_out/r2l_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_c2.json 2> gpurun_out/r2l_c2.err; head -c 400 gpurun_out/r2l_c2.json; echo
python bench.py --members 160000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_c2_160k.json 2> gpurun_out/r2l_c2_160k.err; head -c 300 gpurun_out/r2l_c2_160k.json; echo
python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_c3.json 2> gpurun_out/r2l_c3.err; head -c 300 gpurun_out/r2l_c3.json; echo
python bench.py --config 5 --members 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_c5m8.json 2> gpurun_out/r2l_c5m8.err; head -c 300 gpurun_out/r2l_c5m8.json; echo
