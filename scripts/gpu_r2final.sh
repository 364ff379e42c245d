#!/bin/bash
# round-2 last check on a fresh box: what the driver runs at round end (GPU tests, smoke, both bench arms), plus the
# network configs through bench.py
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2fin_pytest.log; tail -2 gpurun_out/r2fin_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; tail -1 gpurun_out/r2fin_smoke.log
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2fin_ref.json 2> gpurun_out/r2fin_ref.err; head -c 200 gpurun_out/r2fin_ref.json; echo
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2fin_c2.json 2> gpurun_out/r2fin_c2.err; tail -c 200 gpurun_out/r2fin_c2.err; head -c 250 gpurun_out/r2fin_c2.json; echo
python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2fin_c3.json 2> gpurun_out/r2fin_c3.err; head -c 250 gpurun_out/r2fin_c3.json; echo
python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline --e2e-members 3 > gpurun_out/r2fin_c5.json 2> gpurun_out/r2fin_c5.err; head -c 250 gpurun_out/r2fin_c5.json; echo
