#!/bin/bash
# round-2 session Q (final build, one GPU): what the driver runs at round end (GPU tests, smoke, both bench arms), the
# bench lines of every config, the ncu launch list of the bench command and the full capture of its main pass
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2q_pytest.log; tail -3 gpurun_out/r2q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; tail -2 gpurun_out/r2q_smoke.log
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2q_ref_c2.json 2> gpurun_out/r2q_ref_c2.err; head -c 300 gpurun_out/r2q_ref_c2.json; echo
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_c2.json 2> gpurun_out/r2q_c2.err; tail -c 300 gpurun_out/r2q_c2.err; head -c 400 gpurun_out/r2q_c2.json; echo
python bench.py > gpurun_out/r2q_c2_default.json 2> gpurun_out/r2q_c2_default.err; head -c 300 gpurun_out/r2q_c2_default.json; echo
python bench.py --members 160000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_c2_160k.json 2> gpurun_out/r2q_c2_160k.err; head -c 300 gpurun_out/r2q_c2_160k.json; echo
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_c4.json 2> gpurun_out/r2q_c4.err; head -c 300 gpurun_out/r2q_c4.json; echo
python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2q_c3.json 2> gpurun_out/r2q_c3.err; head -c 300 gpurun_out/r2q_c3.json; echo
python bench.py --config 3 --members 256 --steps 2 --warmup 3 --no-cpu-baseline --e2e-members 8 > gpurun_out/r2q_c3_m256.json 2> gpurun_out/r2q_c3_m256.err; head -c 300 gpurun_out/r2q_c3_m256.json; echo
python bench.py --config 5 --steps 2 --warmup 3 --e2e-members 2 > gpurun_out/r2q_c5.json 2> gpurun_out/r2q_c5.err; head -c 300 gpurun_out/r2q_c5.json; echo
python bench.py --config 5 --members 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_c5_m1.json 2> gpurun_out/r2q_c5_m1.err; head -c 300 gpurun_out/r2q_c5_m1.json; echo
for c in 3 5; do python bench.py --impl reference --config $c --steps 3 --warmup 1 > gpurun_out/r2q_ref_c$c.json 2> gpurun_out/r2q_ref_c$c.err; head -c 250 gpurun_out/r2q_ref_c$c.json; echo; done
# profiles (each after its command has run without ncu above)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench.csv $CMD > gpurun_out/r2q_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:simplyp_quad_kernel -s 7 -c 1 -o gpurun_out/r02b_quad_cal_M10000 $CMD > gpurun_out/r2q_ncu_full.log 2>&1
tail -2 gpurun_out/r2q_ncu_full.log
