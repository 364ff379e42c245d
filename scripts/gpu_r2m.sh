#!/bin/bash
# round-2 session M: epoch sweep of networks — invariance test, A/B of epoch lengths on configs 3 and 5, full GPU tier
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "epoch_sweep or network" 2>&1 | tail -8 > gpurun_out/r2m_pytest_a.log; tail -3 gpurun_out/r2m_pytest_a.log
timeout 900 python scripts/exp_epochs.py 3:64 5:8 3:256 5:1 -- 0 128 256 512 1024 2048 > gpurun_out/r2m_epochs.log 2>&1; cat gpurun_out/r2m_epochs.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
