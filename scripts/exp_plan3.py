"""Experiment: placement plans of the latency-bound ensemble sizes — 3 resident blocks per SM (168 registers, nothing
chained) against 2 resident blocks with chained light blocks (184 registers) and the unplanned launches.
  python scripts/exp_plan3.py 10000 11000 12000 13000"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
eng = Engine(0)
VARIANTS = (("default", {}),
            ("2 resident, chained (r2 plan)", {"SIMPLYP_QUAD_MINBLOCKS": "2"}),
            ("3 resident, planned (any Q)", {"SIMPLYP_QUAD_MINBLOCKS": "3", "SIMPLYP_PLAN_QMAX": "147"}),
            ("3 resident, unplanned", {"SIMPLYP_QUAD_MINBLOCKS": "3", "SIMPLYP_SM_PLAN": "0"}))
KEYS = ("SIMPLYP_QUAD_MINBLOCKS", "SIMPLYP_PLAN_QMAX", "SIMPLYP_SM_PLAN")
ref = {}
for M in [int(x) for x in sys.argv[1:]] or [10000]:
    w = bench.build_workload("2004", M)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
    d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(w["member"][:M]); d_sc = eng.to_device(w["sc"][:M])
    d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
    V = w["obs_m"].shape[0]
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    for rep in range(2):
        for name, env in VARIANTS:
            for k in KEYS: os.environ.pop(k, None)
            os.environ.update(env)
            for _ in range(3):
                eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            s = stats.cpu().numpy()
            same = "" if M not in ref else (" bitwise equal" if np.array_equal(s, ref[M], equal_nan=True) else " DIFFERENT RESULTS")
            ref.setdefault(M, s)
            print("M=%6d %-34s %.3f ms  %.3e member-SC-days/s%s" % (M, name, ms, M * 366 / (ms * 1e-3), same), flush=True)
