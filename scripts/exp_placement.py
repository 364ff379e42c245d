"""Experiment: planned placement (SIMPLYP_SM_PLAN) and warps led by one heavy member (SIMPLYP_SOLO_WARPS) — time per
pass and bitwise equality of the results.  usage: exp_solo.py M  plan:solo [plan:solo ...]   (solo 'd' = default)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
eng = Engine(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cfgs = sys.argv[2:] or ["0:0", "1:0", "1:24", "1:d", "1:48", "0:0", "1:d"]
w = bench.build_workload("2004", M)
opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
opt.pilot_days = int(os.environ.get("PILOT_DAYS", "0"))
d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(w["member"][:M]); d_sc = eng.to_device(w["sc"][:M])
d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
V = w["obs_m"].shape[0]
po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
ref = None
for cfg in cfgs:
    plan, solo = cfg.split(":")
    os.environ["SIMPLYP_SM_PLAN"] = plan
    if solo == "d": os.environ.pop("SIMPLYP_SOLO_WARPS", None)
    else: os.environ["SIMPLYP_SOLO_WARPS"] = solo
    stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    for _ in range(3):
        eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = stats.cpu().numpy(); dg = diag.cpu().numpy()
    if ref is None: ref = (st, dg)
    same = np.array_equal(ref[0], st, equal_nan=True) and np.array_equal(ref[1], dg)
    print("M=%d plan=%s solo=%3s  median %.3f ms  min %.3f ms  status %d  bitwise equal to the first: %s" %
          (M, plan, solo, float(np.median(ts)), min(ts), int(dg[:, 0, 3].max()), same), flush=True)
