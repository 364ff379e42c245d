"""Experiment: number of led warps (SIMPLYP_SOLO_WARPS: warps of one heavy member and seven light ones) of the planned
launch.   python scripts/exp_solo.py 10000 -- 0 24 48 96 160 240"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
args = sys.argv[1:]
cut = args.index("--") if "--" in args else len(args)
sizes = [int(a) for a in args[:cut]] or [10000]
settings = args[cut + 1:] or ["24", "48", "96"]
eng = Engine(0)
for M in sizes:
    w = bench.build_workload("2004", M)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
    d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(w["member"][:M]); d_sc = eng.to_device(w["sc"][:M])
    d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
    V = w["obs_m"].shape[0]
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    ref = None
    for rep in range(2):
        for sset in settings:
            for kv in sset.split(","):
                k, v = kv.split("=") if "=" in kv else ("SIMPLYP_SOLO_WARPS", kv)
                os.environ[k] = v
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
            same = "" if ref is None else (" bitwise equal" if np.array_equal(s, ref, equal_nan=True) else " DIFFERENT RESULTS")
            ref = s if ref is None else ref
            print("M=%6d %-40s %.3f ms  %.3e member-SC-days/s%s" % (M, sset, ms, M * 366 / (ms * 1e-3), same), flush=True)
